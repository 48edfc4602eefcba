// tests/emul/cuda_emul.hpp — TEST INFRASTRUCTURE ONLY.
//
// A minimal single-OS-thread emulation of the CUDA execution model, just enough to run the
// kernels of zig-bpe_b200/csrc on the build container (which has no GPU) so their *logic*
// can be checked against the oracle before GPU time is spent. Blocks run one after another;
// the threads of a block are ucontext fibers that yield at __syncthreads(). Warp intrinsics are
// emulated by yielding until all lanes of the (emulated) warp have posted their operand.
// Nothing in the product links or loads this: libbpe_b200.so is compiled by nvcc only, and the
// emulated build is a separate library (tests/emul/libbpe_emul.so) used by `-m "not gpu"` tests.
#pragma once
#include <ucontext.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace emul {
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
void sync_threads();
void yield_thread();
void launch(dim3 grid, dim3 block, const std::function<void()>& body, bool uses_sync);
uint32_t warp_exchange(uint32_t v, int op, uint32_t arg);  // see cuda_emul.cpp
extern uint64_t g_launches;
}  // namespace emul

#define threadIdx (emul::g_threadIdx)
#define blockIdx (emul::g_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)
static inline void __syncthreads() { emul::sync_threads(); }
#define BPE_SPIN_YIELD() emul::yield_thread()
#define BPE_GRID_DEP_WAIT() ((void)0)
#define BPE_GRID_DEP_LAUNCH() ((void)0)
namespace emul { extern int g_sync_or_acc[2]; extern int g_sync_or_phase; }
static inline int __syncthreads_or(int pred) {
    // two alternating accumulators so that back-to-back calls do not interfere
    int ph = emul::g_sync_or_phase;
    if (pred) emul::g_sync_or_acc[ph] = 1;
    emul::sync_threads();
    int r = emul::g_sync_or_acc[ph];
    if (threadIdx.x == 0) { emul::g_sync_or_phase = ph ^ 1; emul::g_sync_or_acc[ph ^ 1] = 0; }
    emul::sync_threads();
    return r;
}
// a warp-wide rendezvous (only meaningful in fiber kernels; plain-loop kernels never call it)
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emul::warp_exchange(0u, 5 /* EMUL_ANY */, 0); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
// 8-byte cells shared between processes (peer mailboxes in the tests): single 64-bit accesses, as on the GPU
static inline void __stcg(uint2* p, uint2 v) { __atomic_store_n(reinterpret_cast<uint64_t*>(p), (uint64_t)v.x | ((uint64_t)v.y << 32), __ATOMIC_RELEASE); }
static inline uint2 __ldcv(const uint2* p) { const uint64_t w = __atomic_load_n(reinterpret_cast<const uint64_t*>(p), __ATOMIC_ACQUIRE); return uint2{(uint32_t)w, (uint32_t)(w >> 32)}; }

// ---- atomics (single OS thread: plain read-modify-write) --------------------------------
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <class T> static inline T atomicSub(T* p, T v) { T o = *p; *p = (T)(o - v); return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = (T)(o | v); return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((__uint128_t)a * (__uint128_t)b) >> 64);
}
static inline unsigned __vcmpeq2(unsigned a, unsigned b) {
    unsigned r = 0;
    if ((a & 0xffffu) == (b & 0xffffu)) r |= 0xffffu;
    if ((a >> 16) == (b >> 16)) r |= 0xffff0000u;
    return r;
}
// packed 16-bit unsigned minimum (VIMNMX.U16x2 / VIMNMX3.U16x2 on the GPU)
static inline unsigned __vminu2(unsigned a, unsigned b) {
    const unsigned lo = std::min(a & 0xffffu, b & 0xffffu), hi = std::min(a >> 16, b >> 16);
    return lo | (hi << 16);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    const uint64_t w = ((uint64_t)hi << 32) | lo;
    return (unsigned)(w >> (shift & 31u));
}
static inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vminu2(a, b), c); }
// warp intrinsics (blockDim.x must be a multiple of 32 where these are used)
enum { EMUL_BALLOT = 0, EMUL_SHFL = 1, EMUL_SHFL_UP = 2, EMUL_SHFL_DOWN = 3, EMUL_SHFL_XOR = 4, EMUL_ANY = 5 };
static inline unsigned __ballot_sync(unsigned, int pred) { return emul::warp_exchange(pred ? 1u : 0u, EMUL_BALLOT, 0); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int lane) { return emul::warp_exchange(v, EMUL_SHFL, (uint32_t)lane); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) { return emul::warp_exchange(v, EMUL_SHFL_UP, d); }
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) { return emul::warp_exchange(v, EMUL_SHFL_DOWN, d); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return emul::warp_exchange(v, EMUL_SHFL_XOR, (uint32_t)m); }
static inline int __any_sync(unsigned, int pred) { return (int)emul::warp_exchange(pred ? 1u : 0u, EMUL_ANY, 0); }

// ---- runtime API subset -----------------------------------------------------------------
typedef int cudaError_t;
typedef int cudaStream_t;
struct EmulEvent { std::chrono::steady_clock::time_point t; };
typedef EmulEvent* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p); strcpy(p->name, "emul"); p->major = 10; p->minor = 0; p->multiProcessorCount = 4;
    p->totalGlobalMem = (size_t)8 << 30; return 0;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmulEvent(); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0;
}

#define BPE_LAUNCH(kern, grid, block, stream, ...) \
    emul::launch(dim3(grid), dim3(block), [&]() { kern(__VA_ARGS__); }, true)
#define BPE_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), [&]() { kern(__VA_ARGS__); }, true)
#define BPE_LAUNCH_PDL(kern, grid, block, stream, pdl, ...) \
    emul::launch(dim3(grid), dim3(block), [&]() { kern(__VA_ARGS__); }, true)
namespace emul { extern uint32_t g_dyn_smem[64 * 1024]; }
static inline uint32_t* bpe_dyn_smem() { return emul::g_dyn_smem; }  // up to 256 KB of "dynamic shared memory"
// TMA bulk copy + mbarrier: the copy happens at issue time, so waiting is a no-op
static inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
static inline void mbar_fence_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
static inline void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
static inline void mbar_wait(uint64_t*, uint32_t) {}
static inline void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
static inline void cp_async_wait_all() {}
// kernels that never call __syncthreads()/warp intrinsics: run threads as a plain loop
#define BPE_LAUNCH_NS(kern, grid, block, stream, ...) \
    emul::launch(dim3(grid), dim3(block), [&]() { kern(__VA_ARGS__); }, false)
