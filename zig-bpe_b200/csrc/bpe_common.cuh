// bpe_common.cuh — shared types, constants and small device helpers of the B200 BPE engine.
//
// Data layout in HBM:
//   tok[]      resident token sequence, u16 ids (the reference's own width,
//              src/basic_tokenizer.zig:162), HOLE = 0xFFFF marks a merged-away slot. train() can
//              only create ids 256..65534 (vocabSize is u16 and ids stop at vocabSize-1), so
//              0xFFFF is never a live id during training; encode switches to the u32 instantiation
//              when a merge list mentions id 65535.
//   pair table open-addressing hash table key -> count, key = first | second << 16 (the same 4
//              bytes the reference hashes, src/basic_tokenizer.zig:40-43)
//   zig cnt[]  per-home-slot population of the *reference's* hash table (Zig 0.13 AutoHashMap,
//              Wyhash low bits), kept so the reference's tie-break (lowest slot among the
//              max-count pairs, src/basic_tokenizer.zig:193,291-303) can be decided on the GPU.
#pragma once

#include <stddef.h>
#include <stdint.h>

#ifdef BPE_EMUL
#include "cuda_emul.hpp"
#else
#include <cuda_runtime.h>
#define BPE_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BPE_LAUNCH_NS(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BPE_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BPE_SPIN_YIELD() ((void)0)
// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may start while its predecessor in the
// stream is still draining; it must not touch anything the predecessor writes before BPE_GRID_DEP_WAIT() returns
// (the predecessor has completed and its writes are visible then). BPE_GRID_DEP_LAUNCH() lets the successor start.
#define BPE_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define BPE_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
template <class... KArgs, class... Args>
static inline cudaError_t bpe_launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, int pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define BPE_LAUNCH_PDL(kern, grid, block, stream, pdl, ...) bpe_launch_pdl(kern, (grid), (block), 0, (stream), (pdl), __VA_ARGS__)
extern __shared__ __align__(128) uint32_t bpe_dyn_smem_[];
__device__ __forceinline__ uint32_t* bpe_dyn_smem() { return bpe_dyn_smem_; }

// ---- TMA bulk copy (global -> shared) completing on an mbarrier: cp.async.bulk, sm_90+ -----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS): no register holds the data while it travels
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

namespace bpe {

// ---- geometry ----------------------------------------------------------------------------
#ifndef BPE_TILE
#define BPE_TILE 8192  // token slots per tile (one CTA pass): 16 KB of u16 ids, 4 x 16 B per thread
#endif
#ifndef BPE_THREADS
#define BPE_THREADS 256
#endif
constexpr int TILE = BPE_TILE;
constexpr int THREADS = BPE_THREADS;
constexpr int SPT = TILE / THREADS;  // slots per thread
static_assert(TILE % THREADS == 0, "tile must be a multiple of the block size");
static_assert(SPT % 8 == 0 || SPT == 4 || SPT == 2 || SPT == 1, "slots per thread");

constexpr int MAXTIE = 1024;          // tied keys listed per step; more -> table replay
constexpr uint32_t EMPTY_KEY = 0xFFFFFFFFu;
constexpr uint32_t ZCHUNK = 1024;     // reference-table slots per chunk in the occupancy scan

template <class T> struct TokTraits;
template <> struct TokTraits<uint16_t> { static constexpr uint32_t hole = 0xFFFFu; };
template <> struct TokTraits<uint32_t> { static constexpr uint32_t hole = 0xFFFFFFFFu; };

__host__ __device__ __forceinline__ uint32_t pair_key(uint32_t first, uint32_t second) {
    return (first & 0xFFFFu) | (second << 16);
}

// error flags raised by kernels (StepCtl::err)
enum : uint32_t {
    ERR_TABLE_FULL = 1u << 0,
    ERR_KEY_MISSING = 1u << 1,
    ERR_COUNT_UNDERFLOW = 1u << 2,
    ERR_VERIFY_MISMATCH = 1u << 3,
    ERR_ZCNT_OVERFLOW = 1u << 4,
    ERR_PEER_TIMEOUT = 1u << 5,   // a peer's deltas did not arrive (multi-GPU peer exchange)
    ERR_COUNT_OVERFLOW = 1u << 6, // a pair occurs more often than the 32-bit counts of the pair table can hold
};

// tie-resolution verdicts (StepCtl::tie_status)
enum : uint32_t { TIE_NONE = 0, TIE_FAST_OK = 1, TIE_NEED_REPLAY = 2 };

// why the device-driven train loop stopped advancing (StepCtl::halt); kernels are no-ops while set
enum : uint32_t {
    H_NONE = 0,
    H_DONE = 1,      // all requested merges learned
    H_HEAVY = 2,     // heavy-key list no longer covers the maximum (or overflowed): host rebuilds it
    H_ZCAP = 3,      // zcnt models another capacity of the reference's table: host rebuilds it
    H_REPLAY = 4,    // tie that needs the full table replay on the host
    H_CLASSIC = 5,   // candidate-scan merge path cannot take this step (A == B, or too many candidates): host runs the tiled pass
    H_AEQB = 6,      // the chosen merge has first == second: the host runs this step with the run-chaining halo pass
};
// StepCtl::flags
enum : uint32_t { F_FORCE_REPLAY = 1u << 0, F_CHECK_TIES = 1u << 1, F_HALT_AEQB = 1u << 2 };

struct MergeRec { uint32_t key; uint32_t count; };  // one learned merge: pair key + its count

// Device-resident control block. The train loop is driven from the device: select / tie kernels
// choose the merge, halo / merge / apply kernels read it from here, and the host only reads this
// block back once per batch of steps.
struct StepCtl {
    // current merge (A,B) -> X
    uint32_t A, B, X;
    // argmax result of the current step
    uint32_t max_count;
    uint32_t ntied;        // number of keys with count == max_count (may exceed MAXTIE)
    // table state
    uint32_t live_keys;    // keys with count > 0 (= distinct adjacent pairs, D)
    uint32_t n_inserted;   // keys ever inserted (load factor of our table)
    // encode passes: merged occurrences accumulate here
    uint32_t cntXX;
    uint32_t cntAB;
    uint32_t err;
    // tie fast path
    uint32_t tie_status;
    uint32_t tie_winner;   // key
    uint32_t last_pair_pos;  // slot index of the left token of the last pair (replay edge case)
    uint32_t verify_mismatch;
    uint32_t n_heavy;      // entries appended to the heavy-key list (may exceed its capacity)
    uint32_t hist_nonzero; // distinct byte pairs found by the initial count
    uint32_t zpop_max;     // upper bound on the keys homed in any one chunk of the reference table
    uint32_t apply_done;   // CTAs of apply_kernel that have finished (last one runs the next selection)
    // device-driven stepping
    uint32_t step;         // merges learned so far (= index of the merge being decided)
    uint32_t want_steps;
    uint32_t halt;         // H_*
    uint32_t theta;        // heavy-list threshold (0 = list invalid)
    uint32_t zcap;         // capacity of the reference table that zcnt models
    uint32_t need_tie;     // the current step's tie is being settled by the zig_* kernels
    uint32_t flags;        // F_*
    uint32_t last_merged;  // occurrences merged by the last applied step
    uint32_t cand_n;       // candidate-scan merge path: queued A positions ...
    uint32_t w_n;          // ... and token writes decided for them
    uint32_t pass_step;    // 1 + index of the last step whose merge pass really ran (not halted), written by that pass
    uint32_t pad0;
    unsigned long long live_tokens;
    unsigned long long fast_ties;   // tie steps settled on the device
    unsigned long long local_live;  // live tokens of this GPU's shard (multi-GPU; live_tokens is the global count)
    uint32_t tie_keys[MAXTIE];
};

// Multi-GPU: what one shard tells the others about its ends, exchanged once per merge step (it rides
// in the same all-reduce as the merge deltas: every rank fills only its own slot, the rest are 0).
struct EdgeInfo {
    uint32_t first[3];  // first live tokens of the shard
    uint32_t nfirst;    // how many of them exist (0..3)
    uint32_t last[2];   // last live tokens, nearest the end first
    uint32_t nlast;     // 0..2
    uint32_t run_tok;   // the last live token ...
    uint32_t run_len;   // ... and how many consecutive live copies of it end the shard
    uint32_t all_same;  // the whole shard consists of that token only
    uint32_t pad[6];
};
static_assert(sizeof(EdgeInfo) == 64, "EdgeInfo is 16 words");

// Multi-GPU peer exchange (NVLink peer memory instead of an NCCL all-reduce): every rank owns a mailbox
// [2 parities][world slots][xw words] + one arrival flag per sender; peers write their deltas straight
// into it and raise their flag, the consumer sums the slots while it applies them.
constexpr int MAX_PEERS = 8;
struct PeerSet {
    uint32_t* mbox[MAX_PEERS];   // mailbox base of every rank (own entry = local pointer): [2 parities][world senders][slot_words] cells
    uint32_t* flags[MAX_PEERS];  // (spare words behind the mailbox; the cells carry their own tags)
    uint32_t slot_words;         // cells per mailbox slot; a cell is 8 bytes {value, epoch}
};

// Per-tile neighbourhood, produced by halo_kernel before each merge pass: what a tile needs to
// know about live tokens outside itself. hole value = "no such token" (sequence start/end).
template <class TokT> struct TileHalo {
    TokT l2, l1;      // last two live tokens before the tile (l1 nearest)
    TokT r0, r1, r2;  // first three live tokens after the tile
    uint32_t runA;    // A==B steps: number of consecutive live A immediately before the tile
};

// ---- pair table --------------------------------------------------------------------------
struct PairTable {
    uint32_t* keys;    // EMPTY_KEY when free
    uint32_t* counts;
    uint32_t mask;     // capacity - 1
};

__host__ __device__ __forceinline__ uint32_t tbl_hash(uint32_t k) {
    k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
    return k;
}

// returns slot of key, or EMPTY_KEY if absent
__device__ __forceinline__ uint32_t tbl_find(const PairTable& t, uint32_t key) {
    uint32_t s = tbl_hash(key) & t.mask;
    for (uint32_t probes = 0; probes <= t.mask; probes++) {
        uint32_t k = ((volatile const uint32_t*)t.keys)[s];
        if (k == key) return s;
        if (k == EMPTY_KEY) return EMPTY_KEY;
        s = (s + 1) & t.mask;
    }
    return EMPTY_KEY;
}

// returns slot of key, inserting it (count 0) if absent; EMPTY_KEY if the table is full
__device__ __forceinline__ uint32_t tbl_find_or_insert(const PairTable& t, uint32_t key, uint32_t* n_inserted) {
    uint32_t s = tbl_hash(key) & t.mask;
    for (uint32_t probes = 0; probes <= t.mask; probes++) {
        uint32_t k = ((volatile const uint32_t*)t.keys)[s];
        if (k == key) return s;
        if (k == EMPTY_KEY) {
            uint32_t old = atomicCAS(&t.keys[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY) { atomicAdd(n_inserted, 1u); return s; }
            if (old == key) return s;
        }
        s = (s + 1) & t.mask;
    }
    return EMPTY_KEY;
}

// ---- Zig 0.13 std.hash.Wyhash of the 4 key bytes, seed 0 ------
// len = 4 -> a = b = (x << 32) | x, x = the key bytes as a little-endian u32 = pair_key().
__host__ __device__ __forceinline__ void wy_mum(uint64_t a, uint64_t b, uint64_t* lo, uint64_t* hi) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    __uint128_t r = (__uint128_t)a * (__uint128_t)b;
    *lo = (uint64_t)r;
    *hi = (uint64_t)(r >> 64);
#endif
}
__host__ __device__ __forceinline__ uint64_t zig_hash_pair(uint32_t key) {
    const uint64_t S0 = 0xa0761d6478bd642fULL, S1 = 0xe7037ed1a0b428dbULL;
    uint64_t lo, hi;
    wy_mum(S0, S1, &lo, &hi);       // seed 0: state0 = 0 ^ mix(0 ^ S0, S1)
    const uint64_t st0 = lo ^ hi;
    uint64_t a = ((uint64_t)key << 32) | (uint64_t)key;
    uint64_t b = a;
    a ^= S1;
    b ^= st0;
    wy_mum(a, b, &lo, &hi);
    uint64_t x = lo ^ S0 ^ 4ULL, y = hi ^ S1;
    wy_mum(x, y, &lo, &hi);
    return lo ^ hi;
}

// capacity of the reference's table holding d distinct keys when no growth is pending
// (Appendix A.3/A.4): smallest power of two >= 8 with floor(cap*80/100) >= d.
__host__ __device__ __forceinline__ uint32_t zig_cap_for(uint32_t d) {
    uint32_t cap = 8;
    while ((uint64_t)cap * 80 / 100 < d) cap <<= 1;
    return cap;
}
__host__ __device__ __forceinline__ uint32_t zig_max_load(uint32_t cap) { return (uint32_t)((uint64_t)cap * 80 / 100); }

// population of reference home slots: zcnt = two u16 counters per u32 word, zpop = keys homed in
// each chunk of ZCHUNK slots (lets the tie resolver bound how far a probe run can reach)
struct ZigPop { uint32_t* zcnt; uint32_t* zpop; uint32_t zmask; };
__device__ __forceinline__ void zcnt_add(const ZigPop& z, uint32_t key, int delta, StepCtl* ctl) {
    uint32_t h = (uint32_t)zig_hash_pair(key) & z.zmask;
    uint32_t sh = (h & 1u) * 16u;
    if (delta > 0) {
        uint32_t old = atomicAdd(&z.zcnt[h >> 1], 1u << sh);
        if (((old >> sh) & 0xFFFFu) == 0xFFFFu) atomicOr(&ctl->err, (uint32_t)ERR_ZCNT_OVERFLOW);
        uint32_t op = atomicAdd(&z.zpop[h / ZCHUNK], 1u);
        if (op + 1 > ((volatile uint32_t*)&ctl->zpop_max)[0]) atomicMax(&ctl->zpop_max, op + 1);
    } else {
        atomicSub(&z.zcnt[h >> 1], 1u << sh);
        atomicSub(&z.zpop[h / ZCHUNK], 1u);
    }
}
__device__ __forceinline__ uint32_t zcnt_get(const uint32_t* zcnt, uint32_t h) {
    return (zcnt[h >> 1] >> ((h & 1u) * 16u)) & 0xFFFFu;
}

}  // namespace bpe
