// dist_comm.hpp — the multi-GPU exchange used by sharded training: NCCL over NVLink/NVSwitch,
// one rank per GPU. NCCL is loaded lazily with dlopen so the single-GPU path has no dependency
// on it (and so the library binds to whichever libnccl.so.2 the host process already loaded,
// e.g. the one bundled with PyTorch).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifndef BPE_EMUL
#include <cuda_runtime.h>
#include <dlfcn.h>
#endif

namespace bpe {

#ifndef BPE_EMUL
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string* err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { if (err) *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
        GetUniqueId = (int (*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
        AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather) {
            if (err) *err = "libnccl is missing required symbols";
            return false;
        }
        return true;
    }
};
inline NcclApi& nccl_api() { static NcclApi api; return api; }

inline bool dist_unique_id(void* out128, std::string* err) {
    NcclApi& api = nccl_api();
    if (!api.load(err)) return false;
    ncclUniqueId id;
    int rc = api.GetUniqueId(&id);
    if (rc != ncclSuccess) { if (err) *err = "ncclGetUniqueId failed"; return false; }
    memcpy(out128, &id, 128);
    return true;
}

enum { DIST_U32_SUM = 0, DIST_U64_MIN = 1, DIST_U64_MAX = 2, DIST_U64_SUM = 3 };
struct DistComm {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = 0;
    bool init(int r, int w, const void* uid, cudaStream_t s, std::string* err) {
        NcclApi& api = nccl_api();
        if (!api.load(err)) return false;
        ncclUniqueId id;
        memcpy(&id, uid, 128);
        int rc = api.CommInitRank(&comm, w, id, r);
        if (rc != ncclSuccess) { if (err) *err = std::string("ncclCommInitRank: ") + (api.GetErrorString ? api.GetErrorString(rc) : "?"); return false; }
        rank = r; world = w; stream = s;
        return true;
    }
    void destroy() { if (comm) { nccl_api().CommDestroy(comm); comm = nullptr; } }
    // in-place all-reduce of a device buffer, enqueued on the context's stream
    bool allreduce(void* buf, size_t count, int kind) {
        if (world == 1) return true;
        int dt = kind == DIST_U32_SUM ? ncclUint32 : ncclUint64;
        int op = (kind == DIST_U32_SUM || kind == DIST_U64_SUM) ? ncclSum : (kind == DIST_U64_MIN ? ncclMin : ncclMax);
        return nccl_api().AllReduce(buf, buf, count, dt, op, comm, stream) == ncclSuccess;
    }
    bool allgather_bytes(const void* src, void* dst, size_t bytes_per_rank) {
        return nccl_api().AllGather(src, dst, bytes_per_rank, ncclUint8, comm, stream) == ncclSuccess;
    }

    // ---- peer-memory mailboxes (NVLink): the per-step exchange of sharded training without NCCL ----
    PeerSet peers;             // valid when peer_ok
    bool peer_ok = false;
    void* mbox_local = nullptr;
    void* peer_mapped[MAX_PEERS] = {nullptr};
    uint32_t epoch_base = 0;   // arrival flags only ever grow, across training calls
    static size_t mbox_slot_words() { return (size_t)2 * 65552 + 16 + 16 * MAX_PEERS; }
    // Allocate this rank's mailbox, exchange IPC handles through NCCL, map every peer's mailbox. Collective: every
    // rank takes part in the all-gather and in the final agreement whatever happened locally (a rank whose
    // allocation or mapping failed sends a zeroed handle and votes 0), so either all ranks use the mailboxes or
    // all of them use the NCCL all-reduce — a split would leave the ranks waiting for each other in different
    // protocols. BPE_TEST_PEER_FAIL_RANK=<r> makes rank r fail locally (tests).
    bool init_peers(std::string* err) {
        if (world < 2 || world > MAX_PEERS) { if (err) *err = "peer exchange supports 2..8 ranks"; return false; }
        const size_t slot = mbox_slot_words();
        const size_t words = (size_t)2 * ((size_t)2 * world * slot) + 64;  // 8-byte cells
        bool local_ok = true;
        std::string why;
        auto local_fail = [&](const char* msg) { if (local_ok) why = msg; local_ok = false; cudaGetLastError(); };
        const char* tf = getenv("BPE_TEST_PEER_FAIL_RANK");
        if (tf && atoi(tf) == rank) local_fail("forced failure (BPE_TEST_PEER_FAIL_RANK)");
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof mine);
        if (local_ok && cudaMalloc(&mbox_local, words * 4) != cudaSuccess) { mbox_local = nullptr; local_fail("mailbox allocation failed"); }
        if (local_ok) {
            cudaMemset(mbox_local, 0, words * 4);
            if (cudaIpcGetMemHandle(&mine, mbox_local) != cudaSuccess) { memset(&mine, 0, sizeof mine); local_fail("cudaIpcGetMemHandle failed"); }
        }
        // [world handles | my handle | vote]: one buffer for the all-gather and the agreement
        void* d_h = nullptr;
        const size_t hb = sizeof(mine);
        if (cudaMalloc(&d_h, hb * (size_t)(world + 1) + 8) != cudaSuccess) {
            // cannot even take part in the collective: this is the one failure that cannot be agreed on
            if (err) *err = "device allocation for the handle exchange failed";
            cudaGetLastError();
            destroy_peers();
            return false;
        }
        cudaMemcpy((char*)d_h + hb * world, &mine, hb, cudaMemcpyHostToDevice);
        bool coll_ok = allgather_bytes((char*)d_h + hb * world, d_h, hb);
        cudaStreamSynchronize(stream);
        std::vector<cudaIpcMemHandle_t> all((size_t)world);
        cudaMemcpy(all.data(), d_h, hb * world, cudaMemcpyDeviceToHost);
        if (!coll_ok) local_fail("all-gather of IPC handles failed");
        static const cudaIpcMemHandle_t zero_handle = {};
        for (int p = 0; p < world && local_ok; p++) {
            void* base = mbox_local;
            if (p != rank) {
                if (memcmp(&all[(size_t)p], &zero_handle, hb) == 0) { local_fail("a peer has no mailbox"); break; }
                if (cudaIpcOpenMemHandle(&base, all[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    local_fail("cudaIpcOpenMemHandle failed (no peer access?)");
                    break;
                }
                peer_mapped[p] = base;
            }
            peers.mbox[p] = (uint32_t*)base;
            peers.flags[p] = (uint32_t*)base + (size_t)2 * ((size_t)2 * world * slot);
        }
        // agreement: minimum of the local verdicts
        unsigned long long vote = local_ok ? 1ull : 0ull;
        void* d_vote = (char*)d_h + hb * (size_t)(world + 1);
        cudaMemcpy(d_vote, &vote, 8, cudaMemcpyHostToDevice);
        const bool red_ok = allreduce(d_vote, 1, DIST_U64_MIN);
        cudaStreamSynchronize(stream);
        cudaMemcpy(&vote, d_vote, 8, cudaMemcpyDeviceToHost);
        cudaFree(d_h);
        if (!red_ok || vote == 0) {
            if (err) *err = local_ok ? "a peer could not set up its mailbox" : why;
            destroy_peers();
            return false;
        }
        peers.slot_words = (uint32_t)slot;
        peer_ok = true;
        return true;
    }
    void destroy_peers() {
        for (int p = 0; p < MAX_PEERS; p++) if (peer_mapped[p]) { cudaIpcCloseMemHandle(peer_mapped[p]); peer_mapped[p] = nullptr; }
        if (mbox_local) { cudaFree(mbox_local); mbox_local = nullptr; }
        peer_ok = false;
    }
};
#else
// emulation build (tests only): the exchange is delegated to a callback so that a world_size-2
// gloo test on the CPU can drive the same host logic
enum { DIST_U32_SUM = 0, DIST_U64_MIN = 1, DIST_U64_MAX = 2, DIST_U64_SUM = 3 };
typedef int (*dist_allreduce_cb)(void* buf, size_t count, int kind);
inline bool dist_unique_id(void* out128, std::string*) { memset(out128, 0, 128); return true; }
struct DistComm {
    int rank = 0, world = 1;
    dist_allreduce_cb cb = nullptr;
    PeerSet peers;
    bool peer_ok = false;
    uint32_t epoch_base = 0;
    void destroy_peers() {}
    static size_t mbox_slot_words() { return (size_t)2 * 65552 + 16 + 16 * MAX_PEERS; }
    static size_t mbox_words(int world) { return (size_t)2 * ((size_t)2 * world * mbox_slot_words()) + 64; }  // 8-byte cells
    // tests: the ranks are processes on one host and `base` is a zero-filled shared-memory segment that holds the
    // mailboxes of all ranks back to back (the GPU build maps the peers' device memory with cudaIpc instead)
    bool init_peers_shm(void* base, size_t bytes) {
        if (world < 2 || world > MAX_PEERS || bytes < (size_t)world * mbox_words(world) * 4) return false;
        for (int p = 0; p < world; p++) {
            uint32_t* mb = (uint32_t*)base + (size_t)p * mbox_words(world);
            peers.mbox[p] = mb;
            peers.flags[p] = mb + (size_t)2 * ((size_t)2 * world * mbox_slot_words());
        }
        peers.slot_words = (uint32_t)mbox_slot_words();
        peer_ok = true;
        return true;
    }
    bool init(int, int, const void*, int, std::string* err) { if (err) *err = "no NCCL in the emulation build"; return false; }
    void destroy() {}
    bool allreduce(void* buf, size_t count, int kind) {
        if (world == 1) return true;
        return cb && cb(buf, count, kind) == 0;
    }
};
#endif

}  // namespace bpe
