// bpe_engine.cu — host side of libbpe_b200.so: context, train / encode / decode drivers and the
// C ABI of include/bpe_b200.h. Kernels live in bpe_kernels.cuh.
//
// Reference call stack replaced (src/basic_tokenizer.zig): train :140-153 -> expandVocabulary
// :172-205 -> {generate :234-255, count :257-278, sort :280-306, replace :207-232} per merge;
// encode :71-88; decode :90-138.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>
#include <unordered_set>

#include "../../include/bpe_b200.h"
#include "bpe_kernels.cuh"
#include "bpe_segenc.cuh"
#include "bpe_tilenc.cuh"
#include "bpe_groupenc.cuh"
#include "tiebreak_host.hpp"
#include "dist_comm.hpp"

using namespace bpe;

// Device allocations are recycled through a per-context cache: a training call allocates several
// GB (two token buffers, the pair table) and cudaMalloc/cudaFree of that size cost more than a
// hundred merge steps. A freed block is reused by a later request of at most twice... its size.
struct BufCache {
    struct Blk { void* p; size_t bytes; };
    std::vector<Blk> free_list;
    size_t cached_bytes = 0;
    static constexpr size_t MAX_CACHED = (size_t)24 << 30;
    void* take(size_t n, size_t* got) {
        size_t best = (size_t)-1;
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].bytes >= n && free_list[i].bytes <= 2 * n + 4096 && (best == (size_t)-1 || free_list[i].bytes < free_list[best].bytes)) best = i;
        if (best == (size_t)-1) return nullptr;
        Blk b = free_list[best];
        free_list.erase(free_list.begin() + (long)best);
        cached_bytes -= b.bytes;
        *got = b.bytes;
        return b.p;
    }
    void give(void* p, size_t bytes) {
        if (bytes < 4096 || cached_bytes + bytes > MAX_CACHED) { cudaFree(p); return; }
        free_list.push_back(Blk{p, bytes});
        cached_bytes += bytes;
    }
    void clear() { for (Blk& b : free_list) cudaFree(b.p); free_list.clear(); cached_bytes = 0; }
    // give memory back to the device until at most `keep` bytes stay cached (largest blocks first)
    void trim(size_t keep) {
        while (cached_bytes > keep && !free_list.empty()) {
            size_t big = 0;
            for (size_t i = 1; i < free_list.size(); i++) if (free_list[i].bytes > free_list[big].bytes) big = i;
            cudaFree(free_list[big].p);
            cached_bytes -= free_list[big].bytes;
            free_list.erase(free_list.begin() + (long)big);
        }
    }
};

// -----------------------------------------------------------------------------------------
// context
// -----------------------------------------------------------------------------------------
struct bpe_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    std::string err;
    // options
    long verify_recount = 0, force_slow_tiebreak = 0, check_tiebreak = 0, compact_pct = 93, encode_compact_pct = 85, table_log2 = 0,
         max_steps = 0, time_phases = 0, profile = 0, debug = 0, merge_impl = 0, xchg_impl = 0, encode_impl = 0, encode_grid = 24, encode_filter = 1, encode_geom = 4, encode_seg_min_steps = 450, encode_tile = 7936, encode_try_tiles = 1, fuse_halo = 1, count_limit_log2 = 32, stream_chunk_mb = 128, stream_chunk_bytes = 0, merge_filter = 0, merge_direct = 3, merge_prestage = 1, merge_pairfilter = 1, merge_loop = 6, merge_dbuf = 0, encode_fuse = 0, batch_steps = 16, pdl = 1, cache_max_mb = 8192;
    int num_sms = 148;
    DistComm dist;  // world == 1 when single GPU
    uint64_t launches = 0;
    BufCache cache;
    std::vector<cudaEvent_t> ev_pool;  // profiling events, created on first use
    struct DecodeCache* decode_cache = nullptr;  // decode tables of the last merge list seen (built once per list)
    cudaStream_t copy_in = 0, copy_out = 0;      // streaming encode: host -> device text chunks, device -> host ids (created on first use)
    void* h_ctl = nullptr;                       // pinned StepCtl read-back buffer of the train loop
};

static std::string g_create_err;

static int fail(bpe_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, BPE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static thread_local BufCache* tl_cache = nullptr;  // set by the API entry points for the call's context

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;      // requested size
    size_t real = 0;       // size of the underlying block
    BufCache* from = nullptr;
    ~DevBuf() { release(); }
    void release() {
        if (!p) return;
        if (from) from->give(p, real); else cudaFree(p);
        p = nullptr;
    }
    cudaError_t alloc(size_t n) {
        release();
        bytes = n;
        from = tl_cache;
        if (from) {
            void* q = from->take(n ? n : 1, &real);
            if (q) { p = q; return cudaSuccess; }
        }
        real = n ? n : 1;
        cudaError_t e = cudaMalloc(&p, real);
        if (e != cudaSuccess && from) {  // out of memory: drop the cache and retry once
            cudaGetLastError();
            from->clear();
            e = cudaMalloc(&p, real);
        }
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
    template <class T> T* as() const { return (T*)p; }
};
struct HostBuf {
    void* p = nullptr;
    ~HostBuf() { if (p) cudaFreeHost(p); }
    cudaError_t alloc(size_t n) { return cudaMallocHost(&p, n ? n : 1); }
    template <class T> T* as() const { return (T*)p; }
};

static inline unsigned grid_for(size_t n, int threads, unsigned max_blocks = 148 * 16) {
    size_t b = (n + (size_t)threads - 1) / (size_t)threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}
static inline size_t round_up(size_t n, size_t m) { return (n + m - 1) / m * m; }
static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// -----------------------------------------------------------------------------------------
// token sequence + per-tile scratch shared by train and encode
// -----------------------------------------------------------------------------------------
template <class TokT> struct Sequence {
    DevBuf buf[2];
    int cur = 0;
    size_t cap_slots = 0;   // allocated slots per buffer (multiple of TILE)
    size_t n_slots = 0;     // slots in use (multiple of TILE)
    uint64_t live = 0;      // live tokens (host view; may lag the device by one step)
    size_t dense_end = 0;   // slots at/after this index are padding holes (set by the load and by every compaction)
    DevBuf halo, run_local, run_full, tile_live, tile_off, total, done_counter;
    TokT* tok() const { return buf[cur].as<TokT>(); }
    TokT* other() const { return buf[cur ^ 1].as<TokT>(); }
    uint32_t ntiles() const { return (uint32_t)(n_slots / TILE); }
};

template <class TokT>
static int seq_init(bpe_ctx* ctx, Sequence<TokT>& sq, const uint8_t* d_text, size_t n) {
    sq.cap_slots = round_up(n ? n : 1, TILE);
    sq.n_slots = sq.cap_slots;
    sq.live = n;
    sq.dense_end = n;
    CU(sq.buf[0].alloc(sq.cap_slots * sizeof(TokT)));
    CU(sq.buf[1].alloc(sq.cap_slots * sizeof(TokT)));
    size_t nt = sq.cap_slots / TILE;
    CU(sq.halo.alloc(nt * sizeof(TileHalo<TokT>)));
    CU(sq.run_local.alloc(nt * sizeof(uint32_t)));
    CU(sq.run_full.alloc(nt));
    CU(sq.tile_live.alloc(nt * sizeof(uint32_t)));
    CU(sq.tile_off.alloc(nt * sizeof(unsigned long long)));
    CU(sq.total.alloc(sizeof(unsigned long long)));
    CU(sq.done_counter.alloc(4));
    CU(cudaMemsetAsync(sq.done_counter.p, 0, 4, ctx->stream));
    BPE_LAUNCH_NS(widen_kernel<TokT>, grid_for(sq.cap_slots, 256), 256, ctx->stream, d_text, n, sq.tok(), sq.cap_slots);
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// squeeze holes out; *new_live receives the live count measured on the device
template <class TokT>
static int seq_compact(bpe_ctx* ctx, Sequence<TokT>& sq, uint64_t* new_live) {
    const uint32_t nt = sq.ntiles();
    BPE_LAUNCH(tile_count_kernel<TokT>, nt, THREADS, ctx->stream, sq.tok(), sq.tile_live.template as<uint32_t>());
    BPE_LAUNCH(tile_scan_kernel, 1, THREADS, ctx->stream, sq.tile_live.template as<uint32_t>(), nt,
               sq.tile_off.template as<unsigned long long>(), sq.total.template as<unsigned long long>());
    BPE_LAUNCH((compact_scatter_kernel<TokT, TokT>), nt, THREADS, ctx->stream, sq.tok(),
               sq.tile_off.template as<unsigned long long>(), sq.other());
    ctx->launches += 3;
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, sq.total.p, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    size_t ns = round_up(total ? (size_t)total : 1, TILE);
    if (ns > (size_t)total) {
        BPE_LAUNCH_NS(fill_holes_kernel<TokT>, grid_for(ns - total, 256), 256, ctx->stream, sq.other(), (size_t)total, ns);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    sq.cur ^= 1;
    sq.n_slots = ns;
    sq.live = total;
    sq.dense_end = (size_t)total;
    if (new_live) *new_live = total;
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// pair table (+ reference-home population)
// -----------------------------------------------------------------------------------------
struct TableMem {
    DevBuf keys, counts, zcnt;  // zcnt: [cap words of u16 home counters][2*cap/ZCHUNK+1 words of chunk populations]
    uint32_t cap = 0;
    uint32_t zcap = 0;  // capacity of the reference's table that zcnt currently models (0 = stale)
    PairTable view() const { PairTable t; t.keys = keys.as<uint32_t>(); t.counts = counts.as<uint32_t>(); t.mask = cap - 1; return t; }
    uint32_t zmask() const { return zcap ? zcap - 1 : 0; }
    size_t zpop_words() const { return (size_t)2 * cap / ZCHUNK + 1; }
    ZigPop zig() const { ZigPop z; z.zcnt = zcnt.as<uint32_t>(); z.zpop = zcnt.as<uint32_t>() + cap; z.zmask = zmask(); return z; }
};

static int table_alloc(bpe_ctx* ctx, TableMem& tm, uint32_t cap) {
    tm.cap = cap;
    tm.zcap = 0;
    CU(tm.keys.alloc((size_t)cap * 4));
    CU(tm.counts.alloc((size_t)cap * 4));
    CU(tm.zcnt.alloc(((size_t)cap + tm.zpop_words()) * 4));  // 2*cap u16 home counters: enough for zig cap <= 2*cap
    CU(cudaMemsetAsync(tm.keys.p, 0xFF, (size_t)cap * 4, ctx->stream));
    CU(cudaMemsetAsync(tm.counts.p, 0, (size_t)cap * 4, ctx->stream));
    return BPE_OK;
}

// make zcnt model the reference table capacity for d live keys (rebuild only when it changes)
static int table_ensure_zcnt(bpe_ctx* ctx, TableMem& tm, StepCtl* d_ctl, uint32_t d) {
    uint32_t want = zig_cap_for(d);
    if (want == tm.zcap) return BPE_OK;
    if ((size_t)want > (size_t)2 * tm.cap) return fail(ctx, BPE_ERR_INTERNAL, "reference table capacity %u exceeds zcnt buffer", want);
    tm.zcap = want;
    CU(cudaMemsetAsync(tm.zcnt.p, 0, std::max<size_t>((size_t)want * 2, 4), ctx->stream));
    CU(cudaMemsetAsync(tm.zcnt.as<uint32_t>() + tm.cap, 0, tm.zpop_words() * 4, ctx->stream));
    CU(cudaMemsetAsync(&d_ctl->zpop_max, 0, 4, ctx->stream));
    BPE_LAUNCH_NS(zig_rebuild_kernel, grid_for(tm.cap, 256), 256, ctx->stream, tm.view(), tm.zig(), d_ctl);
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// train
// -----------------------------------------------------------------------------------------
// chained event marks: the interval between two consecutive marks is attributed to the bucket
// of the earlier one. Nothing synchronises until the pool is exhausted or finish().
enum { K_INIT = 0, K_ARGMAX, K_TIE, K_REPLAY, K_HALO, K_MERGE, K_APPLY, K_COMPACT, K_TABLE, K_HOSTGAP, K_NB = 12 };
struct EvProfile {
    int level = 0;  // 0 off, 1 all buckets, 2 merge kernel only (2 records per step), 3 merge kernel of every 32nd step
    cudaStream_t st = 0;
    std::vector<cudaEvent_t>* ev = nullptr;  // pool owned by the context (reused across calls)
    std::vector<int> bucket;
    size_t used = 0;
    bool sample_now = true;  // level 3: set per step by the caller
    double* ms = nullptr;
    uint64_t* calls = nullptr;
    void init(int lvl, cudaStream_t s, std::vector<cudaEvent_t>* pool, double* ms_out, uint64_t* calls_out) {
        level = lvl; st = s; ms = ms_out; calls = calls_out; ev = pool;
        if (!level) return;
        const size_t POOL = 1u << 15;
        while (ev->size() < POOL) { cudaEvent_t e; cudaEventCreate(&e); ev->push_back(e); }
        bucket.resize(ev->size());
    }
    void drain(bool keep_last) {  // resolve recorded intervals; optionally carry the last mark over
        if (used == 0) return;
        cudaEventSynchronize((*ev)[used - 1]);
        for (size_t i = 0; i + 1 < used; i++) {
            if (bucket[i] < 0 || (level >= 2 && bucket[i] != K_MERGE)) continue;
            float t = 0;
            cudaEventElapsedTime(&t, (*ev)[i], (*ev)[i + 1]);
            ms[bucket[i]] += t;
            calls[bucket[i]]++;
        }
        if (keep_last) { std::swap((*ev)[0], (*ev)[used - 1]); bucket[0] = bucket[used - 1]; used = 1; }
        else used = 0;
    }
    void mark(int b) {
        if (!level) return;
        if (level >= 2 && b != K_MERGE && b != K_APPLY) return;
        if (level == 3 && !sample_now) return;
        if (used == ev->size()) drain(true);
        cudaEventRecord((*ev)[used], st);
        bucket[used] = b;
        used++;
    }
    void finish() {
        if (!level) return;
        if (level == 1) mark(-1);
        drain(false);
    }
};

__global__ void hist_nonzero_kernel(const unsigned long long* __restrict__ hist, StepCtl* ctl) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536u && hist[i]) atomicAdd(&ctl->hist_nonzero, 1u);
}
__global__ void reset_argmax_kernel(StepCtl* ctl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->max_count = 0; ctl->ntied = 0; ctl->tie_status = TIE_NONE; }
}
// host-side decisions written into the control block (one tiny launch instead of several copies)
__global__ void ctl_set_kernel(StepCtl* ctl, MergeRec* rec, int set_theta, uint32_t theta, int set_zcap, uint32_t zcap,
                               int clear_halt, int commit, uint32_t key, uint32_t count) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    if (set_theta) ctl->theta = theta;
    if (set_zcap) ctl->zcap = zcap;
    if (commit) { commit_merge(ctl, rec, key, count); ctl->need_tie = 0; }
    if (clear_halt) { ctl->halt = H_NONE; ctl->cand_n = 0; ctl->w_n = 0; }
}

struct TrainRun {
    bpe_ctx* ctx;
    Sequence<uint16_t> sq;
    TableMem tm;
    DevBuf delta, hist, ctl, rec, firstpos, recount, live_chk, heavy, cand, wr;
    uint32_t cand_cap = 0;  // candidate-scan merge path: queue capacity
    uint32_t vcap = 0;   // stride of the cntL / cntR halves of `delta`
    uint32_t theta = 0;  // heavy-list threshold (0 = list invalid)
    HeavyList hl() const { HeavyList h; h.slots = heavy.as<uint32_t>(); h.cap = (uint32_t)(heavy.bytes / 4); h.theta = theta; return h; }
    void* h_ctl_p = nullptr;  // pinned read-back buffer, owned by the context (page-locking per call costs tens of ms)
    bpe_stats_t st;
    EvProfile prof;
    uint64_t aeqb_steps = 0;                // steps that took the run-chaining halo pass (first == second)
    std::vector<uint32_t> pending_samples;  // profile 3: sampled steps enqueued in the current batch
    uint64_t sampled_noop = 0;              // sampled launches that turned out to be no-ops (batch halted earlier)
    StepCtl* d_ctl() const { return ctl.as<StepCtl>(); }
    StepCtl* hc() const { return (StepCtl*)h_ctl_p; }
    MergeRec* d_rec() const { return rec.as<MergeRec>(); }
    uint32_t* cntL() const { return delta.as<uint32_t>(); }
    uint32_t* cntR() const { return delta.as<uint32_t>() + vcap; }
    uint32_t* nxx() const { return delta.as<uint32_t>() + 2 * (size_t)vcap; }
    uint32_t* nab() const { return delta.as<uint32_t>() + 2 * (size_t)vcap + 1; }
    // multi-GPU: the exchange buffer is [deltas | pad | one EdgeInfo per rank]; one all-reduce per step sums it
    size_t edge_off = 0;  // word offset of the EdgeInfo slots inside `delta`
    EdgeInfo* edges() const { return ctx->dist.world > 1 ? reinterpret_cast<EdgeInfo*>(delta.as<uint32_t>() + edge_off) : nullptr; }
    size_t exchange_words() const { return edge_off + (size_t)ctx->dist.world * 16; }
};

static int read_ctl(bpe_ctx* ctx, TrainRun& R, bool with_ties) {
    size_t bytes = with_ties ? sizeof(StepCtl) : offsetof(StepCtl, tie_keys);
    CU(cudaMemcpyAsync(R.h_ctl_p, R.ctl.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return BPE_OK;
}

// halo (+ run chain) for a host-chosen A (replay / verify passes)
static int launch_halo(bpe_ctx* ctx, Sequence<uint16_t>& sq, uint32_t A, bool aeqb, const EdgeInfo* edges) {
    const uint32_t nt = sq.ntiles();
    BPE_LAUNCH((halo_kernel<uint16_t, false>), (nt + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, ctx->stream, sq.tok(),
               sq.n_slots, nt, sq.halo.as<TileHalo<uint16_t>>(), (const StepCtl*)nullptr, A, aeqb ? 1 : 0,
               sq.run_local.as<uint32_t>(), sq.run_full.as<uint8_t>(), sq.done_counter.as<uint32_t>(), (uint32_t*)nullptr,
               edges, ctx->dist.rank, ctx->dist.world);
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// full table replay for one tie step (see tiebreak_host.hpp)
static int replay_winner(bpe_ctx* ctx, TrainRun& R, uint32_t max_count, uint32_t* winner) {
    int rc = launch_halo(ctx, R.sq, 0, false, R.edges());
    if (rc) return rc;
    const uint32_t cap = R.tm.cap;
    if (R.firstpos.bytes < (size_t)cap * 4) CU(R.firstpos.alloc((size_t)cap * 4));
    CU(cudaMemsetAsync(R.firstpos.p, 0xFF, (size_t)cap * 4, ctx->stream));
    CU(cudaMemsetAsync(&R.d_ctl()->last_pair_pos, 0, 4, ctx->stream));
    FirstPosOp op; op.tbl = R.tm.view(); op.firstpos = R.firstpos.as<uint32_t>(); op.ctl = R.d_ctl();
    BPE_LAUNCH((pair_visit_kernel<uint16_t, FirstPosOp>), R.sq.ntiles(), THREADS, ctx->stream, R.sq.tok(),
               R.sq.halo.as<TileHalo<uint16_t>>(), op);
    ctx->launches++;
    CU(cudaGetLastError());
    std::vector<uint32_t> keys(cap), counts(cap), fpos(cap);
    CU(cudaMemcpyAsync(keys.data(), R.tm.keys.p, (size_t)cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(counts.data(), R.tm.counts.p, (size_t)cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(fpos.data(), R.firstpos.p, (size_t)cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = read_ctl(ctx, R, false);
    if (rc) return rc;
    if (R.hc()->err) return fail(ctx, BPE_ERR_INTERNAL, "device error flags 0x%x during replay", R.hc()->err);
    std::vector<ReplayKey> rk;
    std::vector<uint32_t> tied;
    const uint64_t rank_hi = (uint64_t)ctx->dist.rank << 32;
    const bool multi = ctx->dist.world > 1;
    for (uint32_t i = 0; i < cap; i++) {
        if (keys[i] == EMPTY_KEY || counts[i] == 0) continue;
        if (!multi && fpos[i] == 0xFFFFFFFFu) return fail(ctx, BPE_ERR_INTERNAL, "live pair without occurrence");
        ReplayKey k; k.key = keys[i]; k.home_hash = (uint32_t)zig_hash_pair(keys[i]);
        k.first_pos = fpos[i] == 0xFFFFFFFFu ? ~0ull : (rank_hi | fpos[i]);  // (shard, slot) orders positions globally
        rk.push_back(k);
        if (counts[i] == max_count) tied.push_back(keys[i]);
    }
    uint64_t last_pair = rank_hi | R.hc()->last_pair_pos;
    if (multi) {
        // The replicated tables hold the same keys in different slots, so positions are exchanged in
        // key order: min over ranks of each key's first occurrence, max of the last pair position.
        std::sort(rk.begin(), rk.end(), [](const ReplayKey& a, const ReplayKey& b) { return a.key < b.key; });
        std::vector<uint64_t> pos(rk.size() + 1);
        for (size_t i = 0; i < rk.size(); i++) pos[i] = rk[i].first_pos;
        bool any_pair = false;
        for (size_t i = 0; i < rk.size(); i++) if (rk[i].first_pos != ~0ull) { any_pair = true; break; }
        DevBuf dpos;
        CU(dpos.alloc(pos.size() * 8));
        CU(cudaMemcpyAsync(dpos.p, pos.data(), rk.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->dist.allreduce(dpos.p, rk.size(), DIST_U64_MIN)) return fail(ctx, BPE_ERR_CUDA, "all-reduce (min) failed");
        uint64_t lp = any_pair ? last_pair : 0;
        CU(cudaMemcpyAsync(dpos.as<uint64_t>() + rk.size(), &lp, 8, cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->dist.allreduce(dpos.as<uint64_t>() + rk.size(), 1, DIST_U64_MAX)) return fail(ctx, BPE_ERR_CUDA, "all-reduce (max) failed");
        CU(cudaMemcpyAsync(pos.data(), dpos.p, pos.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < rk.size(); i++) {
            if (pos[i] == ~0ull) return fail(ctx, BPE_ERR_INTERNAL, "live pair without occurrence on any rank");
            rk[i].first_pos = pos[i];
        }
        last_pair = pos[rk.size()];
    }
    ZigTableReplay rep;
    rep.run(rk, last_pair);
    uint32_t w = rep.winner(rk, tied);
    if (w == ZigTableReplay::NONE) return fail(ctx, BPE_ERR_INTERNAL, "replay found no tied key");
    *winner = w;
    return BPE_OK;
}

static int verify_state(bpe_ctx* ctx, TrainRun& R, uint32_t step) {
    int rc = launch_halo(ctx, R.sq, 0, false, R.edges());
    if (rc) return rc;
    const uint32_t cap = R.tm.cap;
    if (R.recount.bytes < (size_t)cap * 4) CU(R.recount.alloc((size_t)cap * 4));
    if (!R.live_chk.p) CU(R.live_chk.alloc(4));
    CU(cudaMemsetAsync(R.recount.p, 0, (size_t)cap * 4, ctx->stream));
    CU(cudaMemsetAsync(R.live_chk.p, 0, 4, ctx->stream));
    RecountOp op; op.tbl = R.tm.view(); op.recount = R.recount.as<uint32_t>(); op.ctl = R.d_ctl();
    BPE_LAUNCH((pair_visit_kernel<uint16_t, RecountOp>), R.sq.ntiles(), THREADS, ctx->stream, R.sq.tok(),
               R.sq.halo.as<TileHalo<uint16_t>>(), op);
    BPE_LAUNCH_NS(verify_counts_kernel, grid_for(cap, 256), 256, ctx->stream, R.tm.view(), R.recount.as<uint32_t>(),
                  R.d_ctl(), R.live_chk.as<uint32_t>());
    ctx->launches += 2;
    CU(cudaGetLastError());
    uint32_t live = 0;
    CU(cudaMemcpyAsync(&live, R.live_chk.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = read_ctl(ctx, R, false);
    if (rc) return rc;
    if (R.hc()->err || R.hc()->verify_mismatch || live != R.hc()->live_keys)
        return fail(ctx, BPE_ERR_INTERNAL, "verify failed after step %u: err=0x%x mismatches=%u live=%u table_live=%u", step,
                    R.hc()->err, R.hc()->verify_mismatch, live, R.hc()->live_keys);
    return BPE_OK;
}

static int collect_heavy(bpe_ctx* ctx, TrainRun& R) {
    CU(cudaMemsetAsync(&R.d_ctl()->n_heavy, 0, 4, ctx->stream));
    BPE_LAUNCH_NS(heavy_collect_kernel, grid_for(R.tm.cap, 256), 256, ctx->stream, R.tm.view(), R.hl(), R.d_ctl());
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// rebuild the pair table (dropping dead keys) so that at least `need_free` more keys fit below 50 % load
static int grow_table(bpe_ctx* ctx, TrainRun& R, uint64_t need_free) {
    TableMem nt;
    uint32_t live = R.hc()->live_keys;
    uint32_t cap = R.tm.cap;
    while (((uint64_t)live + need_free) * 2 > cap || (uint64_t)live * 4 > cap) cap <<= 1;
    int rc = table_alloc(ctx, nt, cap);
    if (rc) return rc;
    CU(cudaMemsetAsync(&R.d_ctl()->n_inserted, 0, 4, ctx->stream));
    BPE_LAUNCH_NS(table_rehash_kernel, grid_for(R.tm.cap, 256), 256, ctx->stream, R.tm.view(), nt.view(), R.d_ctl(), 1);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    auto swap_buf = [](DevBuf& a, DevBuf& b) {
        std::swap(a.p, b.p); std::swap(a.bytes, b.bytes); std::swap(a.real, b.real); std::swap(a.from, b.from);
    };
    swap_buf(R.tm.keys, nt.keys); swap_buf(R.tm.counts, nt.counts); swap_buf(R.tm.zcnt, nt.zcnt);
    R.tm.cap = cap;
    R.tm.zcap = 0;
    R.hc()->n_inserted = live;
    return BPE_OK;
}

// make zcnt model the reference table for d live keys and tell the device
static int sync_zcap(bpe_ctx* ctx, TrainRun& R, uint32_t d) {
    int rc = table_ensure_zcnt(ctx, R.tm, R.d_ctl(), d);
    if (rc) return rc;
    BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 0, 0u, 1, R.tm.zcap, 0, 0, 0u, 0u);
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// one merge pass over the sequence: the register-streaming kernel (default) or the TMA-ring kernel (merge_impl = 1)
template <class TokT, bool DELTAS, bool FROMCTL>
static int launch_merge(bpe_ctx* ctx, TokT* tok, const TileHalo<TokT>* halo, uint32_t nt, const StepCtl* d_ctl, uint32_t* cntL,
                        uint32_t* cntR, uint32_t* nxx, uint32_t* nab, uint32_t A, uint32_t B, uint32_t X, uint32_t bins_min,
                        int backwards) {
    if (ctx->merge_impl == 1) {
        auto kern = merge_tma_kernel<TokT, DELTAS, FROMCTL>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem_bytes<TokT>()));
        const unsigned grid = (unsigned)std::min<uint32_t>(nt, (uint32_t)RING_CTAS_PER_SM * (uint32_t)ctx->num_sms);
        BPE_LAUNCH_SMEM(kern, grid, THREADS, ring_smem_bytes<TokT>(), ctx->stream, tok, halo, d_ctl, cntL, cntR, nxx, nab, A, B, X,
                        bins_min, nt);
    } else {
        const int pdl = (FROMCTL && ctx->pdl) ? 1 : 0;  // train loop: programmatic dependent of the previous apply kernel
        // steps whose pair occurs fewer than merge_direct times per tile on average (over all shards) skip the candidate queue
        const uint64_t dm = FROMCTL ? (uint64_t)std::max<int64_t>(ctx->merge_direct, 0) * nt * (uint64_t)std::max(ctx->dist.world, 1) : 0u;
        const uint32_t direct_max = (uint32_t)std::min<uint64_t>(dm, 0xFFFFFFFFull);
        if (ctx->merge_pairfilter == 0) backwards |= 2;
        if constexpr (FROMCTL && DELTAS && std::is_same<TokT, uint16_t>::value) {
            if (ctx->merge_loop > 0 && !ctx->merge_filter) {  // every CTA takes merge_loop tiles (strided by the grid size)
                const unsigned grid = (unsigned)std::max<uint64_t>(1, ((uint64_t)nt + (uint64_t)ctx->merge_loop - 1) / (uint64_t)ctx->merge_loop);
                if (ctx->merge_dbuf)
                    BPE_LAUNCH_PDL(merge_loop_kernel<true>, grid, THREADS, ctx->stream, pdl, tok, halo, d_ctl, cntL, cntR, nxx, nab, bins_min, backwards, direct_max, nt);
                else
                    BPE_LAUNCH_PDL(merge_loop_kernel<false>, grid, THREADS, ctx->stream, pdl, tok, halo, d_ctl, cntL, cntR, nxx, nab, bins_min, backwards, direct_max, nt);
                ctx->launches++;
                return BPE_OK;
            }
        }
        if (FROMCTL && ctx->merge_prestage && !ctx->merge_filter)
            BPE_LAUNCH_PDL((merge_kernel<TokT, DELTAS, FROMCTL, false, true>), nt, THREADS, ctx->stream, pdl, tok, halo, d_ctl, cntL, cntR, nxx, nab, A, B, X, bins_min,
                           backwards, direct_max);
        else if (ctx->merge_filter)
            BPE_LAUNCH_PDL((merge_kernel<TokT, DELTAS, FROMCTL, true>), nt, THREADS, ctx->stream, pdl, tok, halo, d_ctl, cntL, cntR, nxx, nab, A, B, X, bins_min,
                           backwards, direct_max);
        else
            BPE_LAUNCH_PDL((merge_kernel<TokT, DELTAS, FROMCTL, false>), nt, THREADS, ctx->stream, pdl, tok, halo, d_ctl, cntL, cntR, nxx, nab, A, B, X, bins_min,
                           backwards, direct_max);
    }
    ctx->launches++;
    return BPE_OK;
}

// candidate-scan variant of the step tail (single GPU, A != B): scan, resolve, write, apply
static int enqueue_step_tail_scan(bpe_ctx* ctx, TrainRun& R, uint32_t n_ids, uint32_t step_index, bool fuse_select) {
    const uint32_t nt = R.sq.ntiles();
    R.prof.sample_now = (step_index % 32u) == 0;
    if (R.prof.level == 3 && R.prof.sample_now) R.pending_samples.push_back(step_index);
    R.prof.mark(K_MERGE);
    BPE_LAUNCH_NS(scan_kernel<uint16_t>, nt, THREADS, ctx->stream, R.sq.tok(), R.d_ctl(), R.cand.as<uint32_t>(), R.cand_cap,
                  (int)(step_index & 1u));
    R.prof.mark(K_TIE);    // profile 1: the resolve kernel is booked under "tie", the write kernel under bucket 11
    BPE_LAUNCH_NS(resolve_kernel<uint16_t>, 148 * 4, 256, ctx->stream, R.sq.tok(), R.sq.n_slots, R.d_ctl(), R.cand.as<uint32_t>(), R.cand_cap,
                  R.wr.as<TokWrite>(), R.cntL(), R.cntR(), R.nxx(), R.nab());
    R.prof.mark(11);
    BPE_LAUNCH_NS(write_kernel<uint16_t>, 148 * 4, 256, ctx->stream, R.sq.tok(), R.d_ctl(), R.wr.as<TokWrite>());
    R.prof.mark(K_APPLY);
    PeerSet none;
    memset(&none, 0, sizeof none);
    HaloArgs no_halo;
    memset(&no_halo, 0, sizeof no_halo);
    BPE_LAUNCH(apply_kernel, (4 * n_ids + 3 + 255) / 256, 256, ctx->stream, R.tm.view(), R.delta.as<uint32_t>(), R.vcap, R.d_ctl(),
               R.tm.zig(), n_ids, R.hl(), R.d_rec(), fuse_select ? 1 : 0, 0, none, 0, 1, 0u, 0u, (uint32_t)R.edge_off, no_halo, PushArgs{0, 0, nullptr});
    ctx->launches += 4;
    CU(cudaGetLastError());
    return BPE_OK;
}

// The part of a step after the merge has been chosen: merge pass, exchange (multi-GPU), apply. All read the merge from
// ctl. The halos of the pass were gathered by the previous step's apply kernel (fused halo CTAs) or by the host after
// a compaction; classic_halo = true runs the stand-alone halo pass first — the only one that chains runs of equal
// tokens across tiles, needed when the merge has first == second (the loop halts with H_AEQB for those steps).
static int enqueue_step_tail(bpe_ctx* ctx, TrainRun& R, uint32_t n_ids, uint32_t step_index, bool fuse_select, bool classic_halo = false) {
    const uint32_t nt = R.sq.ntiles();
    R.prof.sample_now = (step_index % 32u) == 0;
    if (R.prof.level == 3 && R.prof.sample_now) R.pending_samples.push_back(step_index);  // credited once the step is known to have run
    if (classic_halo) {
        R.prof.mark(K_HALO);
        BPE_LAUNCH((halo_kernel<uint16_t, true>), (nt + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, ctx->stream, R.sq.tok(),
                   R.sq.n_slots, nt, R.sq.halo.as<TileHalo<uint16_t>>(), (const StepCtl*)R.d_ctl(), 0u, 0, R.sq.run_local.as<uint32_t>(),
                   R.sq.run_full.as<uint8_t>(), R.sq.done_counter.as<uint32_t>(), (uint32_t*)nullptr, R.edges(), ctx->dist.rank, ctx->dist.world);
        ctx->launches++;
    }
    const bool peer = ctx->dist.world > 1 && ctx->dist.peer_ok && ctx->xchg_impl == 0 && ctx->merge_impl != 1 &&
                      R.exchange_words() <= ctx->dist.peers.slot_words;
    const uint32_t parity = step_index & 1u, epoch = ctx->dist.epoch_base + step_index + 1u;
    R.prof.mark(K_MERGE);
    {
        int rcm = launch_merge<uint16_t, true, true>(ctx, R.sq.tok(), R.sq.halo.as<TileHalo<uint16_t>>(), nt, (const StepCtl*)R.d_ctl(),
                                                     R.cntL(), R.cntR(), R.nxx(), R.nab(), 0u, 0u, 0u, 4u * nt, (int)(step_index & 1u));
        if (rcm) return rcm;
    }
    R.prof.mark(K_APPLY);
    if (!peer && ctx->dist.world > 1) {
        // describe this shard's (post-merge) ends, then one all-reduce sums the deltas and gathers the edges
        BPE_LAUNCH(edge_kernel<uint16_t>, 1, 32, ctx->stream, R.sq.tok(), R.sq.n_slots, R.sq.dense_end, R.edges(), ctx->dist.rank, ctx->dist.world,
                      R.d_ctl(), R.nab(), 1);
        ctx->launches++;
        if (!ctx->dist.allreduce(R.delta.p, R.exchange_words(), DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the merge deltas failed");
    }
    // peer path: every apply CTA first stores its cells of this rank's deltas into every peer's mailbox over NVLink
    PushArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.n_slots = R.sq.n_slots; pa.tail_hint = R.sq.dense_end; pa.tok = R.sq.tok();
    HaloArgs ha;
    ha.tok = R.sq.tok(); ha.n_slots = R.sq.n_slots; ha.ntiles = nt; ha.halo = R.sq.halo.as<TileHalo<uint16_t>>();
    ha.apply_blocks = (4 * n_ids + 3 + 255) / 256;
    ha.step1 = step_index + 1u;
    if (ctx->fuse_halo == 0 || ctx->merge_impl != 0) ha.tok = nullptr;
    const uint32_t grid = ha.apply_blocks + (ha.tok ? (nt + 255) / 256 : 0u);
    // (a programmatic dependent of the merge pass when nothing else was enqueued in between)
    BPE_LAUNCH_PDL(apply_kernel, grid, 256, ctx->stream, (ctx->pdl && (peer || ctx->dist.world == 1)) ? 1 : 0, R.tm.view(), R.delta.as<uint32_t>(), R.vcap, R.d_ctl(),
                   R.tm.zig(), n_ids, R.hl(), R.d_rec(), fuse_select ? 1 : 0, peer ? 1 : 0, ctx->dist.peers, ctx->dist.rank, ctx->dist.world,
                   parity, epoch, (uint32_t)R.edge_off, ha, pa);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return BPE_OK;
}

static int train_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, uint16_t vocab_size, bpe_merge_t* out_merges,
                        uint64_t* out_counts, size_t* out_n, bpe_stats_t* stats_out) {
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (vocab_size < 256) return fail(ctx, BPE_ERR_INVALID_VOCAB, "vocab_size %u < 256", (unsigned)vocab_size);
    size_t want = (size_t)vocab_size - 256;
    if (ctx->max_steps > 0 && (size_t)ctx->max_steps < want) want = (size_t)ctx->max_steps;
    if (want && !out_merges) return fail(ctx, BPE_ERR_INVALID_ARG, "out_merges is null");
    if (n >= 0xFFFFFFF0ull) return fail(ctx, BPE_ERR_INVALID_ARG, "shard of %zu bytes exceeds the 32-bit position range", n);
    const double t_host0 = now_ms();
    const uint64_t launches0 = ctx->launches;
    TrainRun R;
    R.ctx = ctx;
    memset(&R.st, 0, sizeof R.st);
    const bool multi = ctx->dist.world > 1;
    if (want == 0 || (!multi && n < 2)) {  // no pairs (:188-191; n == 0 underflows in the reference, defined as no pairs)
        if (stats_out) { *stats_out = R.st; stats_out->total_ms = now_ms() - t_host0; }
        return BPE_OK;
    }
    CU(cudaSetDevice(ctx->device));
    cudaEvent_t ev0, ev1;
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    CU(cudaEventRecord(ev0, ctx->stream));
    const bool debug_sync = ctx->verify_recount != 0 && !multi;  // per-step host checks: batches of one step
    EvProfile& prof = R.prof;
    prof.init(ctx->time_phases ? 1 : (int)ctx->profile, ctx->stream, &ctx->ev_pool, R.st.kernel_ms, R.st.kernel_calls);
    prof.mark(K_INIT);

    int rc = seq_init(ctx, R.sq, d_text, n);
    if (rc) return rc;
    R.vcap = ((uint32_t)vocab_size + 2u) & ~1u;  // even, so the cntL|cntR block is a whole number of 16-byte vectors
    R.edge_off = ((size_t)2 * R.vcap + 2 + 15) / 16 * 16;
    CU(R.delta.alloc(R.exchange_words() * 4)); CU(R.hist.alloc(65536 * 8));
    CU(R.ctl.alloc(sizeof(StepCtl))); CU(R.rec.alloc(want * sizeof(MergeRec)));
    const bool scan_path = ctx->merge_impl == 2 && !multi;
    if (scan_path) {
        R.cand_cap = (uint32_t)std::max<size_t>(1u << 16, n / 32);
        CU(R.cand.alloc((size_t)R.cand_cap * 4));
        CU(R.wr.alloc((size_t)R.cand_cap * 2 * sizeof(TokWrite)));
    }
    uint32_t scan_not_before = 0;  // first step that may use the candidate-scan path again (set after a dense step)
    // fused halo: the apply kernel of step t gathers the tile halos of step t + 1; the host gathers them at the start and
    // after every compaction, and first == second steps halt the loop for the run-chaining halo pass (H_AEQB)
    const bool fused_halo = ctx->fuse_halo != 0 && ctx->merge_impl == 0;
    bool halo_stale = true;
    if (!ctx->h_ctl) CU(cudaMallocHost(&ctx->h_ctl, sizeof(StepCtl)));
    R.h_ctl_p = ctx->h_ctl;
    CU(cudaMemsetAsync(R.delta.p, 0, R.exchange_words() * 4, ctx->stream));
    CU(cudaMemsetAsync(R.hist.p, 0, 65536 * 8, ctx->stream));
    CU(cudaMemsetAsync(R.ctl.p, 0, sizeof(StepCtl), ctx->stream));

    // initial count (countCodePointPairs :257-278 on the byte sequence)
    {
        CU(cudaFuncSetAttribute(byte_pair_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HIST_SMEM));
        if (multi) {  // the pair that straddles two shards belongs to the left one: it needs the next shard's first byte
            BPE_LAUNCH(edge_kernel<uint16_t>, 1, 32, ctx->stream, R.sq.tok(), R.sq.n_slots, R.sq.dense_end, R.edges(), ctx->dist.rank, ctx->dist.world,
                          R.d_ctl(), R.nab(), 0);
            ctx->launches++;
            if (!ctx->dist.allreduce(R.edges(), (size_t)ctx->dist.world * 16, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the shard edges failed");
        }
        const unsigned hgrid = (unsigned)std::max<size_t>(1, std::min<size_t>(148, (n + HIST_PASS - 1) / HIST_PASS));
        BPE_LAUNCH_SMEM(byte_pair_hist_kernel, hgrid, HIST_THREADS, HIST_SMEM, ctx->stream, d_text, n, (const EdgeInfo*)R.edges(),
                        ctx->dist.rank, ctx->dist.world, R.hist.as<unsigned long long>());
        if (multi && !ctx->dist.allreduce(R.hist.p, 65536, DIST_U64_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the byte-pair histogram failed");
    }
    BPE_LAUNCH_NS(hist_nonzero_kernel, 65536 / 256, 256, ctx->stream, R.hist.as<unsigned long long>(), R.d_ctl());
    ctx->launches += 2;
    CU(cudaGetLastError());
    rc = read_ctl(ctx, R, false);
    if (rc) return rc;
    uint32_t cap = 1u << 19;
    if (ctx->table_log2 > 0) cap = 1u << ctx->table_log2;
    else while ((size_t)cap < n / 64 && cap < (1u << 28)) cap <<= 1;
    while ((uint64_t)R.hc()->hist_nonzero * 4 > cap) cap <<= 1;  // the byte pairs alone must fit with room to spare
    rc = table_alloc(ctx, R.tm, cap);
    if (rc) return rc;
    // the heavy list can hold every key of the table, so it cannot overflow even when the whole table ties
    // (e.g. the late phase of a long training on random bytes, where every pair occurs once)
    CU(R.heavy.alloc((size_t)cap * 4));
    // pair counts are 32 bits wide in the table: refuse a corpus whose most frequent pair does not fit ("count_limit_log2"
    // lowers the limit for tests)
    const unsigned long long count_limit = ctx->count_limit_log2 >= 32 ? 0xFFFFFFFFull : ((1ull << ctx->count_limit_log2) - 1ull);
    BPE_LAUNCH_NS(seed_table_kernel, 65536 / 256, 256, ctx->stream, R.hist.as<unsigned long long>(), R.tm.view(), R.d_ctl(), count_limit);
    ctx->launches += 1;
    CU(cudaGetLastError());
    {   // device-side loop state
        StepCtl init;
        memset(&init, 0, sizeof init);
        init.want_steps = (uint32_t)want;
        init.live_tokens = n;  // multi-GPU: only local_live is meaningful
        init.local_live = n;
        init.flags = (ctx->force_slow_tiebreak ? F_FORCE_REPLAY : 0u) | (ctx->check_tiebreak ? F_CHECK_TIES : 0u) |
                     (fused_halo ? F_HALT_AEQB : 0u);
        CU(cudaMemcpyAsync(&R.d_ctl()->step, &init.step, offsetof(StepCtl, tie_keys) - offsetof(StepCtl, step),
                           cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));  // `init` lives on this stack frame
    }
    rc = read_ctl(ctx, R, false);
    if (rc) return rc;
    if (R.hc()->err & ERR_COUNT_OVERFLOW)
        return fail(ctx, BPE_ERR_INTERNAL, "a byte pair occurs more than %llu times: beyond the 32-bit pair counts of this engine (the reference counts in usize)", count_limit);
    rc = sync_zcap(ctx, R, R.hc()->live_keys);
    if (rc) return rc;

    uint32_t steps_done = 0;
    bool finished = false;
    bool selection_pending = false;  // the last enqueued apply_kernel also selects the following merge
    while (!finished) {
        StepCtl* hc = R.hc();
        // ---- housekeeping between batches (the table and the sequence are quiescent here) ----
        // steps enqueued between two read-backs of the control block: after a halt the rest of a batch are no-op launches
        // (~50 us each at 1 GB), so shorter batches waste less; every batch costs one host round trip
        const uint32_t max_batch = debug_sync ? 1u : (uint32_t)std::max<long>(1, std::min<long>(ctx->batch_steps, 256));
        uint32_t K = (uint32_t)std::min<size_t>(max_batch, want - steps_done);
        const uint64_t per_step_inserts = 2ull * (256 + steps_done + K + 1) + 1;
        prof.mark(K_TABLE);
        if (((uint64_t)hc->n_inserted + (uint64_t)(K + 2) * per_step_inserts) * 4 > (uint64_t)R.tm.cap * 3) {
            rc = grow_table(ctx, R, (uint64_t)(K + 2) * per_step_inserts);
            if (rc) return rc;
            rc = sync_zcap(ctx, R, hc->live_keys);
            if (rc) return rc;
            if (R.heavy.bytes < (size_t)R.tm.cap * 4) CU(R.heavy.alloc((size_t)R.tm.cap * 4));
            if (R.theta) { rc = collect_heavy(ctx, R); if (rc) return rc; }  // slot indices changed
        }
        if (R.sq.n_slots > (size_t)TILE && R.sq.live * 100 < (uint64_t)R.sq.n_slots * (uint64_t)ctx->compact_pct) {
            prof.mark(K_COMPACT);
            rc = seq_compact(ctx, R.sq, nullptr);
            if (rc) return rc;
            R.st.compactions++;
            halo_stale = true;
        }
        if (fused_halo && halo_stale) {
            prof.mark(K_HALO);
            rc = launch_halo(ctx, R.sq, 0, false, R.edges());
            if (rc) return rc;
            halo_stale = false;
        }
        // ---- one batch of device-driven steps ----
        for (uint32_t k = 0; k < K; k++) {
            if (!selection_pending) {  // otherwise the previous apply_kernel already chose this step's merge
                prof.mark(K_ARGMAX);
                BPE_LAUNCH(select_kernel, 1, 1024, ctx->stream, R.tm.view(), R.hl(), R.d_ctl(), R.d_rec(), R.tm.zig());
                ctx->launches += 1;
            }
            selection_pending = !debug_sync;
            const bool use_scan = scan_path && steps_done >= scan_not_before && (uint64_t)hc->max_count * 4 < R.cand_cap;
            rc = use_scan ? enqueue_step_tail_scan(ctx, R, 256 + steps_done + k + 1, steps_done + k, selection_pending)
                          : enqueue_step_tail(ctx, R, 256 + steps_done + k + 1, steps_done + k, selection_pending, !fused_halo);
            if (rc) return rc;
        }
        prof.mark(K_HOSTGAP);
        rc = read_ctl(ctx, R, true);
        if (rc) return rc;
        hc = R.hc();
        if (hc->err) return fail(ctx, BPE_ERR_INTERNAL, "device error flags 0x%x near step %u", hc->err, hc->step);
        R.st.scanned_slots += (uint64_t)(hc->step - steps_done) * R.sq.n_slots;
        for (uint32_t sidx : R.pending_samples) {
            if (sidx < hc->step) R.st.kernel_ms[10] += (double)R.sq.n_slots; else R.sampled_noop++;
        }
        R.pending_samples.clear();
        steps_done = hc->step;
        R.sq.live = multi ? hc->local_live : hc->live_tokens;
        if (debug_sync && hc->halt == H_NONE) { rc = verify_state(ctx, R, steps_done); if (rc) return rc; hc = R.hc(); }
        if (ctx->debug && (hc->halt != H_NONE || ctx->debug > 1))
            fprintf(stderr, "[bpe r%d] step %u halt %u (%u,%u)->%u max %u ntied %u live_keys %u inserted %u cap %u n_heavy %u theta %u zcap %u slots %zu live %llu err 0x%x\n",
                    ctx->dist.rank, hc->step, hc->halt, hc->A, hc->B, hc->X, hc->max_count, hc->ntied, hc->live_keys, hc->n_inserted, R.tm.cap, hc->n_heavy,
                    R.theta, R.tm.zcap, R.sq.n_slots, (unsigned long long)R.sq.live, hc->err);
        if (hc->halt != H_NONE) selection_pending = false;  // the host intervenes: start the next batch with a select_kernel
        switch (hc->halt) {
            case H_NONE: break;
            case H_DONE: finished = true; break;
            case H_ZCAP:
                prof.mark(K_TABLE);
                rc = sync_zcap(ctx, R, hc->live_keys);
                if (rc) return rc;
                BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 0, 0u, 0, 0u, 1, 0, 0u, 0u);
                ctx->launches++;
                break;
            case H_HEAVY: {
                // the list no longer covers the maximum: one full pass, then rebuild it
                prof.mark(K_ARGMAX);
                BPE_LAUNCH_NS(reset_argmax_kernel, 1, 1, ctx->stream, R.d_ctl());
                BPE_LAUNCH(argmax_kernel, grid_for(R.tm.cap, THREADS), THREADS, ctx->stream, R.tm.view(), R.d_ctl());
                ctx->launches += 2;
                CU(cudaGetLastError());
                rc = read_ctl(ctx, R, false);
                if (rc) return rc;
                hc = R.hc();
                if (hc->max_count == 0) { finished = true; break; }  // "No more pairs to merge" (:188-191)
                // theta = 3/4 of the maximum: a short list (the per-step select scans it) at the price of a few more rebuilds
                R.theta = hc->max_count - hc->max_count / 4;
                if (R.theta < 1) R.theta = 1;
                rc = collect_heavy(ctx, R);
                if (rc) return rc;
                BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 1, R.theta, 0, 0u, 1, 0, 0u, 0u);
                ctx->launches++;
                break;
            }
            case H_CLASSIC: {
                // the candidate-scan path declined this step (A == B or too many candidates); nothing was modified:
                // run the tiled pass for the merge already chosen
                if (hc->A != hc->B) scan_not_before = steps_done + 64;
                BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 0, 0u, 0, 0u, 1, 0, 0u, 0u);
                ctx->launches++;
                rc = enqueue_step_tail(ctx, R, 256 + steps_done + 1, steps_done, false, true);
                if (rc) return rc;
                rc = read_ctl(ctx, R, false);
                if (rc) return rc;
                hc = R.hc();
                if (hc->err) return fail(ctx, BPE_ERR_INTERNAL, "device error flags 0x%x near step %u", hc->err, hc->step);
                R.st.scanned_slots += R.sq.n_slots;
                for (uint32_t sidx : R.pending_samples) {
                    if (sidx < hc->step) R.st.kernel_ms[10] += (double)R.sq.n_slots; else R.sampled_noop++;
                }
                R.pending_samples.clear();
                steps_done = hc->step;
                R.sq.live = hc->live_tokens;
                if (debug_sync) { rc = verify_state(ctx, R, steps_done); if (rc) return rc; }
                break;
            }
            case H_AEQB: {
                // the merge is chosen and recorded; its pass needs the run-chaining halo. Nothing is read back: the
                // step is known to run, and its apply kernel already selects the following merge.
                BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 0, 0u, 0, 0u, 1, 0, 0u, 0u);
                ctx->launches++;
                rc = enqueue_step_tail(ctx, R, 256 + steps_done + 1, steps_done, !debug_sync, true);
                if (rc) return rc;
                R.aeqb_steps++;
                if (debug_sync) {
                    rc = read_ctl(ctx, R, false);
                    if (rc) return rc;
                    hc = R.hc();
                    if (hc->err) return fail(ctx, BPE_ERR_INTERNAL, "device error flags 0x%x near step %u", hc->err, hc->step);
                    R.st.scanned_slots += R.sq.n_slots;
                    steps_done = hc->step;
                    R.sq.live = multi ? hc->local_live : hc->live_tokens;
                    rc = verify_state(ctx, R, steps_done);
                    if (rc) return rc;
                } else {
                    R.st.scanned_slots += R.sq.n_slots;
                    steps_done += 1;
                    selection_pending = true;
                }
                break;
            }
            case H_REPLAY: {
                R.st.tie_steps++;
                prof.mark(K_REPLAY);
                const uint32_t maxc = hc->max_count;
                const bool had_fast = hc->tie_status == TIE_FAST_OK;
                const uint32_t fast_winner = hc->tie_winner;
                uint32_t w = 0;
                rc = replay_winner(ctx, R, maxc, &w);
                if (rc) return rc;
                if (had_fast && w != fast_winner)
                    return fail(ctx, BPE_ERR_INTERNAL, "tie fast path chose (%u,%u), replay chose (%u,%u) at step %u",
                                fast_winner & 0xFFFF, fast_winner >> 16, w & 0xFFFF, w >> 16, steps_done);
                if (!had_fast) R.st.tie_slow_steps++;
                BPE_LAUNCH_NS(ctl_set_kernel, 1, 1, ctx->stream, R.d_ctl(), R.d_rec(), 0, 0u, 0, 0u, 1, 1, w, maxc);
                ctx->launches++;
                rc = enqueue_step_tail(ctx, R, 256 + steps_done + 1, steps_done, false, !fused_halo || (w & 0xFFFFu) == (w >> 16));
                if (rc) return rc;
                rc = read_ctl(ctx, R, false);
                if (rc) return rc;
                hc = R.hc();
                if (hc->err) return fail(ctx, BPE_ERR_INTERNAL, "device error flags 0x%x near step %u", hc->err, hc->step);
                R.st.scanned_slots += R.sq.n_slots;
                for (uint32_t sidx : R.pending_samples) {
                    if (sidx < hc->step) R.st.kernel_ms[10] += (double)R.sq.n_slots; else R.sampled_noop++;
                }
                R.pending_samples.clear();
                steps_done = hc->step;
                R.sq.live = multi ? hc->local_live : hc->live_tokens;
                if (debug_sync) { rc = verify_state(ctx, R, steps_done); if (rc) return rc; }
                break;
            }
            default: return fail(ctx, BPE_ERR_INTERNAL, "unknown halt code %u", hc->halt);
        }
        CU(cudaGetLastError());
        if (steps_done >= want) finished = true;
    }
    ctx->dist.epoch_base += (uint32_t)want + 8u;  // arrival flags of the peer exchange never go backwards
    // merge list back to the host
    std::vector<MergeRec> recs(steps_done);
    if (steps_done) CU(cudaMemcpyAsync(recs.data(), R.rec.p, steps_done * sizeof(MergeRec), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ev1, ctx->stream));
    CU(cudaEventSynchronize(ev1));
    prof.finish();
    if (prof.level == 3 && R.st.kernel_calls[K_MERGE] >= R.sampled_noop) R.st.kernel_calls[K_MERGE] -= R.sampled_noop;
    float dev_ms = 0;
    CU(cudaEventElapsedTime(&dev_ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    for (size_t i = 0; i < steps_done; i++) {
        out_merges[i].first = (uint16_t)(recs[i].key & 0xFFFFu);
        out_merges[i].second = (uint16_t)(recs[i].key >> 16);
        out_merges[i].new_token = (uint16_t)(256 + i);
        if (out_counts) out_counts[i] = recs[i].count;
    }
    *out_n = steps_done;
    R.st.tie_steps += R.hc()->fast_ties;
    R.st.aeqb_steps = R.aeqb_steps;
    R.st.device_ms = dev_ms;
    R.st.total_ms = now_ms() - t_host0;
    R.st.kernel_launches = ctx->launches - launches0;
    // reference TimeStats buckets from the kernel classes
    R.st.sort_pairs_ms = R.st.kernel_ms[K_ARGMAX] + R.st.kernel_ms[K_TIE] + R.st.kernel_ms[K_REPLAY];
    R.st.sort_pairs_calls = steps_done;
    R.st.replace_pair_ms = R.st.kernel_ms[K_HALO] + R.st.kernel_ms[K_MERGE] + R.st.kernel_ms[K_COMPACT];
    R.st.replace_pair_calls = steps_done;
    R.st.just_count_pairs_ms = R.st.kernel_ms[K_INIT] + R.st.kernel_ms[K_APPLY];
    R.st.just_count_pairs_calls = steps_done + 1;
    if (stats_out) *stats_out = R.st;
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// encode (src/basic_tokenizer.zig:71-88): the list in order, commuting merges grouped into level passes (build_encode_schedule)
// -----------------------------------------------------------------------------------------
// Encode schedule. The reference applies the merges one at a time in list order (:71-88). Two merges q < r
// commute on every token sequence when neither uses the token the other produces and no token is the second
// component of one pair and the first component of the other (their occurrences then never overlap and
// neither pass creates or destroys an occurrence of the other). Giving every merge the level
// 1 + max(level of the earlier merges it does NOT commute with) and running the levels in order is
// therefore a reordering of the list by swaps of adjacent commuting merges, i.e. the same function.
// All pairs of a level with first != second go through ONE level_kernel pass; pairs with first == second
// (greedy runs) keep their own merge pass inside their level. Lists a trained tokenizer cannot produce
// (new id below 256, equal to a component, already mentioned earlier, or a repeated pair) act as barriers
// and run alone, in list position.
struct EncStep { int single; uint32_t off, cnt; };  // single >= 0: merges[single] as one merge pass; else a level pass over ents[off, off+cnt)

// Level of every merge (1-based) and whether it has to run alone. Returns true when the whole list is
// *regular* (what a trained tokenizer writes: every new id >= 256, different from its components, not
// mentioned earlier, no repeated pair) — the precondition of the segment-resident encoder.
static bool merge_levels(const bpe_merge_t* merges, size_t m, std::vector<uint32_t>& lvl, std::vector<uint8_t>& alone) {
    std::vector<uint32_t> lastZ(65536, 0), lastF(65536, 0), lastS(65536, 0);
    std::vector<uint8_t> used(65536, 0);
    std::unordered_set<uint32_t> seen;
    seen.reserve(m * 2 + 16);
    lvl.assign(m, 0);
    alone.assign(m, 0);
    uint32_t floor_lvl = 0, max_lvl = 0;
    bool all_regular = true;
    for (size_t i = 0; i < m; i++) {
        const uint32_t a = merges[i].first, b = merges[i].second, z = merges[i].new_token, key = pair_key(a, b);
        const bool regular = z >= 256 && z != a && z != b && !used[z] && !seen.count(key);
        uint32_t l;
        if (regular) {
            l = 1 + std::max(std::max(floor_lvl, lastZ[a]), std::max(lastZ[b], std::max(lastS[a], lastF[b])));
            alone[i] = (a == b) ? 1 : 0;
        } else {
            l = max_lvl + 1;
            floor_lvl = l;
            alone[i] = 1;
            all_regular = false;
        }
        lvl[i] = l;
        lastF[a] = std::max(lastF[a], l); lastS[b] = std::max(lastS[b], l); lastZ[z] = std::max(lastZ[z], l);
        used[a] = used[b] = used[z] = 1;
        seen.insert(key);
        max_lvl = std::max(max_lvl, l);
    }
    return all_regular;
}

static void build_encode_schedule(const bpe_merge_t* merges, size_t m, bool levels, std::vector<EncStep>& steps,
                                  std::vector<LevelEntry>& ents) {
    steps.clear();
    ents.clear();
    if (!levels) {
        for (size_t i = 0; i < m; i++) steps.push_back(EncStep{(int)i, 0u, 0u});
        return;
    }
    std::vector<uint32_t> lvl;
    std::vector<uint8_t> alone;
    merge_levels(merges, m, lvl, alone);
    std::vector<uint32_t> order(m);
    for (size_t i = 0; i < m; i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return lvl[x] < lvl[y]; });
    size_t i = 0;
    while (i < m) {
        size_t j = i;
        while (j < m && lvl[order[j]] == lvl[order[i]]) j++;
        uint32_t off = (uint32_t)ents.size(), cnt = 0;
        for (size_t k = i; k < j; k++) {
            const bpe_merge_t& mg = merges[order[k]];
            if (alone[order[k]]) continue;
            ents.push_back(LevelEntry{pair_key(mg.first, mg.second), mg.new_token});
            if (++cnt == (uint32_t)LVL_MAX) { steps.push_back(EncStep{-1, off, cnt}); off = (uint32_t)ents.size(); cnt = 0; }
        }
        if (cnt) steps.push_back(EncStep{-1, off, cnt});
        for (size_t k = i; k < j; k++)
            if (alone[order[k]]) steps.push_back(EncStep{(int)order[k], 0u, 0u});
        i = j;
    }
}

// -----------------------------------------------------------------------------------------
// segment-resident encode (bpe_segenc.cuh): one launch over the text + the final squeeze
// -----------------------------------------------------------------------------------------
template <int C, int M, int NT>
static int launch_segenc(bpe_ctx* ctx, const uint8_t* d_text, long long n, const uint8_t* halo_l, int hl, const uint8_t* halo_r, int hr,
                         long long g_lo, long long k_first, long long k_last, const SegTab& T, uint16_t* slots, uint32_t* d_fail) {
    auto kern = segenc_kernel<C, M, NT>;
    const size_t smem = segenc_smem_bytes<C, M, NT>();
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long cores = k_last - k_first + 1;
    const long long tiles = (cores + (NT - 2) - 1) / (NT - 2);
    if (tiles > 0x7FFFFFFFll) return fail(ctx, BPE_ERR_INVALID_ARG, "input too large for one launch");
    BPE_LAUNCH_SMEM(kern, (unsigned)tiles, NT, smem, ctx->stream, d_text, n, halo_l, hl, halo_r, hr, g_lo, k_first, k_last, T, slots, d_fail);
    ctx->launches++;
    CU(cudaGetLastError());
    return BPE_OK;
}

// ---- pieces shared by the segment-resident and the tile-resident encoders ----
// pair -> (level, new id) tables on the device: byte pairs dense, the rest hashed; byte length of every id
struct SegTables {
    DevBuf tab;
    SegTab T;
    uint32_t max_level = 0;
};
static int build_seg_tables(bpe_ctx* ctx, const bpe_merge_t* merges, size_t m, const std::vector<uint32_t>& lvl, SegTables& st) {
    size_t nhash = 0;
    for (size_t i = 0; i < m; i++) nhash += (merges[i].first | merges[i].second) >= 256;
    uint32_t hlog = 6;
    while (((size_t)1 << hlog) < 2 * nhash + 2) hlog++;
    const size_t hcap = (size_t)1 << hlog;
    std::vector<uint32_t> host((size_t)65536 + 2 * hcap + 65536 / 2);
    uint32_t* bp = host.data();
    uint32_t* hk = bp + 65536;
    uint16_t* len = reinterpret_cast<uint16_t*>(hk + 2 * hcap);
    for (size_t i = 0; i < 65536; i++) bp[i] = SEG_NONE;
    for (size_t i = 0; i < hcap; i++) { hk[2 * i] = EMPTY_KEY; hk[2 * i + 1] = SEG_NONE; }
    for (size_t i = 0; i < 65536; i++) len[i] = 1;
    st.max_level = 0;
    for (size_t i = 0; i < m; i++) {
        const uint32_t a = merges[i].first, b = merges[i].second, z = merges[i].new_token;
        const uint32_t val = lvl[i] | (z << 16);
        st.max_level = std::max(st.max_level, lvl[i]);
        if ((a | b) < 256) bp[a | (b << 8)] = val;
        else {
            const uint32_t key = pair_key(a, b);
            uint32_t s = (key * 0x9E3779B1u) >> (32 - hlog);
            while (hk[2 * s] != EMPTY_KEY) s = (s + 1) & (uint32_t)(hcap - 1);
            hk[2 * s] = key; hk[2 * s + 1] = val;
        }
        len[z] = (uint16_t)std::min<uint32_t>(65535u, (uint32_t)len[a] + (uint32_t)len[b]);
    }
    CU(st.tab.alloc(host.size() * 4));
    CU(cudaMemcpyAsync(st.tab.p, host.data(), host.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // `host` lives on this frame
    st.T.bp = st.tab.as<uint32_t>();
    st.T.hk = reinterpret_cast<const uint2*>(st.tab.as<uint32_t>() + 65536);
    st.T.hmask = (uint32_t)(hcap - 1);
    st.T.hshift = 32 - hlog;
    st.T.len = reinterpret_cast<const uint16_t*>(st.tab.as<uint32_t>() + 65536 + 2 * hcap);
    return BPE_OK;
}

// multi-GPU: where this shard lies in the whole text, and up to SEG_HALO bytes of text on both sides of it
struct ShardHalo {
    DevBuf halo;  // [SEG_HALO left, right-aligned data moved to the front | SEG_HALO right]
    std::vector<uint8_t> host;  // the same bytes on the host (the streaming encoder builds per-chunk halos from them)
    int hl = 0, hr = 0;
    long long g_lo = 0, g_total = 0, min_shard = 0;
    const uint8_t* left() const { return halo.as<uint8_t>(); }
    const uint8_t* right() const { return halo.as<uint8_t>() + SEG_HALO; }
};
static int exchange_shard_halo(bpe_ctx* ctx, const uint8_t* d_text, size_t n, ShardHalo& sh, bool text_on_host = false) {
    const int world = ctx->dist.world, rank = ctx->dist.rank;
    sh.g_lo = 0; sh.g_total = (long long)n; sh.hl = sh.hr = 0; sh.min_shard = (long long)n;
    sh.host.assign(2 * SEG_HALO, 0);
    CU(sh.halo.alloc(2 * SEG_HALO));
    if (world == 1) return BPE_OK;
    // every rank contributes [n (2 words) | first SEG_HALO bytes | last SEG_HALO bytes]; the sum over ranks of
    // buffers that are zero outside the own slot is an all-gather
    const size_t W = 2 + 2 * SEG_HALO / 4;
    std::vector<uint32_t> all((size_t)world * W, 0u);
    std::vector<uint8_t> ends(2 * SEG_HALO, 0);
    const size_t nf = std::min<size_t>(n, SEG_HALO);
    if (nf && text_on_host) {
        memcpy(ends.data(), d_text, nf);
        memcpy(ends.data() + SEG_HALO, d_text + (n - nf), nf);
    } else if (nf) {
        CU(cudaMemcpyAsync(ends.data(), d_text, nf, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ends.data() + SEG_HALO, d_text + (n - nf), nf, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    uint32_t* mine = all.data() + (size_t)rank * W;
    mine[0] = (uint32_t)(n & 0xFFFFFFFFu); mine[1] = (uint32_t)((uint64_t)n >> 32);
    memcpy(mine + 2, ends.data(), 2 * SEG_HALO);
    DevBuf xb;
    CU(xb.alloc(all.size() * 4));
    CU(cudaMemcpyAsync(xb.p, all.data(), all.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->dist.allreduce(xb.p, all.size(), DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-gather of the shard ends failed");
    CU(cudaMemcpyAsync(all.data(), xb.p, all.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    auto size_of = [&](int r) { return (long long)(((uint64_t)all[(size_t)r * W + 1] << 32) | all[(size_t)r * W]); };
    sh.g_total = 0;
    for (int r = 0; r < world; r++) { if (r == rank) sh.g_lo = sh.g_total; sh.g_total += size_of(r); sh.min_shard = std::min(sh.min_shard, size_of(r)); }
    std::vector<uint8_t> hb(2 * SEG_HALO, 0);
    int hl = 0, hr = 0;
    // left halo: the last bytes of the ranks before me, nearest first, written right-aligned into hb[0, SEG_HALO)
    for (int r = rank - 1; r >= 0 && hl < SEG_HALO; r--) {
        const long long sz = size_of(r);
        const int have = (int)std::min<long long>(sz, SEG_HALO);        // bytes of rank r held in its "last" block
        const uint8_t* last = reinterpret_cast<const uint8_t*>(all.data() + (size_t)r * W + 2) + SEG_HALO;
        const int take = std::min(have, SEG_HALO - hl);
        memcpy(hb.data() + SEG_HALO - hl - take, last + (have - take), (size_t)take);
        hl += take;
    }
    if (hl < SEG_HALO) memmove(hb.data(), hb.data() + SEG_HALO - hl, (size_t)hl);  // left-align: kernels read halo_l[hl + p], p in [-hl, 0)
    for (int r = rank + 1; r < world && hr < SEG_HALO; r++) {
        const long long sz = size_of(r);
        const int have = (int)std::min<long long>(sz, SEG_HALO);
        const uint8_t* first = reinterpret_cast<const uint8_t*>(all.data() + (size_t)r * W + 2);
        const int take = std::min(have, SEG_HALO - hr);
        memcpy(hb.data() + SEG_HALO + hr, first, (size_t)take);
        hr += take;
    }
    sh.hl = hl; sh.hr = hr;
    sh.host = hb;
    CU(cudaMemcpyAsync(sh.halo.p, hb.data(), 2 * SEG_HALO, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // hb lives on this frame
    return BPE_OK;
}

// squeeze the holes out of a u16 slot array of n slots (allocated to a whole number of tiles), straight into d_out
static int squeeze_slots(bpe_ctx* ctx, DevBuf& slots, size_t n, uint16_t* d_out, size_t* out_n) {
    const size_t n_slots = round_up(n ? n : 1, TILE);
    if (n_slots > n) {
        BPE_LAUNCH_NS(fill_holes_kernel<uint16_t>, grid_for(n_slots - n, 256), 256, ctx->stream, slots.as<uint16_t>(), n, n_slots);
        ctx->launches++;
    }
    const size_t ntl = n_slots / TILE;
    if (ntl > 0x7FFFFFFFull) return fail(ctx, BPE_ERR_INVALID_ARG, "input too large");
    DevBuf tile_live, tile_off, total;
    CU(tile_live.alloc(ntl * 4)); CU(tile_off.alloc(ntl * 8)); CU(total.alloc(8));
    BPE_LAUNCH(tile_count_kernel<uint16_t>, (unsigned)ntl, THREADS, ctx->stream, slots.as<uint16_t>(), tile_live.as<uint32_t>());
    BPE_LAUNCH(tile_scan_kernel, 1, THREADS, ctx->stream, tile_live.as<uint32_t>(), (uint32_t)ntl, tile_off.as<unsigned long long>(),
               total.as<unsigned long long>());
    BPE_LAUNCH((compact_scatter_kernel<uint16_t, uint16_t>), (unsigned)ntl, THREADS, ctx->stream, slots.as<uint16_t>(),
               tile_off.as<unsigned long long>(), d_out);
    ctx->launches += 3;
    CU(cudaGetLastError());
    unsigned long long tot = 0;
    CU(cudaMemcpyAsync(&tot, total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out_n = (size_t)tot;
    return BPE_OK;
}

static bool mentions_id_65535(const bpe_merge_t* merges, size_t m) {
    for (size_t i = 0; i < m; i++)
        if (merges[i].first == 0xFFFF || merges[i].second == 0xFFFF || merges[i].new_token == 0xFFFF) return true;
    return false;
}

// -----------------------------------------------------------------------------------------
// tile-resident encode (bpe_tilenc.cuh): one launch over the text + the final squeeze
// -----------------------------------------------------------------------------------------
static int tile_size_for(const bpe_ctx* ctx, bool has_bridges) {
    int tile_max = (int)ctx->encode_tile;
    return std::max(has_bridges ? 1024 : 512, std::min(tile_max, (int)TN_TMAX)) / 16 * 16;
}
// one launch of tilenc_kernel over n bytes at d_text (plus hl / hr bytes of text before / after it); zone / flags are
// (re)allocated as needed; *d_fail points at the counter of windows that gave up (read it after the stream has drained)
static int launch_tilenc(bpe_ctx* ctx, cudaStream_t stream, const uint8_t* d_text, size_t n, const uint8_t* d_halo_l, int hl, const uint8_t* d_halo_r,
                         int hr, bool left_text, bool right_text, const SegTables& tb, int tile_max, uint16_t* d_slots, DevBuf& zone, DevBuf& flags,
                         uint32_t** d_fail) {
    const TileGeom g = tn_geometry((long long)n, hl, hr, left_text, right_text, tile_max);
    const long long nwin = g.bwl + g.ntile + g.bwr;
    if (nwin > 0x7FFFFFFFll) return fail(ctx, BPE_ERR_INVALID_ARG, "input too large for one launch");
    if (zone.bytes < (size_t)nwin * TN_ZONE * 2) CU(zone.alloc((size_t)nwin * TN_ZONE * 2));
    if (flags.bytes < ((size_t)nwin + 6) * 4) CU(flags.alloc(((size_t)nwin + 6) * 4));  // [flags | ticket | fail | 4 debug counters]
    CU(cudaMemsetAsync(flags.p, 0, ((size_t)nwin + 6) * 4, stream));
    uint32_t* d_ticket = flags.as<uint32_t>() + nwin;
    *d_fail = d_ticket + 1;
    const size_t smem = tilenc_smem_bytes(tile_max, tb.max_level);
    CU(cudaFuncSetAttribute(tilenc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n > 0) {
        BPE_LAUNCH_SMEM(tilenc_kernel, (unsigned)nwin, TN_THREADS, smem, stream, d_text, d_halo_l, d_halo_r, g, tile_max, tb.T,
                        tb.max_level, d_slots, zone.as<uint16_t>(), flags.as<uint32_t>(), d_ticket, *d_fail,
                        ctx->debug ? *d_fail + 1 : (uint32_t*)nullptr);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return BPE_OK;
}

// *used = false when this path does not apply (irregular list, id 65535, more than TL_MAXLVL levels, a shard
// shorter than TN_MIN_SHARD on several GPUs) or gave up (seam without a common token, a token longer than 255
// bytes, a run of more than 255 equal tokens): nothing has been written to d_out then and the caller goes on
// to the next encoder.
static int encode_tiles(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m, uint16_t* d_out,
                        size_t* out_n, bpe_stats_t* st, bool* used) {
    *used = false;
    if (m == 0) return BPE_OK;
    std::vector<uint32_t> lvl;
    std::vector<uint8_t> alone;
    if (!merge_levels(merges, m, lvl, alone) || mentions_id_65535(merges, m)) return BPE_OK;
    SegTables tb;
    int rc = build_seg_tables(ctx, merges, m, lvl, tb);
    if (rc) return rc;
    if (tb.max_level > TL_MAXLVL) return BPE_OK;
    const int world = ctx->dist.world, rank = ctx->dist.rank;
    const bool multi = world > 1;
    ShardHalo sh;
    rc = exchange_shard_halo(ctx, d_text, n, sh);
    if (rc) return rc;
    if (multi && sh.min_shard < (long long)TN_MIN_SHARD) return BPE_OK;  // (the same decision on every rank)
    const bool left_text = multi && sh.g_lo > 0, right_text = multi && sh.g_lo + (long long)n < sh.g_total;
    const int tile_max = tile_size_for(ctx, left_text || right_text);
    const size_t n_slots = round_up(n ? n : 1, TILE);
    DevBuf slots, zone, flags;
    CU(slots.alloc(n_slots * 2));
    uint32_t* d_fail = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, ctx->stream));
    rc = launch_tilenc(ctx, ctx->stream, d_text, n, sh.left(), sh.hl, sh.right(), sh.hr, left_text, right_text, tb, tile_max, slots.as<uint16_t>(),
                       zone, flags, &d_fail);
    if (rc) return rc;
    CU(cudaEventRecord(e1, ctx->stream));
    // a window that gave up anywhere (on any rank) sends every rank to the next encoder
    if (multi && !ctx->dist.allreduce(d_fail, 1, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the seam status failed");
    uint32_t nfail = 0;
    CU(cudaMemcpyAsync(&nfail, d_fail, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float kms = 0;
    cudaEventElapsedTime(&kms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st) { st->kernel_ms[11] = kms; st->kernel_calls[11] = nfail ? 2 : 1; }
    if (ctx->debug) {
        uint32_t dbg[4] = {0, 0, 0, 0};
        cudaMemcpy(dbg, d_fail + 1, sizeof dbg, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[bpe r%d] tile encode: %zu bytes, tile %d, levels %u, %.3f ms; per window: %.1f ranges, %.1f one-warp rounds, %.1f CTA rounds, %.1f partial rounds; %u windows gave up\n",
                rank, n, tile_max, tb.max_level, kms, dbg[0] / (double)std::max<size_t>(1, n / tile_max), dbg[1] / (double)std::max<size_t>(1, n / tile_max),
                dbg[2] / (double)std::max<size_t>(1, n / tile_max), dbg[3] / (double)std::max<size_t>(1, n / tile_max), nfail);
    }
    if (nfail) return BPE_OK;
    *used = true;
    if (n == 0) { *out_n = 0; return BPE_OK; }
    rc = squeeze_slots(ctx, slots, n, d_out, out_n);
    if (rc) return rc;
    if (st) st->scanned_slots += n;
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// streaming encode from host buffers (the reference reads the whole file into memory, src/utils/read_file.zig:3-13, and
// encodes it in place; here the text never has to fit the GPU): the text is cut into chunks, chunk c + 1 travels
// host -> device and the ids of chunk c - 1 device -> host while the tile-resident kernel encodes chunk c. A chunk is
// encoded exactly like a shard of a multi-GPU run — 512 bytes of the neighbouring chunks as halos, bridge windows
// at both ends — so the concatenation of the chunks' ids is the encoding of the whole text. Device memory: two text
// chunks, one slot array, two id buffers (8 bytes per chunk byte) whatever the size of the input.
// *used = false when the path does not apply or a window gave up: the caller then encodes the resident way.
// -----------------------------------------------------------------------------------------
static int encode_stream(bpe_ctx* ctx, const uint8_t* text, size_t n, const bpe_merge_t* merges, size_t m, uint16_t* out, size_t* out_n,
                         bpe_stats_t* st, bool* used) {
    *used = false;
    const size_t chunk_target = ctx->stream_chunk_bytes > 0 ? (size_t)ctx->stream_chunk_bytes : (size_t)ctx->stream_chunk_mb << 20;
    if (m == 0 || chunk_target < 4096 || n < 2 * chunk_target) return BPE_OK;
    std::vector<uint32_t> lvl;
    std::vector<uint8_t> alone;
    if (!merge_levels(merges, m, lvl, alone) || mentions_id_65535(merges, m)) return BPE_OK;
    SegTables tb;
    int rc = build_seg_tables(ctx, merges, m, lvl, tb);
    if (rc) return rc;
    if (tb.max_level > TL_MAXLVL) return BPE_OK;
    const bool multi = ctx->dist.world > 1;
    ShardHalo sh;
    rc = exchange_shard_halo(ctx, text, n, sh, true);
    if (rc) return rc;
    if (multi && sh.min_shard < (long long)(2 * chunk_target)) return BPE_OK;  // every rank streams, or none (the same decision everywhere)
    const size_t nchunk = (n + chunk_target - 1) / chunk_target;
    const size_t ch = round_up((n + nchunk - 1) / nchunk, 16);  // even chunks: none is shorter than half the target
    const int tile_max = tile_size_for(ctx, true);
    if (!ctx->copy_in) CU(cudaStreamCreate(&ctx->copy_in));
    if (!ctx->copy_out) CU(cudaStreamCreate(&ctx->copy_out));
    DevBuf d_text[2], d_halo[2], d_ids[2], slots, zone, flags;
    for (int b = 0; b < 2; b++) {
        CU(d_text[b].alloc(ch));
        CU(d_halo[b].alloc(2 * SEG_HALO));
        CU(d_ids[b].alloc(ch * 2));
    }
    CU(slots.alloc(round_up(ch, TILE) * 2));
    cudaEvent_t h2d_ev[2], d2h_ev[2];
    for (int b = 0; b < 2; b++) { CU(cudaEventCreate(&h2d_ev[b])); CU(cudaEventCreate(&d2h_ev[b])); }
    std::vector<uint8_t> hb[2];
    auto chunk_lo = [&](size_t c) { return std::min(n, c * ch); };
    auto send = [&](size_t c) -> int {  // text and halos of chunk c, host -> device, on the copy-in stream
        const int b = (int)(c & 1);
        const size_t lo = chunk_lo(c), hi = chunk_lo(c + 1);
        hb[b].assign(2 * SEG_HALO, 0);
        int hl, hr;
        if (c == 0) { hl = sh.hl; memcpy(hb[b].data(), sh.host.data(), (size_t)hl); }
        else { hl = (int)std::min<size_t>(SEG_HALO, lo); memcpy(hb[b].data(), text + lo - hl, (size_t)hl); }
        if (c + 1 == nchunk) { hr = sh.hr; memcpy(hb[b].data() + SEG_HALO, sh.host.data() + SEG_HALO, (size_t)hr); }
        else { hr = (int)std::min<size_t>(SEG_HALO, n - hi); memcpy(hb[b].data() + SEG_HALO, text + hi, (size_t)hr); }
        CU(cudaMemcpyAsync(d_halo[b].p, hb[b].data(), 2 * SEG_HALO, cudaMemcpyHostToDevice, ctx->copy_in));
        CU(cudaMemcpyAsync(d_text[b].p, text + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_in));
        CU(cudaEventRecord(h2d_ev[b], ctx->copy_in));
        return BPE_OK;
    };
    const double t0 = now_ms();
    rc = send(0);
    if (rc) return rc;
    size_t out_off = 0, prev_cnt = 0;
    bool gave_up = false;
    for (size_t c = 0; c < nchunk && !gave_up; c++) {
        const int b = (int)(c & 1);
        const size_t lo = chunk_lo(c), hi = chunk_lo(c + 1), len = hi - lo;
        const bool left_text = c > 0 || (multi && sh.g_lo > 0), right_text = c + 1 < nchunk || (multi && sh.g_lo + (long long)n < sh.g_total);
        const int hl = c == 0 ? sh.hl : (int)std::min<size_t>(SEG_HALO, lo), hr = c + 1 == nchunk ? sh.hr : (int)std::min<size_t>(SEG_HALO, n - hi);
        CU(cudaStreamWaitEvent(ctx->stream, h2d_ev[b], 0));
        uint32_t* d_fail = nullptr;
        rc = launch_tilenc(ctx, ctx->stream, d_text[b].as<uint8_t>(), len, d_halo[b].as<uint8_t>(), hl, d_halo[b].as<uint8_t>() + SEG_HALO, hr,
                           left_text, right_text, tb, tile_max, slots.as<uint16_t>(), zone, flags, &d_fail);
        if (rc) return rc;
        if (c + 1 < nchunk) { rc = send(c + 1); if (rc) return rc; }  // travels while chunk c is being encoded
        if (c > 0) {  // the ids of chunk c - 1 leave while chunk c is being encoded
            CU(cudaMemcpyAsync(out + out_off, d_ids[b ^ 1].p, prev_cnt * 2, cudaMemcpyDeviceToHost, ctx->copy_out));
            CU(cudaEventRecord(d2h_ev[b ^ 1], ctx->copy_out));
            out_off += prev_cnt;
        }
        if (c >= 2) CU(cudaStreamWaitEvent(ctx->stream, d2h_ev[b], 0));  // d_ids[b] still held the ids of chunk c - 2
        size_t cnt = 0;
        rc = squeeze_slots(ctx, slots, len, d_ids[b].as<uint16_t>(), &cnt);  // (synchronises the compute stream)
        if (rc) return rc;
        uint32_t nfail = 0;
        CU(cudaMemcpyAsync(&nfail, d_fail, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (nfail) gave_up = true;
        prev_cnt = cnt;
    }
    // a window that gave up anywhere (on any rank) sends every rank to the resident encoders
    if (multi) {
        DevBuf fl;
        CU(fl.alloc(4));
        uint32_t v = gave_up ? 1u : 0u;
        CU(cudaMemcpyAsync(fl.p, &v, 4, cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->dist.allreduce(fl.p, 1, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the seam status failed");
        CU(cudaMemcpyAsync(&v, fl.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        gave_up = v != 0;
    }
    if (!gave_up) {
        const int b = (int)((nchunk - 1) & 1);
        CU(cudaMemcpyAsync(out + out_off, d_ids[b].p, prev_cnt * 2, cudaMemcpyDeviceToHost, ctx->copy_out));
        out_off += prev_cnt;
    }
    CU(cudaStreamSynchronize(ctx->copy_in));
    CU(cudaStreamSynchronize(ctx->copy_out));
    for (int b = 0; b < 2; b++) { cudaEventDestroy(h2d_ev[b]); cudaEventDestroy(d2h_ev[b]); }
    if (st) { st->kernel_ms[11] = now_ms() - t0; st->kernel_calls[11] = gave_up ? 2 : 1; st->kernel_calls[9] = nchunk; }
    if (gave_up) return BPE_OK;
    *used = true;
    *out_n = out_off;
    if (st) { st->kernel_calls[10] = 2; st->scanned_slots += n; }
    return BPE_OK;
}

// *used = false when this path does not apply (irregular list, id 65535, seam without a common token):
// nothing has been written to d_out then and the caller runs the level-scheduled passes.
static int encode_segments(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m, uint16_t* d_out,
                           size_t* out_n, bpe_stats_t* st, bool* used) {
    *used = false;
    if (m == 0) return BPE_OK;
    std::vector<uint32_t> lvl;
    std::vector<uint8_t> alone;
    if (!merge_levels(merges, m, lvl, alone) || mentions_id_65535(merges, m)) return BPE_OK;
    const int world = ctx->dist.world, rank = ctx->dist.rank;
    const bool multi = world > 1;
    SegTables tb;
    int rc = build_seg_tables(ctx, merges, m, lvl, tb);
    if (rc) return rc;
    const SegTab& T = tb.T;
    ShardHalo sh;
    rc = exchange_shard_halo(ctx, d_text, n, sh);
    if (rc) return rc;
    const long long g_lo = sh.g_lo, g_total = sh.g_total;
    const int hl = sh.hl, hr = sh.hr;
    // ---- geometry ----
    int C = 64;
    if (ctx->encode_geom == 1) C = 128;
    else if (ctx->encode_geom == 2 || ctx->encode_geom == 5) C = 96;
    else if (ctx->encode_geom == 3) C = 32;
    const long long g_hi = g_lo + (long long)n;
    const long long k_text_last = g_total > 0 ? (g_total - 1) / C : 0;
    long long k_first = g_lo / C - 2, k_last = g_hi / C + 2;
    if (k_first < 0) k_first = 0;
    if (k_last > k_text_last) k_last = k_text_last;
    const size_t n_slots = round_up(n ? n : 1, TILE);
    DevBuf slots, failbuf;
    CU(slots.alloc(n_slots * 2));
    CU(failbuf.alloc(4));
    CU(cudaMemsetAsync(failbuf.p, 0, 4, ctx->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, ctx->stream));
    if (n > 0 && g_total > 0) {
        const uint8_t* hlp = sh.left();
        const uint8_t* hrp = sh.right();
        uint16_t* sl = slots.as<uint16_t>();
        uint32_t* fl = failbuf.as<uint32_t>();
        switch (ctx->encode_geom) {
            case 4: rc = launch_segenc<64, 32, 64>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
            case 5: rc = launch_segenc<96, 32, 64>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
            case 1: rc = launch_segenc<128, 32, 128>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
            case 2: rc = launch_segenc<96, 32, 128>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
            case 3: rc = launch_segenc<32, 12, 32>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
            default: rc = launch_segenc<64, 32, 128>(ctx, d_text, (long long)n, hlp, hl, hrp, hr, g_lo, k_first, k_last, T, sl, fl); break;
        }
        if (rc) return rc;
    }
    CU(cudaEventRecord(e1, ctx->stream));
    // a seam without a common token anywhere (on any rank) sends every rank to the level-scheduled passes
    if (multi && !ctx->dist.allreduce(failbuf.p, 1, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the seam status failed");
    uint32_t nfail = 0;
    CU(cudaMemcpyAsync(&nfail, failbuf.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float kms = 0;
    cudaEventElapsedTime(&kms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st) { st->kernel_ms[11] = kms; st->kernel_calls[11] = nfail ? 2 : 1; }
    if (nfail) {
        if (ctx->debug) fprintf(stderr, "[bpe r%d] segment encode: %u seams without a common token, falling back to level passes\n", rank, nfail);
        return BPE_OK;
    }
    *used = true;
    if (n == 0) { *out_n = 0; return BPE_OK; }
    rc = squeeze_slots(ctx, slots, n, d_out, out_n);
    if (rc) return rc;
    if (st) st->scanned_slots += n;
    return BPE_OK;
}

template <class TokT>
static int encode_passes(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m,
                         uint16_t* d_out, size_t* out_n, bpe_stats_t* st) {
    Sequence<TokT> sq;
    int rc = seq_init(ctx, sq, d_text, n);
    if (rc) return rc;
    DevBuf ctl;
    CU(ctl.alloc(sizeof(StepCtl)));
    CU(cudaMemsetAsync(ctl.p, 0, sizeof(StepCtl), ctx->stream));
    StepCtl* d_ctl = ctl.as<StepCtl>();
    // multi-GPU: shards exchange their end tokens before every pass (one small all-reduce), so a pair
    // that straddles two shards is merged exactly as on one GPU (X stays left, the hole goes right)
    const bool multi = ctx->dist.world > 1;
    DevBuf edges_buf, flag_buf;
    EdgeInfo* edges = nullptr;
    if (multi) {
        CU(edges_buf.alloc((size_t)ctx->dist.world * sizeof(EdgeInfo)));
        CU(flag_buf.alloc(4));
        CU(cudaMemsetAsync(edges_buf.p, 0, (size_t)ctx->dist.world * sizeof(EdgeInfo), ctx->stream));
        edges = edges_buf.as<EdgeInfo>();
    }
    uint32_t merged_seen = 0;  // device cntAB accumulates over passes; host subtracts what it has seen
    uint32_t pass_index = 0;   // alternate the scan direction so consecutive passes reuse the L2
    auto read_merged = [&](uint32_t* fresh) -> int {
        uint32_t acc = 0;
        CU(cudaMemcpyAsync(&acc, &d_ctl->cntAB, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *fresh = acc - merged_seen;
        merged_seen = acc;
        sq.live -= *fresh;
        return BPE_OK;
    };
    std::vector<EncStep> steps;
    std::vector<LevelEntry> ents;
    build_encode_schedule(merges, m, ctx->encode_impl != 1, steps, ents);
    DevBuf ents_buf;
    if (!ents.empty()) {
        CU(ents_buf.alloc(ents.size() * sizeof(LevelEntry)));
        CU(cudaMemcpyAsync(ents_buf.p, ents.data(), ents.size() * sizeof(LevelEntry), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaFuncSetAttribute(level_kernel<TokT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)level_smem_bytes<TokT>(LVL_HASH_MAX)));
        CU(cudaFuncSetAttribute(level_kernel<TokT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)level_smem_bytes<TokT>(LVL_HASH_MAX, LVL_BYTE_IDS_MAX / 4)));
        CU(cudaFuncSetAttribute(level_kernel<TokT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)level_smem_bytes<TokT>(LVL_HASH_MAX, 2 * LVL_ROLE_WORDS)));
    }
    // encode_filter = 1: byte role map + successor filter (level_kernel MODE 1) when every id fits the byte map
    uint32_t max_id = 255;
    for (size_t i = 0; i < m; i++) max_id = std::max(max_id, (uint32_t)std::max(merges[i].first, std::max(merges[i].second, merges[i].new_token)));
    const bool pair_bloom = ctx->encode_filter == 2;
    const bool byte_roles = ctx->encode_filter == 1 && max_id < (uint32_t)LVL_BYTE_IDS_MAX;
    const uint32_t role_words = pair_bloom ? 2u * (uint32_t)LVL_ROLE_WORDS : byte_roles ? ((max_id + 1u + 15u) / 16u) * 4u : (uint32_t)LVL_ROLE_WORDS;
    // ent_cnt == 0: one merge pass for (A,B) -> X; else one level pass over ents[ent_off, ent_off + ent_cnt)
    auto one_pass = [&](uint32_t A, uint32_t B, uint32_t X, uint32_t ent_off = 0, uint32_t ent_cnt = 0) -> int {
        const uint32_t nt = sq.ntiles();
        if (ent_cnt) { A = 1; B = X = 0; }  // the halo of a level pass carries no run of equal tokens
        if (multi) {
            BPE_LAUNCH(edge_kernel<TokT>, 1, 32, ctx->stream, sq.tok(), sq.n_slots, sq.dense_end, edges, ctx->dist.rank, ctx->dist.world, d_ctl,
                       &d_ctl->cntAB, 0);
            ctx->launches += 1;
            if (!ctx->dist.allreduce(edges, (size_t)ctx->dist.world * 16, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce of the shard edges failed");
        }
        BPE_LAUNCH((halo_kernel<TokT, false>), (nt + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, ctx->stream, sq.tok(), sq.n_slots,
                   nt, sq.halo.template as<TileHalo<TokT>>(), (const StepCtl*)nullptr, A, A == B ? 1 : 0,
                   sq.run_local.template as<uint32_t>(), sq.run_full.template as<uint8_t>(), sq.done_counter.template as<uint32_t>(),
                   (uint32_t*)nullptr, (const EdgeInfo*)edges, ctx->dist.rank, ctx->dist.world);
        ctx->launches += 1;
        if (ent_cnt) {
            const int backwards = (int)(pass_index++ & 1u);
            uint32_t hash_log2 = 6;
            while ((1u << hash_log2) < 2u * ent_cnt) hash_log2++;
            // encode_grid CTAs per SM take the tiles round-robin (the level's tables are built once per CTA); 0: one CTA per tile
            const uint32_t grid = ctx->encode_grid > 0   ? std::min<uint32_t>(nt, (uint32_t)ctx->encode_grid * (uint32_t)ctx->num_sms)
                                  : ctx->encode_grid < 0 ? std::min<uint32_t>(nt, (uint32_t)(-ctx->encode_grid))  // absolute CTA count (tests)
                                                         : nt;
            const size_t smem = level_smem_bytes<TokT>((int)(1u << hash_log2), (int)role_words);
            const TileHalo<TokT>* halo_p = sq.halo.template as<TileHalo<TokT>>();
            const LevelEntry* ents_p = (const LevelEntry*)ents_buf.template as<LevelEntry>() + ent_off;
            if (pair_bloom)
                BPE_LAUNCH_SMEM((level_kernel<TokT, 2>), grid, THREADS, smem, ctx->stream, sq.tok(), halo_p, ents_p, ent_cnt, &d_ctl->cntAB, backwards,
                                nt, hash_log2, role_words);
            else if (byte_roles)
                BPE_LAUNCH_SMEM((level_kernel<TokT, 1>), grid, THREADS, smem, ctx->stream, sq.tok(), halo_p, ents_p, ent_cnt, &d_ctl->cntAB, backwards,
                                nt, hash_log2, role_words);
            else
                BPE_LAUNCH_SMEM((level_kernel<TokT, 0>), grid, THREADS, smem, ctx->stream, sq.tok(), halo_p, ents_p, ent_cnt, &d_ctl->cntAB, backwards,
                                nt, hash_log2, role_words);
            ctx->launches++;
        } else {
            int rcm = launch_merge<TokT, false, false>(ctx, sq.tok(), sq.halo.template as<TileHalo<TokT>>(), nt, (const StepCtl*)nullptr,
                                                       (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, &d_ctl->cntAB, A, B, X, 0u,
                                                       (int)(pass_index++ & 1u));
            if (rcm) return rcm;
        }
        CU(cudaGetLastError());
        if (st) st->scanned_slots += sq.n_slots;
        return BPE_OK;
    };
    // encode_fuse = K > 1: up to K consecutive level steps go through ONE residency of every tile (bpe_groupenc.cuh)
    const int fuse_k = (std::is_same<TokT, uint16_t>::value && !multi && max_id < (uint32_t)LVL_BYTE_IDS_MAX)
                           ? (int)std::min<long>(std::max<long>(ctx->encode_fuse, 0), GRP_MAXLV) : 0;
    DevBuf deep_halo, grp_res;
    // groups start small (the first levels merge a third of the tokens each, halos included), grow by one level per group that went
    // through and are halved when a group's halos did not cover its levels
    int fuse_now = std::min(fuse_k, 2);
    // steps [s0, s1): level steps with adjacent entries in `ents`. *ok = false: the taint reached a tile, nothing was changed.
    auto group_pass = [&](size_t s0, size_t s1, bool* ok) -> int {
        const uint32_t nt = sq.ntiles();
        if (!deep_halo.p) {
            CU(deep_halo.alloc((size_t)nt * sizeof(DeepHalo)));  // (the first group sees the largest tile count)
            CU(grp_res.alloc(sizeof(GroupResult)));
            CU(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem_bytes(GRP_HASH_MAX, LVL_BYTE_IDS_MAX)));
        }
        GroupDesc gd;
        memset(&gd, 0, sizeof gd);
        uint32_t total = 0;
        for (size_t k = s0; k < s1; k++) { total += steps[k].cnt; gd.end[k - s0] = total; }
        gd.nlev = (uint32_t)(s1 - s0);
        for (uint32_t k = gd.nlev; k < (uint32_t)GRP_MAXLV; k++) gd.end[k] = total;
        uint32_t hash_log2 = 6;
        while ((1u << hash_log2) < 2u * total) hash_log2++;
        const uint32_t role_bytes = ((max_id + 1u + 15u) / 16u) * 16u;
        CU(cudaMemsetAsync(grp_res.p, 0, sizeof(GroupResult), ctx->stream));
        BPE_LAUNCH(deep_halo_kernel, (nt + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, ctx->stream, reinterpret_cast<const uint16_t*>(sq.tok()),
                   sq.n_slots, nt, deep_halo.template as<DeepHalo>());
        const uint32_t grid = ctx->encode_grid > 0   ? std::min<uint32_t>(nt, (uint32_t)ctx->encode_grid * (uint32_t)ctx->num_sms)
                              : ctx->encode_grid < 0 ? std::min<uint32_t>(nt, (uint32_t)(-ctx->encode_grid))
                                                     : nt;
        const int backwards = (int)(pass_index++ & 1u);  // (not inside the launch: the emulation evaluates the arguments once per thread)
        BPE_LAUNCH_SMEM(group_kernel, grid, THREADS, group_smem_bytes((int)(1u << hash_log2), (int)role_bytes), ctx->stream,
                        reinterpret_cast<const uint16_t*>(sq.tok()), reinterpret_cast<uint16_t*>(sq.other()),
                        (const DeepHalo*)deep_halo.template as<DeepHalo>(),
                        (const LevelEntry*)ents_buf.template as<LevelEntry>() + steps[s0].off, gd, grp_res.template as<GroupResult>(), backwards, nt,
                        hash_log2, role_bytes);
        ctx->launches += 2;
        CU(cudaGetLastError());
        GroupResult gr;
        CU(cudaMemcpyAsync(&gr, grp_res.p, sizeof gr, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (st) st->scanned_slots += sq.n_slots;
        *ok = gr.fail == 0;
        if (st) st->kernel_calls[*ok ? 9 : 8]++;  // encode: fused level groups that went through / that were redone level by level
        if (*ok) {  // the other buffer now holds the sequence
            sq.cur ^= 1;
            sq.live -= gr.merged;
        }
        return BPE_OK;
    };
    size_t since_check = 0;
    for (size_t si = 0; si < steps.size() && (n > 1 || multi); si++) {
        const EncStep& es = steps[si];
        bool last_step = si + 1 == steps.size();
        if (es.single < 0 && fuse_k > 1) {
            size_t sj = si + 1;
            uint32_t total = es.cnt;
            while (sj < steps.size() && steps[sj].single < 0 && (int)(sj - si) < fuse_now && steps[sj].off == steps[sj - 1].off + steps[sj - 1].cnt &&
                   total + steps[sj].cnt <= (uint32_t)GRP_PAIRS_MAX) { total += steps[sj].cnt; sj++; }
            bool grouped = false;
            if (sj - si > 1) {
                rc = group_pass(si, sj, &grouped);
                if (rc) return rc;
                fuse_now = grouped ? std::min(fuse_k, fuse_now + 1) : std::max(1, fuse_now / 2);
            } else if (fuse_now < fuse_k) {
                fuse_now++;  // (single steps after a failure: try groups again, growing)
            }
            if (grouped) {
                si = sj - 1;
                last_step = si + 1 == steps.size();
                since_check = 32;  // (the group's merges are in sq.live already; the compaction check below runs)
            } else {
                rc = one_pass(0, 0, 0, es.off, es.cnt);
                if (rc) return rc;
                since_check += (pass_index <= 8u) ? 32 : 8;
            }
        } else if (es.single < 0) {
            // a level pass can remove a large share of the tokens: the first ones do most of the merging, so the live count
            // is read back (one host round trip) after each of the first 8 level passes and after every 4th later on
            rc = one_pass(0, 0, 0, es.off, es.cnt);
            if (rc) return rc;
            since_check += (pass_index <= 8u) ? 32 : 8;
        } else {
            const uint32_t A = merges[es.single].first, B = merges[es.single].second, X = merges[es.single].new_token;
            rc = one_pass(A, B, X);
            if (rc) return rc;
            since_check++;
            // the reference does not advance `i` after a hit (:78-81), so a merge whose new token
            // equals its own first component keeps absorbing: repeat until a pass changes nothing
            while (X == A) {
                uint32_t fresh = 0;
                rc = read_merged(&fresh);
                if (rc) return rc;
                since_check = 0;
                if (multi) {  // "nothing changed" must hold on every shard
                    CU(cudaMemcpyAsync(flag_buf.p, &fresh, 4, cudaMemcpyHostToDevice, ctx->stream));
                    if (!ctx->dist.allreduce(flag_buf.p, 1, DIST_U32_SUM)) return fail(ctx, BPE_ERR_CUDA, "all-reduce failed");
                    CU(cudaMemcpyAsync(&fresh, flag_buf.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                }
                if (!fresh) break;
                rc = one_pass(A, B, X);
                if (rc) return rc;
            }
        }
        if (since_check >= 32 || last_step) {
            uint32_t fresh = 0;
            rc = read_merged(&fresh);
            if (rc) return rc;
            since_check = 0;
            if (sq.n_slots > (size_t)TILE && sq.live * 100 < (uint64_t)sq.n_slots * (uint64_t)ctx->encode_compact_pct && !last_step) {
                rc = seq_compact(ctx, sq, nullptr);
                if (rc) return rc;
                if (st) st->compactions++;
            }
        }
    }
    // final squeeze straight into the caller's u16 buffer
    const uint32_t nt = sq.ntiles();
    BPE_LAUNCH(tile_count_kernel<TokT>, nt, THREADS, ctx->stream, sq.tok(), sq.tile_live.template as<uint32_t>());
    BPE_LAUNCH(tile_scan_kernel, 1, THREADS, ctx->stream, sq.tile_live.template as<uint32_t>(), nt,
               sq.tile_off.template as<unsigned long long>(), sq.total.template as<unsigned long long>());
    BPE_LAUNCH((compact_scatter_kernel<TokT, uint16_t>), nt, THREADS, ctx->stream, sq.tok(),
               sq.tile_off.template as<unsigned long long>(), d_out);
    ctx->launches += 3;
    CU(cudaGetLastError());
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, sq.total.p, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out_n = (size_t)total;
    return BPE_OK;
}

static int encode_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m,
                         uint16_t* d_out, size_t* out_n, bpe_stats_t* stats_out) {
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (m && !merges) return fail(ctx, BPE_ERR_INVALID_ARG, "merges is null");
    bpe_stats_t st;
    memset(&st, 0, sizeof st);
    const double t0 = now_ms();
    const uint64_t l0 = ctx->launches;
    if (n == 0 && ctx->dist.world == 1) { if (stats_out) *stats_out = st; return BPE_OK; }
    if (n && !d_out) return fail(ctx, BPE_ERR_INVALID_ARG, "out is null");
    CU(cudaSetDevice(ctx->device));
    cudaEvent_t ev0, ev1;
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    CU(cudaEventRecord(ev0, ctx->stream));
    int rc = BPE_OK;
    bool done = false;
    // encode_impl 0 (default) picks by measured cost (B200). The level passes sweep the resident
    // sequence once per level of the list (C3's 7,936 merges: ~190 sweeps, 0.09 s per GB on one GPU) but need one exchange
    // and one host round trip per sweep across GPUs and 32-bit positions; the tile-resident kernel (one launch, ~0.15 s
    // per GB for such a list, nothing exchanged between GPUs) takes the multi-GPU runs, inputs of 4 GiB or more and the
    // lists whose schedule is longer than "encode_seg_min_steps". 4 / 3 force the tile / segment kernel, 2 the level
    // passes, 1 one pass per merge. kernel_calls[10] records the encoder that produced the ids (1 segment, 2 tile; 0 passes).
    size_t sched_steps = 0;
    if (ctx->encode_impl == 0 && m > 0) {
        std::vector<EncStep> steps;
        std::vector<LevelEntry> ents;
        build_encode_schedule(merges, m, true, steps, ents);
        sched_steps = steps.size();
    }
    const bool prefer_tiles = ctx->dist.world > 1 || n >= 0xFFFFFFF0ull || sched_steps > (size_t)ctx->encode_seg_min_steps || ctx->encode_try_tiles > 1;
    if ((ctx->encode_impl == 0 && ctx->encode_try_tiles && prefer_tiles) || ctx->encode_impl == 4) {
        rc = encode_tiles(ctx, d_text, n, merges, m, d_out, out_n, &st, &done);
        if (rc) return rc;
        if (done) st.kernel_calls[10] = 2;
        else if (ctx->encode_impl == 4) return fail(ctx, BPE_ERR_INTERNAL, "encode_impl = 4: the tile-resident encoder does not apply to this input");
    }
    bool try_segments = ctx->encode_impl == 3;
    if (!done && ctx->encode_impl == 0 && m > 0) try_segments = sched_steps > (size_t)ctx->encode_seg_min_steps || n >= 0xFFFFFFF0ull;
    if (!done && try_segments) {
        const uint64_t tile_verdict = st.kernel_calls[11];
        rc = encode_segments(ctx, d_text, n, merges, m, d_out, out_n, &st, &done);
        if (rc) return rc;
        if (done) st.kernel_calls[10] = 1;
        else if (ctx->encode_impl == 3) return fail(ctx, BPE_ERR_INTERNAL, "encode_impl = 3: the segment-resident encoder does not apply to this input");
        if (!done && tile_verdict == 2) st.kernel_calls[11] = 2;
    }
    if (!done) {
        if (n >= 0xFFFFFFF0ull) return fail(ctx, BPE_ERR_INVALID_ARG, "input of %zu bytes exceeds the 32-bit position range of the pass-based encoder", n);
        // id 65535 is the u16 hole marker: lists that mention it run on u32 slots
        bool wide = false;
        for (size_t i = 0; i < m; i++)
            if (merges[i].first == 0xFFFF || merges[i].second == 0xFFFF || merges[i].new_token == 0xFFFF) wide = true;
        rc = wide ? encode_passes<uint32_t>(ctx, d_text, n, merges, m, d_out, out_n, &st)
                  : encode_passes<uint16_t>(ctx, d_text, n, merges, m, d_out, out_n, &st);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ev1, ctx->stream));
    CU(cudaEventSynchronize(ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    st.device_ms = ms;
    st.total_ms = now_ms() - t0;
    st.kernel_launches = ctx->launches - l0;
    st.replace_pair_calls = m;
    st.replace_pair_ms = ms;
    if (stats_out) *stats_out = st;
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// decode (src/basic_tokenizer.zig:90-138)
// -----------------------------------------------------------------------------------------
// findMerge (:109-116) returns the FIRST merge whose new_token matches; decodeMerge (:118-138) expands first then
// second, recursively. Ids whose expansion is undefined (unknown component or a cycle, which would overflow the
// reference's stack) get length 0 and fail only if they are used. Lengths are computed over the merge DAG (no
// expansion is materialised for that), expansions of up to VOC_FLAT_MAX bytes are flattened for the scatter kernel,
// longer ones (e.g. a doubling chain over a run of one byte) are decoded by descending through the DAG, so there
// is no limit on the length of a token's expansion short of the 32-bit length itself.
struct DecodeVocabHost {
    std::vector<uint32_t> len, off, pair;  // per id
    std::vector<uint8_t> bytes;
};
static const uint32_t VOC_FLAT_MAX = 4096;
static const size_t VOC_FLAT_TOTAL = (size_t)64 << 20;
static const uint32_t VOC_LEN_SAT = 0xFFFFFFFEu;  // expansions of 4 GiB or more: the decode of such a token cannot be addressed

static void build_vocab(const bpe_merge_t* merges, size_t m, DecodeVocabHost& v) {
    const uint32_t NID = 65536;
    std::vector<int32_t> def(NID, -1);
    for (size_t i = 0; i < m; i++)
        if (def[merges[i].new_token] < 0) def[merges[i].new_token] = (int32_t)i;
    std::vector<uint8_t> state(NID, 0);  // 0 new, 1 open, 2 done
    v.len.assign(NID, 0);
    v.pair.assign(NID, 0);
    v.off.assign(NID, VOC_NOT_FLAT);
    for (uint32_t b = 0; b < 256; b++) { state[b] = 2; v.len[b] = 1; }
    std::vector<uint32_t> stack, order;  // order: ids in an order where components come first
    for (uint32_t id = 256; id < NID; id++) {
        if (state[id] || def[id] < 0) continue;
        stack.push_back(id);
        while (!stack.empty()) {
            uint32_t t = stack.back();
            if (state[t] == 2) { stack.pop_back(); continue; }
            state[t] = 1;
            const bpe_merge_t& mg = merges[def[t]];
            uint32_t parts[2] = {mg.first, mg.second};
            bool wait = false, bad = false;
            for (uint32_t p : parts) {
                if (p < 256) continue;
                if (def[p] < 0) { bad = true; break; }
                if (state[p] == 0) { stack.push_back(p); wait = true; break; }
                if (state[p] == 1) { bad = true; break; }  // cycle
                if (v.len[p] == 0) { bad = true; break; }
            }
            if (wait) continue;
            if (bad) v.len[t] = 0;
            else {
                const uint64_t l = (uint64_t)v.len[parts[0]] + v.len[parts[1]];
                v.len[t] = l >= VOC_LEN_SAT ? VOC_LEN_SAT : (uint32_t)l;
                v.pair[t] = parts[0] | (parts[1] << 16);
                order.push_back(t);
            }
            state[t] = 2;
            stack.pop_back();
        }
    }
    // flatten the short expansions (components first, so a token's bytes are the concatenation of two flattened ones)
    v.bytes.resize(256);
    for (uint32_t b = 0; b < 256; b++) { v.bytes[b] = (uint8_t)b; v.off[b] = b; }
    for (uint32_t t : order) {
        const uint32_t f = v.pair[t] & 0xFFFFu, s2 = v.pair[t] >> 16;
        if (v.len[t] > VOC_FLAT_MAX || v.off[f] == VOC_NOT_FLAT || v.off[s2] == VOC_NOT_FLAT || v.bytes.size() + v.len[t] > VOC_FLAT_TOTAL) continue;
        v.off[t] = (uint32_t)v.bytes.size();
        const size_t at = v.bytes.size();
        v.bytes.resize(at + v.len[t]);
        memcpy(&v.bytes[at], &v.bytes[v.off[f]], v.len[f]);
        memcpy(&v.bytes[at + v.len[f]], &v.bytes[v.off[s2]], v.len[s2]);
    }
}

// device copy of the decode tables, kept in the context for the merge list they were built from
struct DecodeCache {
    uint64_t fingerprint = 0;
    size_t m = 0;
    bool valid = false;
    DevBuf len, off, pair, bytes;
    bool has_saturated = false;
    DecVocab view() const { DecVocab d; d.len = len.as<uint32_t>(); d.off = off.as<uint32_t>(); d.pair = pair.as<uint32_t>(); d.bytes = bytes.as<uint8_t>(); return d; }
};
static uint64_t merges_fingerprint(const bpe_merge_t* merges, size_t m) {
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)m;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(merges);
    for (size_t i = 0; i < m * sizeof(bpe_merge_t); i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
static int ensure_decode_vocab(bpe_ctx* ctx, const bpe_merge_t* merges, size_t m);

__global__ void decode_flag_big_kernel(const uint16_t* __restrict__ toks, size_t n, const uint32_t* __restrict__ voc_len, StepCtl* ctl) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t l = voc_len[toks[i]];
        if (l == 0) atomicOr(&ctl->err, (uint32_t)ERR_KEY_MISSING);
        else if (l >= 0xFFFFFFFEu) atomicOr(&ctl->err, (uint32_t)ERR_TABLE_FULL);
    }
}

// d_out == nullptr: size query only
static int decode_device(bpe_ctx* ctx, const uint16_t* d_toks, size_t n, const bpe_merge_t* merges, size_t m,
                         uint8_t* d_out, size_t cap, size_t* out_n, bpe_stats_t* stats_out) {
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (m && !merges) return fail(ctx, BPE_ERR_INVALID_ARG, "merges is null");
    bpe_stats_t st;
    memset(&st, 0, sizeof st);
    const double t0 = now_ms();
    const uint64_t l0 = ctx->launches;
    if (n == 0) { if (stats_out) *stats_out = st; return BPE_OK; }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_decode_vocab(ctx, merges, m);
    if (rc) return rc;
    DecodeCache& dc = *ctx->decode_cache;
    DevBuf ctl, tile_bytes, tile_off, total;
    const size_t nt = (n + TILE - 1) / TILE;
    if (nt > 0x7FFFFFFFull) return fail(ctx, BPE_ERR_INVALID_ARG, "input too large");
    CU(ctl.alloc(sizeof(StepCtl))); CU(tile_bytes.alloc(nt * 8)); CU(tile_off.alloc(nt * 8)); CU(total.alloc(8));
    cudaEvent_t ev0, ev1;
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    CU(cudaEventRecord(ev0, ctx->stream));
    CU(cudaMemsetAsync(ctl.p, 0, sizeof(StepCtl), ctx->stream));
    BPE_LAUNCH(decode_len_kernel, (unsigned)nt, THREADS, ctx->stream, d_toks, n, dc.len.as<uint32_t>(), tile_bytes.as<unsigned long long>(), ctl.as<StepCtl>());
    ctx->launches++;
    if (dc.has_saturated) {
        BPE_LAUNCH_NS(decode_flag_big_kernel, grid_for(n, 256), 256, ctx->stream, d_toks, n, dc.len.as<uint32_t>(), ctl.as<StepCtl>());
        ctx->launches++;
    }
    BPE_LAUNCH(tile_scan64_kernel, 1, THREADS, ctx->stream, tile_bytes.as<unsigned long long>(), (uint32_t)nt,
               tile_off.as<unsigned long long>(), total.as<unsigned long long>());
    ctx->launches++;
    CU(cudaGetLastError());
    uint32_t err = 0;
    unsigned long long tot = 0;
    CU(cudaMemcpyAsync(&err, &ctl.as<StepCtl>()->err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&tot, total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (err & ERR_KEY_MISSING) return fail(ctx, BPE_ERR_INVALID_TOKEN, "token id without a merge (error.InvalidToken)");
    if (err & ERR_TABLE_FULL) return fail(ctx, BPE_ERR_OOM, "a token expands to 4 GiB or more");
    *out_n = (size_t)tot;
    if (d_out) {
        if ((size_t)tot > cap) return fail(ctx, BPE_ERR_OOM, "decode needs %llu bytes, buffer has %zu", tot, cap);
        BPE_LAUNCH(decode_scatter_kernel, (unsigned)nt, THREADS, ctx->stream, d_toks, n, dc.view(), tile_off.as<unsigned long long>(),
                   tile_bytes.as<unsigned long long>(), d_out, cap);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev1, ctx->stream));
    CU(cudaEventSynchronize(ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    st.device_ms = ms;
    st.total_ms = now_ms() - t0;
    st.kernel_launches = ctx->launches - l0;
    if (stats_out) *stats_out = st;
    return BPE_OK;
}

// -----------------------------------------------------------------------------------------
// C ABI
// -----------------------------------------------------------------------------------------
static int ensure_decode_vocab(bpe_ctx* ctx, const bpe_merge_t* merges, size_t m) {
    if (!ctx->decode_cache) ctx->decode_cache = new DecodeCache();
    DecodeCache& dc = *ctx->decode_cache;
    const uint64_t fp = merges_fingerprint(merges, m);
    if (dc.valid && dc.m == m && dc.fingerprint == fp) return BPE_OK;
    dc.valid = false;
    DecodeVocabHost v;
    build_vocab(merges, m, v);
    dc.has_saturated = false;
    for (uint32_t l : v.len) if (l >= VOC_LEN_SAT) dc.has_saturated = true;
    BufCache* saved = tl_cache;
    tl_cache = nullptr;  // these buffers outlive the call: plain cudaMalloc, freed with the context
    cudaError_t e = dc.len.alloc(65536 * 4);
    if (e == cudaSuccess) e = dc.off.alloc(65536 * 4);
    if (e == cudaSuccess) e = dc.pair.alloc(65536 * 4);
    if (e == cudaSuccess) e = dc.bytes.alloc(v.bytes.size());
    tl_cache = saved;
    if (e != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of the decode tables failed");
    CU(cudaMemcpyAsync(dc.len.p, v.len.data(), 65536 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc.off.p, v.off.data(), 65536 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc.pair.p, v.pair.data(), 65536 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc.bytes.p, v.bytes.data(), v.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // `v` lives on this frame
    dc.fingerprint = fp;
    dc.m = m;
    dc.valid = true;
    return BPE_OK;
}

struct CacheScope {
    BufCache* prev;
    bpe_ctx* ctx;
    explicit CacheScope(bpe_ctx* c) : prev(tl_cache), ctx(c) { tl_cache = c ? &c->cache : nullptr; }
    // at the end of every API call the cache is cut back to "cache_max_mb" so that other allocators of the process
    // (torch, a second context, the caller) are not starved by memory that sits idle here
    ~CacheScope() { if (ctx) ctx->cache.trim((size_t)std::max<long>(0, ctx->cache_max_mb) << 20); tl_cache = prev; }
};

extern "C" {

const char* bpe_version(void) {
#ifdef BPE_EMUL
    return "bpe_b200 0.1 (CPU emulation build: tests only)";
#else
    return "bpe_b200 0.1 sm_100a";
#endif
}

const char* bpe_last_error(const bpe_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bpe_ctx_create(bpe_ctx** out, int device) {
    bpe_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, BPE_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, BPE_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(nullptr, BPE_ERR_INVALID_ARG, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, BPE_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 10) return fail(nullptr, BPE_ERR_CUDA, "device %d (%s) is sm_%d%d; this build needs sm_100", device, prop.name, prop.major, prop.minor);
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, BPE_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    ctx = new bpe_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if (cudaStreamCreate(&ctx->stream) != cudaSuccess) { delete ctx; return fail(nullptr, BPE_ERR_CUDA, "cudaStreamCreate failed"); }
    *out = ctx;
    return BPE_OK;
}

int bpe_nccl_unique_id(void* out128) { return dist_unique_id(out128, &g_create_err) ? BPE_OK : BPE_ERR_CUDA; }

int bpe_ctx_create_dist(bpe_ctx** out, int device, int rank, int world, const void* nccl_unique_id) {
    int rc = bpe_ctx_create(out, device);
    if (rc) return rc;
    if (world > 1) {
        std::string err;
        if (!(*out)->dist.init(rank, world, nccl_unique_id, (*out)->stream, &err)) {
            bpe_ctx_destroy(*out);
            *out = nullptr;
            return fail(nullptr, BPE_ERR_CUDA, "NCCL init failed: %s", err.c_str());
        }
#ifndef BPE_EMUL
        // NVLink peer mailboxes for the per-step exchange; if peer access is unavailable the NCCL all-reduce is used
        if (!(*out)->dist.init_peers(&err)) g_create_err = "peer exchange disabled: " + err;
#endif
    }
    return BPE_OK;
}

#ifdef BPE_EMUL
// emulation build only (tests): a "multi-GPU" context whose exchange is a caller-supplied callback
int bpe_ctx_create_dist_cb(bpe_ctx** out, int rank, int world, dist_allreduce_cb cb) {
    int rc = bpe_ctx_create(out, 0);
    if (rc) return rc;
    (*out)->dist.rank = rank;
    (*out)->dist.world = world;
    (*out)->dist.cb = cb;
    return BPE_OK;
}
// emulation build only (tests): the per-step exchange through peer mailboxes in a shared-memory segment, exactly the
// kernels the GPU build runs over NVLink peer memory (the last CTA of merge_kernel pushes, apply_kernel waits and sums)
int bpe_ctx_set_peer_shm(bpe_ctx* ctx, void* base, size_t bytes) {
    if (!ctx || !base) return BPE_ERR_INVALID_ARG;
    return ctx->dist.init_peers_shm(base, bytes) ? BPE_OK : BPE_ERR_INVALID_ARG;
}
size_t bpe_peer_shm_bytes(int world) { return (size_t)world * DistComm::mbox_words(world) * 4; }
#endif

void bpe_ctx_destroy(bpe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->dist.destroy_peers();
    ctx->dist.destroy();
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx->decode_cache;
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    ctx->cache.clear();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int bpe_ctx_set_option(bpe_ctx* ctx, const char* name, long value) {
    if (!ctx || !name) return BPE_ERR_INVALID_ARG;
    std::string s(name);
    if (s == "verify_recount") ctx->verify_recount = value;
    else if (s == "force_slow_tiebreak") ctx->force_slow_tiebreak = value;
    else if (s == "check_tiebreak") ctx->check_tiebreak = value;
    else if (s == "compact_pct") ctx->compact_pct = value;
    else if (s == "encode_compact_pct") ctx->encode_compact_pct = value;
    else if (s == "table_log2") ctx->table_log2 = value;
    else if (s == "max_steps") ctx->max_steps = value;
    else if (s == "time_phases") ctx->time_phases = value;
    else if (s == "profile") ctx->profile = value;
    else if (s == "debug") ctx->debug = value;
    else if (s == "merge_impl") ctx->merge_impl = value;
    else if (s == "xchg_impl") ctx->xchg_impl = value;
    else if (s == "encode_impl") ctx->encode_impl = value;
    else if (s == "encode_grid") ctx->encode_grid = value;
    else if (s == "encode_filter") ctx->encode_filter = value;
    else if (s == "encode_geom") ctx->encode_geom = value;
    else if (s == "encode_seg_min_steps") ctx->encode_seg_min_steps = value;
    else if (s == "encode_tile") ctx->encode_tile = value;
    else if (s == "encode_try_tiles") ctx->encode_try_tiles = value;
    else if (s == "fuse_halo") ctx->fuse_halo = value;
    else if (s == "count_limit_log2") ctx->count_limit_log2 = value;
    else if (s == "merge_filter") ctx->merge_filter = value;
    else if (s == "merge_direct") ctx->merge_direct = value;
    else if (s == "merge_prestage") ctx->merge_prestage = value;
    else if (s == "merge_pairfilter") ctx->merge_pairfilter = value;
    else if (s == "merge_loop") ctx->merge_loop = value;
    else if (s == "merge_dbuf") ctx->merge_dbuf = value;
    else if (s == "encode_fuse") ctx->encode_fuse = value;
    else if (s == "batch_steps") ctx->batch_steps = value;
    else if (s == "pdl") ctx->pdl = value;
    else if (s == "cache_max_mb") ctx->cache_max_mb = value;
    else if (s == "stream_chunk_mb") ctx->stream_chunk_mb = value;
    else if (s == "stream_chunk_bytes") ctx->stream_chunk_bytes = value;
    else return fail(ctx, BPE_ERR_INVALID_ARG, "unknown option '%s'", name);
    return BPE_OK;
}

int bpe_train_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, uint16_t vocab_size, bpe_merge_t* out_merges,
                     uint64_t* out_counts, size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    return train_device(ctx, d_text, n, vocab_size, out_merges, out_counts, out_n, stats);
}

int bpe_train(bpe_ctx* ctx, const uint8_t* text, size_t n, uint16_t vocab_size, bpe_merge_t* out_merges,
              uint64_t* out_counts, size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    if (n && !text) return fail(ctx, BPE_ERR_INVALID_ARG, "text is null");
    const double t0 = now_ms();
    CU(cudaSetDevice(ctx->device));
    DevBuf d;  // (through the context's cache: cudaMalloc + cudaFree of a GB cost ~0.2 s per call)
    if (d.alloc(n ? n : 1) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", n);
    CU(cudaMemcpyAsync(d.p, text, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = train_device(ctx, d.as<uint8_t>(), n, vocab_size, out_merges, out_counts, out_n, stats);
    if (stats) stats->total_ms = now_ms() - t0;
    return rc;
}

int bpe_encode_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m,
                      uint16_t* d_out, size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    return encode_device(ctx, d_text, n, merges, m, d_out, out_n, stats);
}

int bpe_encode(bpe_ctx* ctx, const uint8_t* text, size_t n, const bpe_merge_t* merges, size_t m, uint16_t* out,
               size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (n == 0 && ctx->dist.world == 1) { if (stats) memset(stats, 0, sizeof *stats); return BPE_OK; }
    if (n && (!text || !out)) return fail(ctx, BPE_ERR_INVALID_ARG, "text/out is null");
    const double t0 = now_ms();
    CU(cudaSetDevice(ctx->device));
    if (m && !merges) return fail(ctx, BPE_ERR_INVALID_ARG, "merges is null");
    if (ctx->encode_impl == 0 || ctx->encode_impl == 4) {
        // large inputs: chunks stream through the GPU (copies overlap the encoder, device memory stays bounded)
        bpe_stats_t sst;
        memset(&sst, 0, sizeof sst);
        const uint64_t l0 = ctx->launches;
        bool streamed = false;
        int rcs = encode_stream(ctx, text, n, merges, m, out, out_n, &sst, &streamed);
        if (rcs) return rcs;
        if (streamed) {
            sst.total_ms = now_ms() - t0;
            sst.device_ms = sst.kernel_ms[11];
            sst.kernel_launches = ctx->launches - l0;
            sst.replace_pair_calls = m;
            if (stats) *stats = sst;
            return BPE_OK;
        }
        *out_n = 0;
    }
    DevBuf d_in, d_out;
    if (d_in.alloc(n ? n : 1) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", n);
    if (d_out.alloc(n ? n * 2 : 2) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", n * 2);
    CU(cudaMemcpyAsync(d_in.p, text, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = encode_device(ctx, d_in.as<uint8_t>(), n, merges, m, d_out.as<uint16_t>(), out_n, stats);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, d_out.p, *out_n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) stats->total_ms = now_ms() - t0;
    return BPE_OK;
}

int bpe_decode_device(bpe_ctx* ctx, const uint16_t* d_toks, size_t n, const bpe_merge_t* merges, size_t m,
                      uint8_t* d_out, size_t cap, size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    if (n && !d_out) return fail(ctx, BPE_ERR_INVALID_ARG, "out is null");
    return decode_device(ctx, d_toks, n, merges, m, d_out, cap, out_n, stats);
}

int bpe_decode_size(bpe_ctx* ctx, const uint16_t* toks, size_t n, const bpe_merge_t* merges, size_t m, size_t* out_n) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (n == 0) return BPE_OK;
    if (!toks) return fail(ctx, BPE_ERR_INVALID_ARG, "toks is null");
    CU(cudaSetDevice(ctx->device));
    DevBuf d_in;
    if (d_in.alloc(n * 2) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", n * 2);
    CU(cudaMemcpyAsync(d_in.p, toks, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    return decode_device(ctx, d_in.as<uint16_t>(), n, merges, m, nullptr, 0, out_n, nullptr);
}

int bpe_decode(bpe_ctx* ctx, const uint16_t* toks, size_t n, const bpe_merge_t* merges, size_t m, uint8_t* out,
               size_t cap, size_t* out_n, bpe_stats_t* stats) {
    if (!ctx) return BPE_ERR_INVALID_ARG;
    CacheScope cache_scope(ctx);
    if (!out_n) return fail(ctx, BPE_ERR_INVALID_ARG, "out_n is null");
    *out_n = 0;
    if (n == 0) { if (stats) memset(stats, 0, sizeof *stats); return BPE_OK; }
    if (!toks || !out) return fail(ctx, BPE_ERR_INVALID_ARG, "toks/out is null");
    const double t0 = now_ms();
    CU(cudaSetDevice(ctx->device));
    DevBuf d_in, d_out;
    if (d_in.alloc(n * 2) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", n * 2);
    if (d_out.alloc(cap ? cap : 1) != cudaSuccess) return fail(ctx, BPE_ERR_OOM, "device allocation of %zu bytes failed", cap);
    CU(cudaMemcpyAsync(d_in.p, toks, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    int rc = decode_device(ctx, d_in.as<uint16_t>(), n, merges, m, d_out.as<uint8_t>(), cap, out_n, stats);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, d_out.p, *out_n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) stats->total_ms = now_ms() - t0;
    return BPE_OK;
}

}  // extern "C"
