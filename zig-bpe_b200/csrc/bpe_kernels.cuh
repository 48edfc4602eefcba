// bpe_kernels.cuh — device kernels of the B200 BPE engine (train / encode hot path).
//
// Reference functions replaced (all in /root/reference/src/basic_tokenizer.zig):
//   generateInitialTokens :155-170   -> widen_kernel
//   generateCodePointPairs :234-255  -> fused away (pairs are formed from the staged tile)
//   countCodePointPairs :257-278     -> byte_pair_hist_kernel + seed_table_kernel once, then
//                                       incremental deltas from the merge pass via apply_kernel
//   sortCodePointPairs :280-306, [0] :193 -> select_body (fused into apply_kernel) / select_kernel + the exact host replay
//   replaceTopPairWithNewToken :207-232   -> the merge pass: merge_loop_kernel in the train loop (looped CTAs, next tile
//                                       prefetched), merge_kernel elsewhere; halos gathered by apply_kernel / halo_kernel;
//                                       compaction kernels between batches
//   encode :71-88                    -> level_kernel (one pass per level of commuting merges), merge_kernel for
//                                       first == second pairs and irregular entries; bpe_tilenc.cuh / bpe_segenc.cuh
//   decode :90-138                   -> decode_len / tile_scan64 / decode_scatter
#pragma once
#include "bpe_common.cuh"

namespace bpe {

// =========================================================================================
// small block-level helpers (blockDim.x == THREADS unless stated)
// =========================================================================================
template <int N>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh /*[N]*/, uint32_t* total) {
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int off = 1; off < N; off <<= 1) {
        uint32_t add = (t >= off) ? sh[t - off] : 0u;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    uint32_t incl = sh[t];
    if (total) *total = sh[N - 1];
    __syncthreads();
    return incl - v;
}

// =========================================================================================
// load: u8 text -> token slots, padded with holes up to a whole number of tiles
// =========================================================================================
template <class TokT>
__global__ void widen_kernel(const uint8_t* __restrict__ text, size_t n, TokT* __restrict__ tok, size_t n_pad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_pad; i += stride) tok[i] = i < n ? (TokT)text[i] : (TokT)TokTraits<TokT>::hole;
}

template <class TokT>
__global__ void fill_holes_kernel(TokT* tok, size_t from, size_t to) {
    size_t i = from + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < to; i += stride) tok[i] = (TokT)TokTraits<TokT>::hole;
}

// =========================================================================================
// initial pair count over bytes: dense 256x256 histogram, hist[first | second << 8]
// next_byte: first byte of the following shard (multi-GPU), -1 if none
// =========================================================================================
// Block-private bins in shared memory: 65,536 u16 counters (128 KB, two per word), flushed to the
// global histogram after every pass of HIST_PASS bytes so that no counter can overflow. The global histogram is
// 64 bits wide (the reference counts in usize, :47,265; a 10 GB corpus can hold a pair more than 2^32 times, and the
// sum over the shards is taken in 64 bits too); the pair table keeps 32-bit counts, so seed_table_kernel refuses a
// histogram that does not fit — no later count can exceed the largest initial one (a new pair occurs at most as
// often as the pair it came from), so this one check covers the whole training.
constexpr int HIST_THREADS = 1024;
constexpr int HIST_BPT = 32;                       // bytes per thread per pass
constexpr int HIST_PASS = HIST_THREADS * HIST_BPT;  // 32,768 pairs per pass < 65,536
constexpr size_t HIST_SMEM = 65536 * 2;

__global__ void __launch_bounds__(HIST_THREADS, 1) byte_pair_hist_kernel(const uint8_t* __restrict__ text, size_t n,
                                                                         const EdgeInfo* __restrict__ edges, int rank, int world,
                                                                         unsigned long long* __restrict__ hist) {
    uint32_t* bins = bpe_dyn_smem();  // [32768] words
    int next_byte = -1;  // first byte of the following shards (multi-GPU), none otherwise
    if (edges) for (int q = rank + 1; q < world; q++) if (edges[q].nfirst) { next_byte = (int)edges[q].first[0]; break; }
    for (int i = (int)threadIdx.x; i < 32768; i += HIST_THREADS) bins[i] = 0;
    __syncthreads();
    const size_t npass = (n + HIST_PASS - 1) / HIST_PASS;
    const bool aligned = ((size_t)text & 15) == 0;
    for (size_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
        const size_t base = pass * HIST_PASS + (size_t)threadIdx.x * HIST_BPT;
        if (base < n) {
            uint8_t b[HIST_BPT + 1];
            const size_t avail = n - base;  // bytes from base to the end of the shard
            if (aligned && avail >= HIST_BPT) {
                const uint4* p = reinterpret_cast<const uint4*>(text + base);
                uint4 v0 = p[0], v1 = p[1];
                uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    b[4 * k + 0] = (uint8_t)(w[k]); b[4 * k + 1] = (uint8_t)(w[k] >> 8);
                    b[4 * k + 2] = (uint8_t)(w[k] >> 16); b[4 * k + 3] = (uint8_t)(w[k] >> 24);
                }
            } else {
                for (int k = 0; k < HIST_BPT; k++) b[k] = (size_t)k < avail ? text[base + k] : 0;
            }
            int cnt = avail >= HIST_BPT ? HIST_BPT : (int)avail;  // left tokens handled by this thread
            int last = -1;                                        // successor of my last byte
            if (avail > (size_t)cnt) last = text[base + cnt];
            else last = next_byte;                                // end of shard: first byte of the next shard, or none
            for (int k = 0; k < cnt; k++) {
                int nb = (k + 1 < cnt) ? (int)b[k + 1] : last;
                if (nb < 0) break;
                uint32_t bin = (uint32_t)b[k] | ((uint32_t)nb << 8);
                atomicAdd(&bins[bin >> 1], 1u << ((bin & 1u) * 16u));
            }
        }
        __syncthreads();
        for (int i = (int)threadIdx.x; i < 32768; i += HIST_THREADS) {
            uint32_t wv = bins[i];
            if (wv) {
                bins[i] = 0;
                if (wv & 0xFFFFu) atomicAdd(&hist[2 * i], (unsigned long long)(wv & 0xFFFFu));
                if (wv >> 16) atomicAdd(&hist[2 * i + 1], (unsigned long long)(wv >> 16));
            }
        }
        __syncthreads();
    }
}

// hist (already summed over GPUs) -> pair table + reference-home population
__global__ void seed_table_kernel(const unsigned long long* __restrict__ hist, PairTable tbl, StepCtl* ctl, unsigned long long count_limit) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536u) return;
    const unsigned long long c64 = hist[i];
    if (!c64) return;
    if (c64 > count_limit) { atomicOr(&ctl->err, (uint32_t)ERR_COUNT_OVERFLOW); return; }
    const uint32_t c = (uint32_t)c64;
    uint32_t key = pair_key(i & 255u, i >> 8);
    uint32_t s = tbl_find_or_insert(tbl, key, &ctl->n_inserted);
    if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_TABLE_FULL); return; }
    atomicAdd(&tbl.counts[s], c);
    atomicAdd(&ctl->live_keys, 1u);
}

// =========================================================================================
// argmax over the pair table, then the list of keys that share the maximum
// =========================================================================================
__global__ void argmax_kernel(PairTable tbl, StepCtl* ctl) {
    __shared__ uint32_t sh[THREADS];
    uint32_t m = 0;
    size_t cap = (size_t)tbl.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t c = tbl.counts[i];
        m = c > m ? c : m;
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int off = THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { uint32_t o = sh[threadIdx.x + off]; if (o > sh[threadIdx.x]) sh[threadIdx.x] = o; }
        __syncthreads();
    }
    if (threadIdx.x == 0 && sh[0]) atomicMax(&ctl->max_count, sh[0]);
}

__global__ void ties_kernel(PairTable tbl, StepCtl* ctl) {
    const uint32_t m = ctl->max_count;
    if (m == 0) return;
    size_t cap = (size_t)tbl.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        if (tbl.counts[i] == m) {
            uint32_t idx = atomicAdd(&ctl->ntied, 1u);
            if (idx < (uint32_t)MAXTIE) ctl->tie_keys[idx] = tbl.keys[i];
        }
    }
}

// -----------------------------------------------------------------------------------------
// Heavy-key list. The maximum pair count never increases (old pairs only lose occurrences and a
// new pair (p,X)/(X,n)/(X,X) cannot occur more often than the (A,B) it came from), so it is
// enough to scan the keys whose count was >= theta when the list was built, plus keys created
// later with a count >= theta (appended by the apply kernel). The list is exact while
// max >= theta; the host rebuilds it with a lower theta when the maximum falls below.
// -----------------------------------------------------------------------------------------
struct HeavyList { uint32_t* slots; uint32_t cap; uint32_t theta; };

__global__ void heavy_collect_kernel(PairTable tbl, HeavyList hl, StepCtl* ctl) {
    size_t cap = (size_t)tbl.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        if (tbl.counts[i] >= hl.theta) {
            uint32_t idx = atomicAdd(&ctl->n_heavy, 1u);
            if (idx < hl.cap) hl.slots[idx] = (uint32_t)i;
        }
    }
}

__global__ void heavy_argmax_kernel(PairTable tbl, HeavyList hl, StepCtl* ctl) {
    __shared__ uint32_t sh[THREADS];
    uint32_t n = ctl->n_heavy < hl.cap ? ctl->n_heavy : hl.cap;
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t c = tbl.counts[hl.slots[i]];
        m = c > m ? c : m;
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int off = THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { uint32_t o = sh[threadIdx.x + off]; if (o > sh[threadIdx.x]) sh[threadIdx.x] = o; }
        __syncthreads();
    }
    if (threadIdx.x == 0 && sh[0]) atomicMax(&ctl->max_count, sh[0]);
}

// single CTA: maximum and tie list in one launch (the heavy list is short)
__global__ void heavy_argmax_ties_kernel(PairTable tbl, HeavyList hl, StepCtl* ctl) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t s_ntied;
    const uint32_t n = ctl->n_heavy < hl.cap ? ctl->n_heavy : hl.cap;
    uint32_t m = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t c = tbl.counts[hl.slots[i]];
        m = c > m ? c : m;
    }
    sh[threadIdx.x] = m;
    if (threadIdx.x == 0) s_ntied = 0;
    __syncthreads();
    for (int off = (int)blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { uint32_t o = sh[threadIdx.x + off]; if (o > sh[threadIdx.x]) sh[threadIdx.x] = o; }
        __syncthreads();
    }
    m = sh[0];
    if (m) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t s = hl.slots[i];
            if (tbl.counts[s] == m) {
                uint32_t idx = atomicAdd(&s_ntied, 1u);
                if (idx < (uint32_t)MAXTIE) ctl->tie_keys[idx] = tbl.keys[s];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { ctl->max_count = m; ctl->ntied = s_ntied; ctl->tie_status = TIE_NONE; }
}

// record the decision for the current step
__device__ __forceinline__ void commit_merge(StepCtl* ctl, MergeRec* rec, uint32_t key, uint32_t count) {
    ctl->A = key & 0xFFFFu;
    ctl->B = key >> 16;
    ctl->X = 256u + ctl->step;
    rec[ctl->step].key = key;
    rec[ctl->step].count = count;
    // first == second: the merge pass needs run lengths across tiles, which the fused halo does not produce
    if ((ctl->flags & F_HALT_AEQB) && (key & 0xFFFFu) == (key >> 16)) ctl->halt = H_AEQB;
}

// ---- tie fast path (device) -------------------------------------------------------------
// The reference picks, among the pairs sharing the maximum count, the one in the lowest slot of its
// hash table (Appendix A). For linear probing without deletions the *set* of occupied slots does not
// depend on insertion order: with cnt[x] = keys whose home is x, the carry recurrence
//     o[x] = max(0, o[x-1] + cnt[x] - 1)
// says exactly which slots are free (slot x is free iff o[x-1] + cnt[x] == 0). If the tied key with
// the smallest home h1 reaches a free slot before the next tied home, and no tied key's probe run
// wraps past the end of the table, that key owns the lowest slot. Anything else goes to the replay.
//
// The carry entering a slot is computed locally: walking back over chunks of ZCHUNK slots, the carry
// a chunk can pass on is at most its population, so once  ZCHUNK + sum(pop - ZCHUNK)  over the
// chunks walked is <= 0 nothing further left can reach us (all chunk populations are <= ZCHUNK,
// checked through zpop_max), and the recurrence restarted with o = 0 at that point is exact.
struct ChunkFn { int32_t add; int32_t lo; };  // f(o) = max(lo, o + add)
__device__ __forceinline__ ChunkFn fn_compose(ChunkFn f, ChunkFn g) {  // g after f
    ChunkFn r;
    r.add = f.add + g.add;
    int32_t a = f.lo + g.add;
    r.lo = g.lo > a ? g.lo : a;
    return r;
}
__device__ __forceinline__ int32_t fn_eval(ChunkFn f, int32_t o) { int32_t a = o + f.add; return f.lo > a ? f.lo : a; }

constexpr uint32_t TIE_MAX_BACK = 64;  // chunks walked back before giving up (-> replay)

// one warp: first free slot at/after the home of `key`. returns false if it cannot be decided locally
__device__ __forceinline__ bool zig_first_free(const ZigPop& z, uint32_t zcap, uint32_t key, uint32_t* home_out,
                                               uint32_t* free_out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t zchunk = zcap < ZCHUNK ? zcap : ZCHUNK;
    const uint32_t nchunks = zcap / zchunk;
    const uint32_t h = (uint32_t)zig_hash_pair(key) & (zcap - 1);
    const uint32_t c = h / zchunk;
    // how many whole chunks to walk back (warp-uniform: every lane computes the same)
    uint32_t K = 0;
    bool full_lap = false;
    {
        int64_t P = 0;
        while (true) {
            K++;
            if (K >= nchunks) { full_lap = true; break; }  // small table: take a whole lap (plus one more below)
            if (K > TIE_MAX_BACK) return false;
            const uint32_t cc = (c + nchunks - K) % nchunks;
            P += (int64_t)z.zpop[cc] - (int64_t)zchunk;
            if ((int64_t)zchunk + P <= 0) break;
        }
    }
    // slots to evaluate before h: K chunks plus the part of chunk c before h; a full lap is taken
    // twice (the first lap delivers the steady-state carry, load < 1 makes it a fixed point)
    uint64_t len;
    uint32_t start;
    if (full_lap) { len = (uint64_t)2 * zcap; start = h; }  // start at h, go around twice, end at h
    else { len = (uint64_t)K * zchunk + (h - c * zchunk); start = ((c + nchunks - K) % nchunks) * zchunk; }
    // each lane folds a contiguous part, then an ordered shuffle reduction composes them
    const uint64_t per = (len + 31) / 32;
    uint64_t lo = per * lane, hi = lo + per;
    if (lo > len) lo = len;
    if (hi > len) hi = len;
    ChunkFn f; f.add = 0; f.lo = 0;
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t x = (uint32_t)((start + i) & (uint64_t)(zcap - 1));
        ChunkFn g; g.add = (int32_t)zcnt_get(z.zcnt, x) - 1; g.lo = 0;
        f = fn_compose(f, g);
    }
    for (int off = 1; off < 32; off <<= 1) {
        ChunkFn g;
        g.add = (int32_t)__shfl_down_sync(0xffffffffu, (uint32_t)f.add, (unsigned)off);
        g.lo = (int32_t)__shfl_down_sync(0xffffffffu, (uint32_t)f.lo, (unsigned)off);
        if (lane + (uint32_t)off < 32u) f = fn_compose(f, g);
    }
    int32_t o = fn_eval(f, 0);  // lane 0: carry entering slot h
    o = (int32_t)__shfl_sync(0xffffffffu, (uint32_t)o, 0);
    uint32_t x = h;
    bool ok = true;
    while (true) {
        int32_t occ = o + (int32_t)zcnt_get(z.zcnt, x);
        if (occ == 0) break;  // free slot
        o = occ - 1;
        x++;
        if (x == zcap) { ok = false; break; }  // the run wraps: order is not decided by homes alone
    }
    *home_out = h;
    *free_out = x;
    return ok;
}

// select_kernel (single CTA): the device-driven replacement of "sort, take [0]" (:186-193). Scans
// the heavy list for the maximum and its ties, then commits the merge (unique maximum, or a tie
// the occupancy test settles) or halts the loop for the host.
__device__ __forceinline__ void select_body(const PairTable& tbl, HeavyList hl, StepCtl* ctl, MergeRec* rec, ZigPop z) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t s_ntied;
    __shared__ uint32_t s_mode;  // 0 done, 1 settle the tie here
    __shared__ uint32_t s_home[MAXTIE];
    __shared__ uint32_t s_free[MAXTIE];
    __shared__ uint32_t s_bad;
    if (ctl->halt) return;
    if (ctl->step >= ctl->want_steps) {
        if (threadIdx.x == 0) ctl->halt = H_DONE;
        return;
    }
    const uint32_t nthr = blockDim.x;
    hl.theta = ctl->theta;
    const uint32_t n = ctl->n_heavy < hl.cap ? ctl->n_heavy : hl.cap;
    uint32_t m = 0;
    constexpr int SC = 16;  // list entries kept in registers per thread (covers lists up to 16K keys)
    uint32_t cs[SC], cc[SC];
#pragma unroll
    for (int k = 0; k < SC; k++) {
        const uint32_t i = threadIdx.x + (uint32_t)k * nthr;
        cs[k] = i < n ? hl.slots[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SC; k++) {
        const uint32_t i = threadIdx.x + (uint32_t)k * nthr;
        cc[k] = i < n ? tbl.counts[cs[k]] : 0u;
        m = cc[k] > m ? cc[k] : m;
    }
    for (uint32_t i = threadIdx.x + (uint32_t)SC * nthr; i < n; i += nthr) {
        uint32_t c = tbl.counts[hl.slots[i]];
        m = c > m ? c : m;
    }
    sh[threadIdx.x] = m;
    if (threadIdx.x == 0) { s_ntied = 0; s_mode = 0; s_bad = 0; }
    __syncthreads();
    for (int off = (int)blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { uint32_t o = sh[threadIdx.x + off]; if (o > sh[threadIdx.x]) sh[threadIdx.x] = o; }
        __syncthreads();
    }
    m = sh[0];
    const bool list_ok = (ctl->theta != 0) && (ctl->n_heavy <= hl.cap) && (m >= ctl->theta);
    if (list_ok) {
#pragma unroll
        for (int k = 0; k < SC; k++) {
            if (cc[k] == m && threadIdx.x + (uint32_t)k * nthr < n) {
                uint32_t idx = atomicAdd(&s_ntied, 1u);
                if (idx < (uint32_t)MAXTIE) ctl->tie_keys[idx] = tbl.keys[cs[k]];
            }
        }
        for (uint32_t i = threadIdx.x + (uint32_t)SC * nthr; i < n; i += nthr) {
            uint32_t s = hl.slots[i];
            if (tbl.counts[s] == m) {
                uint32_t idx = atomicAdd(&s_ntied, 1u);
                if (idx < (uint32_t)MAXTIE) ctl->tie_keys[idx] = tbl.keys[s];
            }
        }
    }
    __syncthreads();
    const uint32_t ntied = s_ntied;
    if (threadIdx.x == 0) {
        ctl->max_count = m;
        ctl->ntied = ntied;
        ctl->tie_status = TIE_NONE;
        ctl->need_tie = 0;
        if (!list_ok) ctl->halt = H_HEAVY;
        else if (ntied == 1) commit_merge(ctl, rec, ctl->tie_keys[0], m);
        else {
            const uint32_t D = ctl->live_keys;
            const uint32_t zc = zig_cap_for(D);
            if ((ctl->flags & F_FORCE_REPLAY) || ntied > (uint32_t)MAXTIE || D == zig_max_load(zc)) ctl->halt = H_REPLAY;
            else if (zc != ctl->zcap) ctl->halt = H_ZCAP;
            else if (ctl->zpop_max > (zc < ZCHUNK ? zc : ZCHUNK)) ctl->halt = H_REPLAY;  // a chunk could overflow into the next
            else s_mode = 1;
        }
    }
    __syncthreads();
    if (s_mode != 1) return;
    // settle the tie: one warp per tied key
    const uint32_t zcap = ctl->zcap;
    z.zmask = zcap - 1;
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t i = warp; i < ntied; i += nwarps) {
        uint32_t hm = 0, fr = 0;
        bool ok = zig_first_free(z, zcap, ctl->tie_keys[i], &hm, &fr);
        if ((threadIdx.x & 31u) == 0) {
            s_home[i] = hm;
            s_free[i] = fr;
            if (!ok) atomicOr(&s_bad, 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t status = TIE_NEED_REPLAY, winner = 0;
    if (!s_bad) {
        uint32_t b = 0;
        for (uint32_t i = 1; i < ntied; i++) if (s_home[i] < s_home[b]) b = i;
        bool ok = true;
        for (uint32_t i = 0; i < ntied; i++) if (i != b && s_home[i] < s_free[b]) { ok = false; break; }
        if (ok) { status = TIE_FAST_OK; winner = ctl->tie_keys[b]; }
    }
    ctl->tie_status = status;
    ctl->tie_winner = winner;
    if (status == TIE_FAST_OK && !(ctl->flags & F_CHECK_TIES)) { commit_merge(ctl, rec, winner, m); ctl->fast_ties += 1; }
    else ctl->halt = H_REPLAY;
}

__global__ void __launch_bounds__(1024) select_kernel(PairTable tbl, HeavyList hl, StepCtl* ctl, MergeRec* rec, ZigPop z) {
    select_body(tbl, hl, ctl, rec, z);
}

__global__ void heavy_ties_kernel(PairTable tbl, HeavyList hl, StepCtl* ctl) {
    const uint32_t m = ctl->max_count;
    if (m == 0) return;
    uint32_t n = ctl->n_heavy < hl.cap ? ctl->n_heavy : hl.cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t s = hl.slots[i];
        if (tbl.counts[s] == m) {
            uint32_t idx = atomicAdd(&ctl->ntied, 1u);
            if (idx < (uint32_t)MAXTIE) ctl->tie_keys[idx] = tbl.keys[s];
        }
    }
}

// =========================================================================================
// halo_kernel: one thread per tile gathers the live tokens around the tile.
// ext_*: live tokens beyond this GPU's shard (hole = none). For A==B steps it also measures
// the run of A that ends at the tile's left edge (run_local / run_full, chained by
// run_chain_kernel so the cost stays linear on degenerate inputs such as "aaaa...").
// =========================================================================================
// FROMCTL: the merge is read from the device control block (device-driven train loop) and the
// kernel does nothing while the loop is halted; otherwise A / aeqb come from the arguments.
// For A==B steps the last CTA to finish chains the per-tile run lengths:
//   runA[t] = run_local[t] + (run_full[t] ? runA[t-1] : 0), runA[0] = ext_run
// (each thread folds a contiguous range of tiles into (add, full), thread 0 chains the ranges, then
// every thread replays its range from its carry-in — linear work even on "aaaa..." inputs).
// zero2 (nullable): the two scalar merge deltas (cntXX, cntAB) to clear before the merge pass.
constexpr int HALO_THREADS = 128;

// Where a shard's neighbours' end tokens come from: the local copy of the exchanged EdgeInfo slots (single GPU:
// none; NCCL path and host-driven passes: the exchange buffer) or, inside the fused apply + halo kernel of the
// peer-memory path, the mailbox slots themselves (valid once the sender's arrival flag shows this step's epoch).
// Mailbox cells are 8 bytes: {value, epoch}. One 8-byte store carries the value together with the tag that says which
// step it belongs to, so a reader needs neither a separate arrival flag nor a fence between data and flag (the
// "flag in data" scheme of low-latency collectives): it polls the cell until the tag is the step's epoch. Epochs grow
// monotonically over steps and trainings, so a stale cell can never be mistaken for a fresh one.
#ifdef BPE_EMUL
constexpr uint32_t PEER_SPIN_LIMIT = 0xFFFFFFF0u;  // (the emulated peers are processes that take seconds per step)
#else
constexpr uint32_t PEER_SPIN_LIMIT = 1u << 24;
#endif
__device__ __forceinline__ void peer_put(uint32_t* mbox, size_t cell, uint32_t value, uint32_t epoch) {
    __stcg(reinterpret_cast<uint2*>(mbox) + cell, make_uint2(value, epoch));
}
__device__ __forceinline__ uint32_t peer_get(const uint32_t* mbox, size_t cell, uint32_t epoch, uint32_t* err) {
    uint32_t spins = 0;
    while (true) {
        const uint2 v = __ldcv(reinterpret_cast<const uint2*>(mbox) + cell);
        if (v.y == epoch) return v.x;
        BPE_SPIN_YIELD();
        if (++spins > PEER_SPIN_LIMIT) { atomicOr(err, (uint32_t)ERR_PEER_TIMEOUT); return 0u; }  // (seconds: a peer is gone)
    }
}
struct EdgeSrc {
    const EdgeInfo* local;      // nullptr: single GPU
    const uint32_t* mbox;       // != nullptr: read rank r's EdgeInfo from the tagged cells r * slot_words + edge_off + r * 16 ...
    const uint32_t* flags;      // (unused)
    uint32_t slot_words, edge_off, epoch;
    uint32_t* err;
};
__device__ __forceinline__ EdgeInfo edge_get(const EdgeSrc& es, int r) {
    if (!es.mbox) return es.local[r];
    EdgeInfo e;
    const size_t c0 = (size_t)r * es.slot_words + es.edge_off + (uint32_t)r * 16u;
    uint32_t* dst = reinterpret_cast<uint32_t*>(&e);
    // all 16 cells are requested before any tag is looked at (one memory latency instead of sixteen); a cell that does
    // not carry the epoch yet is polled on its own
    const uint2* cells = reinterpret_cast<const uint2*>(es.mbox) + c0;
    uint2 v[16];
#pragma unroll
    for (int w = 0; w < 16; w++) v[w] = __ldcv(cells + w);
#pragma unroll
    for (int w = 0; w < 16; w++) dst[w] = v[w].y == es.epoch ? v[w].x : peer_get(es.mbox, c0 + (size_t)w, es.epoch, es.err);
    return e;
}

// the live tokens around tile t (holes are skipped, the search continues in the neighbouring shards)
template <class TokT>
__device__ __forceinline__ void halo_gather(const TokT* __restrict__ tok, size_t n_slots, uint32_t t, TileHalo<TokT>* halo,
                                            const EdgeSrc& es, int rank, int world) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    const bool have_edges = es.local != nullptr || es.mbox != nullptr;
    TileHalo<TokT> h;
    TokT l[2] = {H, H};
    int nl = 0;
    for (size_t i = (size_t)t * TILE; i > 0 && nl < 2;) {
        --i;
        TokT v = tok[i];
        if (v != H) l[nl++] = v;
    }
    if (nl < 2 && have_edges) {  // fell off the shard: continue in the shards before this one
        for (int r = rank - 1; r >= 0 && nl < 2; r--) {
            const EdgeInfo e = edge_get(es, r);
            for (uint32_t k = 0; k < e.nlast && nl < 2; k++) l[nl++] = (TokT)e.last[k];
        }
    }
    h.l1 = l[0];
    h.l2 = (l[0] == H) ? H : l[1];
    TokT r3[3] = {H, H, H};
    int nr = 0;
    for (size_t i = (size_t)(t + 1) * TILE; i < n_slots && nr < 3; i++) {
        TokT v = tok[i];
        if (v != H) r3[nr++] = v;
    }
    if (nr < 3 && have_edges) {  // continue in the shards after this one
        for (int q = rank + 1; q < world && nr < 3; q++) {
            const EdgeInfo e = edge_get(es, q);
            for (uint32_t k = 0; k < e.nfirst && nr < 3; k++) r3[nr++] = (TokT)e.first[k];
        }
    }
    h.r0 = r3[0]; h.r1 = r3[1]; h.r2 = r3[2];
    h.runA = 0;
    halo[t] = h;
}

template <class TokT, bool FROMCTL>
__global__ void __launch_bounds__(HALO_THREADS) halo_kernel(const TokT* __restrict__ tok, size_t n_slots, uint32_t ntiles,
                            TileHalo<TokT>* halo, const StepCtl* __restrict__ ctl, uint32_t Au, int aeqb,
                            uint32_t* run_local, uint8_t* run_full, uint32_t* done_counter, uint32_t* zero2,
                            const EdgeInfo* __restrict__ edges, int rank, int world) {
    __shared__ uint32_t s_add[HALO_THREADS];
    __shared__ uint32_t s_full[HALO_THREADS];
    __shared__ uint32_t s_last;
    const TokT H = (TokT)TokTraits<TokT>::hole;
    if (FROMCTL) {
        if (ctl->halt) return;
        Au = ctl->A;
        aeqb = (ctl->A == ctl->B) ? 1 : 0;
    }
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0 && zero2) { zero2[0] = 0; zero2[1] = 0; }
    if (t < ntiles) {
        EdgeSrc es;
        es.local = edges; es.mbox = nullptr; es.flags = nullptr; es.slot_words = es.edge_off = es.epoch = 0; es.err = nullptr;
        halo_gather<TokT>(tok, n_slots, t, halo, es, rank, world);
        if (aeqb) {
            // live A's walking back from the tile's left edge, inside the previous tile only
            const TokT A = (TokT)Au;
            uint32_t cnt = 0;
            uint8_t full = 1;
            if (t == 0) { full = 0; }
            else {
                size_t lo = (size_t)(t - 1) * TILE;
                for (size_t i = (size_t)t * TILE; i > lo;) {
                    --i;
                    TokT v = tok[i];
                    if (v == H) continue;
                    if (v == A) cnt++; else { full = 0; break; }
                }
            }
            run_local[t] = cnt;
            run_full[t] = full;
        }
    }
    if (!aeqb) return;
    // ---- the last CTA to get here chains the runs ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t prev = atomicAdd(done_counter, 1u);
        s_last = (prev == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t nthr = blockDim.x;
    const uint32_t q = (ntiles + nthr - 1) / nthr;
    const uint32_t lo = threadIdx.x * q;
    uint32_t hi = lo + q;
    if (hi > ntiles) hi = ntiles;
    uint32_t add = 0, full = 1;
    for (uint32_t i = (lo == 0 ? 1u : lo); i < hi; i++) {
        const uint32_t rl = ((volatile uint32_t*)run_local)[i];
        if (((volatile uint8_t*)run_full)[i]) add += rl;
        else { add = rl; full = 0; }
    }
    s_add[threadIdx.x] = add;
    s_full[threadIdx.x] = full;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t carry = 0;  // run of A entering tile 0: it may continue through the shards before this one
        if (edges) {
            for (int r = rank - 1; r >= 0; r--) {
                if (edges[r].nlast == 0) continue;           // empty shard
                if (edges[r].run_tok != Au) break;
                carry += edges[r].run_len;
                if (!edges[r].all_same) break;
            }
        }
        for (uint32_t i = 0; i < nthr; i++) {
            uint32_t a = s_add[i], f = s_full[i];
            s_add[i] = carry;
            carry = a + (f ? carry : 0u);
        }
        *done_counter = 0;  // ready for the next launch
    }
    __syncthreads();
    uint32_t run = s_add[threadIdx.x];
    if (lo == 0 && lo < hi) halo[0].runA = run;
    for (uint32_t i = (lo == 0 ? 1u : lo); i < hi; i++) {
        run = ((volatile uint32_t*)run_local)[i] + (((volatile uint8_t*)run_full)[i] ? run : 0u);
        halo[i].runA = run;
    }
}

// =========================================================================================
// edge_kernel (multi-GPU, one warp): describe this shard's ends in its EdgeInfo slot of the exchange
// buffer and zero the other ranks' slots (the buffer is summed by the all-reduce). It also keeps the
// shard's own live-token count, because after the all-reduce only the global merged count is left.
// =========================================================================================
// one warp; zero_others: clear the other ranks' slots (all-reduce path: the buffer is summed)
template <class TokT>
__device__ __forceinline__ void edge_body(const TokT* __restrict__ tok, size_t n_slots, size_t tail_hint, EdgeInfo* all, int rank,
                                          int world, StepCtl* ctl, const uint32_t* nab_local, int account, bool zero_others) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    const uint32_t lane = threadIdx.x & 31u;
    if (zero_others) {
        uint32_t* words = reinterpret_cast<uint32_t*>(all);
        for (uint32_t w = lane; w < (uint32_t)world * 16u; w += 32u)
            if ((int)(w / 16u) != rank) words[w] = 0;
    }
    EdgeInfo* e = &all[rank];
    // ---- first three live tokens: the warp scans 32 slots at a time ----
    uint32_t nf = 0, fv[3] = {0, 0, 0};
    for (size_t base = 0; base < n_slots && nf < 3; base += 32) {
        const size_t i = base + lane;
        const uint32_t v = i < n_slots ? (uint32_t)tok[i] : (uint32_t)H;
        uint32_t live = __ballot_sync(0xffffffffu, v != (uint32_t)H);
        while (live && nf < 3) {
            const int src = __ffs((int)live) - 1;
            live &= live - 1;
            fv[nf++] = __shfl_sync(0xffffffffu, v, src);
        }
    }
    // ---- last two live tokens and the run of equal tokens that ends the shard ----
    uint32_t n = 0, run = 0, all_same = 1, lastv = 0, lasts[2] = {0, 0};
    bool stop = false;
    // everything at or beyond tail_hint is padding (holes written by the load / the last compaction)
    for (size_t top = tail_hint < n_slots ? tail_hint : n_slots; top > 0 && !stop;) {
        const size_t base = top >= 32 ? top - 32 : 0;
        const size_t i = base + lane;
        const uint32_t v = i < top ? (uint32_t)tok[i] : (uint32_t)H;
        uint32_t live = __ballot_sync(0xffffffffu, v != (uint32_t)H);
        // fast path for long runs: every live token of this chunk continues the run
        if (live && n >= 2 && run == n) {
            const uint32_t same = __ballot_sync(0xffffffffu, v == lastv);
            if (same == live) { const uint32_t c = (uint32_t)__popc(live); n += c; run += c; top = base; continue; }
        }
        while (live) {
            const int src = 31 - __clz((int)live);
            live &= ~(1u << src);
            const uint32_t tv = __shfl_sync(0xffffffffu, v, src);
            if (n < 2) lasts[n] = tv;
            if (n == 0) lastv = tv;
            n++;
            if (tv == lastv && run + 1 == n) run++;
            else { all_same = 0; if (n >= 2) { stop = true; break; } }
        }
        top = base;
    }
    if (n == 0) all_same = 0;
    if (lane == 0) {
        for (uint32_t k = 0; k < 3; k++) e->first[k] = fv[k];
        e->nfirst = nf;
        e->last[0] = lasts[0]; e->last[1] = lasts[1];
        e->nlast = n < 2 ? n : 2;
        e->run_tok = n ? lastv : 0u;
        e->run_len = run;
        e->all_same = all_same;
        for (int k = 0; k < 6; k++) e->pad[k] = 0;
        if (account && !ctl->halt) ctl->local_live -= *nab_local;
    }
}

template <class TokT>
__global__ void edge_kernel(const TokT* __restrict__ tok, size_t n_slots, size_t tail_hint, EdgeInfo* all, int rank, int world,
                            StepCtl* ctl, const uint32_t* nab_local, int account) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    edge_body<TokT>(tok, n_slots, tail_hint, all, rank, world, ctl, nab_local, account, true);
}

// Peer-memory exchange (multi-GPU), inside the apply kernel. Every apply thread owns one delta cell (the one it folds
// into the table). It stores this rank's value of the cell, tagged with the step's epoch, into slot `rank` of every
// peer's mailbox over NVLink (CTA 0 also describes the shard's ends and sends its EdgeInfo the same way), then reads
// the cell as local value + the peers' values, polling each peer cell until it carries the epoch. A thread depends
// only on the same thread of the other ranks — never on another CTA of its own grid, no flags, no fences — so
// transfer, reduction and table update are one kernel and the exchange costs one NVLink store latency.
struct PushArgs {
    size_t n_slots, tail_hint;  // sequence geometry for the shard-end scan
    const uint16_t* tok;
};

// =========================================================================================
// tile staging: ext[] = [holes | l2 l1 | TILE slots | r0 r1 r2 | holes], tile data at OFF
// =========================================================================================
constexpr int OFF = 8;
constexpr int EXT = TILE + 16;
constexpr int EXT_END = OFF + TILE + 3;  // one past the last index that may hold a token

template <class TokT>
__device__ __forceinline__ void stage_tile(TokT* ext, const TokT* __restrict__ tok, size_t base,
                                           const TileHalo<TokT>& h) {
    constexpr int VEC = 16 / sizeof(TokT);
    const uint4* src = reinterpret_cast<const uint4*>(tok + base);
    uint4* dst = reinterpret_cast<uint4*>(ext + OFF);
    for (int i = threadIdx.x; i < TILE / VEC; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) {
        const TokT H = (TokT)TokTraits<TokT>::hole;
        for (int i = 0; i < OFF - 2; i++) ext[i] = H;
        ext[OFF - 2] = h.l2;
        ext[OFF - 1] = h.l1;
        ext[OFF + TILE + 0] = h.r0;
        ext[OFF + TILE + 1] = h.r1;
        ext[OFF + TILE + 2] = h.r2;
        for (int i = OFF + TILE + 3; i < EXT; i++) ext[i] = H;
    }
}
template <class TokT> __device__ __forceinline__ int next_live(const TokT* ext, int i) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    for (int j = i + 1; j < EXT_END; j++) if (ext[j] != H) return j;
    return -1;
}
template <class TokT> __device__ __forceinline__ int prev_live(const TokT* ext, int i) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    for (int j = i - 1; j >= OFF - 2; j--) if (ext[j] != H) return j;
    return -1;
}

// =========================================================================================
// merge_kernel: replace every occurrence of (A,B) by X, left to right, non-overlapping, exactly
// as replaceTopPairWithNewToken (:207-232) / the encode pass (:75-85) do. One CTA per tile.
// Each thread writes only its own slots: the slot holding A becomes X, the slot holding the
// consumed B becomes a hole. With DELTAS it also emits the count changes of the neighbouring
// pairs: cntL[p] merges had unmerged left neighbour p ((p,A)-1,(p,X)+1), cntR[n] had right
// neighbour n that does not start another occurrence ((B,n)-1,(X,n)+1), cntXX adjacent
// occurrences ((B,A)-1,(X,X)+1), cntAB occurrences ((A,B)-1 each).
// =========================================================================================
// bit i of the result is set when slot i of the 16-byte vector equals `a`
template <class TokT> __device__ __forceinline__ uint32_t vec_mask(const uint4& v, uint32_t a);
template <> __device__ __forceinline__ uint32_t vec_mask<uint16_t>(const uint4& v, uint32_t a) {
    // bit (i + 16*half) <-> slot 2*i + half: one AND per word instead of a per-slot shuffle of bits
    // exact zero-halfword test without cross-lane carries: z = ~(((x & 0x7FFF7FFF) + 0x7FFF7FFF) | x) & 0x80008000
    // (sm_100 has no SIMD video compare: __vcmpeq2 expands to ~6 instructions per word)
    const uint32_t aa = a | (a << 16);
    auto z = [&](uint32_t w) { const uint32_t x = w ^ aa; return ~(((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) & 0x80008000u; };
    return (z(v.x) >> 15) | (z(v.y) >> 14) | (z(v.z) >> 13) | (z(v.w) >> 12);
}
// slot index of a set bit of vec_mask
template <class TokT> __device__ __forceinline__ int mask_bit_to_slot(int bit);
template <> __device__ __forceinline__ int mask_bit_to_slot<uint16_t>(int bit) { return ((bit & 15) << 1) | (bit >> 4); }
template <> __device__ __forceinline__ int mask_bit_to_slot<uint32_t>(int bit) { return bit; }
template <> __device__ __forceinline__ uint32_t vec_mask<uint32_t>(const uint4& v, uint32_t a) {
    return (v.x == a ? 1u : 0u) | (v.y == a ? 2u : 0u) | (v.z == a ? 4u : 0u) | (v.w == a ? 8u : 0u);
}
template <class TokT> __device__ __forceinline__ void unpack_vec(const uint4& v, uint32_t* out);
template <> __device__ __forceinline__ void unpack_vec<uint16_t>(const uint4& v, uint32_t* out) {
    out[0] = v.x & 0xFFFFu; out[1] = v.x >> 16; out[2] = v.y & 0xFFFFu; out[3] = v.y >> 16;
    out[4] = v.z & 0xFFFFu; out[5] = v.z >> 16; out[6] = v.w & 0xFFFFu; out[7] = v.w >> 16;
}
template <> __device__ __forceinline__ void unpack_vec<uint32_t>(const uint4& v, uint32_t* out) {
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <class TokT> __device__ __forceinline__ bool vec_has(const uint4& v, uint32_t a);
template <> __device__ __forceinline__ bool vec_has<uint16_t>(const uint4& v, uint32_t a) {
    // a slot equals `a` iff the 16-bit-lane minimum of the four words ^ aa has a zero lane: two packed-minimum
    // instructions (VIMNMX3.U16x2 / VIMNMX.U16x2, the DPX unit of sm_90+) instead of three bit operations per word
    // (round 2's first profile had 23 % of the merge pass's instructions in this test).
    // "has a zero halfword": (m - 0x00010001) & ~m & 0x80008000. A borrow can only flag a neighbouring lane falsely
    // when the lower lane really is zero, so the any-test is exact.
    const uint32_t aa = a | (a << 16);
    uint32_t m = __vimin3_u16x2(v.x ^ aa, v.y ^ aa, v.z ^ aa);
    m = __vminu2(m, v.w ^ aa);
    return ((m - 0x00010001u) & ~m & 0x80008000u) != 0u;
}
template <> __device__ __forceinline__ bool vec_has<uint32_t>(const uint4& v, uint32_t a) {
    return v.x == a || v.y == a || v.z == a || v.w == a;
}

// Candidate filter. An A can only start an occurrence if its next live token is B; looking no further than the next
// SLOT (inside the 16-byte vector, or the first slot of the following vector, which a shuffle brings in), an A whose
// next slot holds a live token other than B is ruled out on the spot, in registers. What remains — next slot is B,
// a hole (the next live token lies further on) or unknown (last lane of a warp) — is a candidate for the staged path.
// Returns the candidates in vec_mask's bit layout. next_first: first slot of the following vector; next_known: it is valid.
template <class TokT> __device__ __forceinline__ uint32_t vec_candidates(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first, bool next_known);
template <> __device__ __forceinline__ uint32_t vec_candidates<uint16_t>(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first, bool next_known) {
    const uint32_t ma = vec_mask<uint16_t>(v, a);
    if (!ma) return 0u;
    const uint32_t ms = vec_mask<uint16_t>(v, b) | vec_mask<uint16_t>(v, 0xFFFFu);  // slots holding B or a hole
    const bool last_ok = !next_known || next_first == b || next_first == 0xFFFFu;
    // bit i = slot 2i, bit 16 + i = slot 2i + 1: the successor of slot 2i is bit 16 + i, of slot 2i + 1 bit i + 1
    const uint32_t succ = ((ms >> 16) & 0xFu) | (((ms & 0xEu) >> 1) << 16) | (last_ok ? (1u << 19) : 0u);
    return ma & succ;
}
template <> __device__ __forceinline__ uint32_t vec_candidates<uint32_t>(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first, bool next_known) {
    const uint32_t ma = vec_mask<uint32_t>(v, a);
    if (!ma) return 0u;
    const uint32_t ms = vec_mask<uint32_t>(v, b) | vec_mask<uint32_t>(v, 0xFFFFFFFFu);
    const bool last_ok = !next_known || next_first == b || next_first == 0xFFFFFFFFu;
    return ma & (((ms >> 1) & 0x7u) | (last_ok ? 8u : 0u));
}
// The A's of a vector that can start an occurrence as far as the next SLOT tells: the next slot holds B or a hole (the next
// live token lies further on). next_first = the slot after the vector's last one. Same bit layout as vec_mask. Conservative
// (an A whose next live token is B always qualifies) and cheap enough to run on every vector that holds an A: the packed
// minimum of "slot != A or next slot != B" and "slot != A or next slot != hole" is zero exactly at the candidates.
template <class TokT> __device__ __forceinline__ uint32_t vec_pair_candidates(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first);
template <> __device__ __forceinline__ uint32_t vec_pair_candidates<uint16_t>(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first) {
    const uint32_t aa = a | (a << 16), bb = b | (b << 16);
    // y: every slot's successor slot
    const uint32_t y0 = __funnelshift_r(v.x, v.y, 16), y1 = __funnelshift_r(v.y, v.z, 16), y2 = __funnelshift_r(v.z, v.w, 16),
                   y3 = __funnelshift_r(v.w, next_first, 16);
    auto c = [&](uint32_t w, uint32_t y) { const uint32_t x = w ^ aa; return __vminu2(x | (y ^ bb), x | ~y); };
    const uint32_t c0 = c(v.x, y0), c1 = c(v.y, y1), c2 = c(v.z, y2), c3 = c(v.w, y3);
    const uint32_t m = __vminu2(__vimin3_u16x2(c0, c1, c2), c3);
    if (!((m - 0x00010001u) & ~m & 0x80008000u)) return 0u;  // no zero lane anywhere (exact, see vec_has)
    auto z = [&](uint32_t x) { return ~(((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) & 0x80008000u; };
    return (z(c0) >> 15) | (z(c1) >> 14) | (z(c2) >> 13) | (z(c3) >> 12);
}
template <> __device__ __forceinline__ uint32_t vec_pair_candidates<uint32_t>(const uint4& v, uint32_t a, uint32_t b, uint32_t next_first) {
    const uint32_t H = 0xFFFFFFFFu;
    auto ok = [&](uint32_t n) { return n == b || n == H; };
    return ((v.x == a && ok(v.y)) ? 1u : 0u) | ((v.y == a && ok(v.z)) ? 2u : 0u) | ((v.z == a && ok(v.w)) ? 4u : 0u) |
           ((v.w == a && ok(next_first)) ? 8u : 0u);
}
template <class TokT> __device__ __forceinline__ uint32_t vec_first_slot(const uint4& v);
template <> __device__ __forceinline__ uint32_t vec_first_slot<uint16_t>(const uint4& v) { return v.x & 0xFFFFu; }
template <> __device__ __forceinline__ uint32_t vec_first_slot<uint32_t>(const uint4& v) { return v.x; }

// +1 for `key` in the block-private bins (open addressing, 8 probes), falling back to the global
// arrays when the neighbourhood is too crowded
template <int NBIN>
__device__ __forceinline__ void bin_add(uint32_t* bin_key, uint32_t* bin_val, uint32_t key, uint32_t* cntL, uint32_t* cntR) {
    uint32_t s = (key * 0x9E3779B1u) >> 23;  // 9 bits
    s &= (uint32_t)(NBIN - 1);
    for (int probe = 0; probe < 8; probe++) {
        uint32_t k = ((volatile uint32_t*)bin_key)[s];
        if (k == EMPTY_KEY) k = atomicCAS(&bin_key[s], EMPTY_KEY, key), k = (k == EMPTY_KEY) ? key : k;
        if (k == key) { atomicAdd(&bin_val[s], 1u); return; }
        s = (s + 1) & (uint32_t)(NBIN - 1);
    }
    atomicAdd((key & 0x10000u) ? &cntR[key & 0xFFFFu] : &cntL[key], 1u);
}

// Ownership of writes: the thread that holds the A of an occurrence writes X over it and, when the
// consumed partner lies in the same tile, the hole over the partner. A partner that lies in the
// next tile is blanked by that tile ("head duty": its left halo ends in an A that starts an
// occurrence). So a tile without any A and without head duty is streamed and left untouched.
// DELTAS: cntL / cntR / *nxx_out receive the neighbour-pair deltas (train); *nab_out always
// receives the number of merged occurrences. FROMCTL as in halo_kernel.
constexpr int MERGE_NBIN = 512;   // block-private delta bins (train)
constexpr int MERGE_QCAP = 2048;  // queued A positions per tile (denser tiles overflow to in-place handling)

// The staged path of one tile, shared by the register-streaming and the TMA-ring kernels: write the
// halo margins, queue the A positions, resolve occurrences, write X / holes, emit the deltas.
// h is valid in thread 0 only. Ends with all threads past their last read of ext.
// stage the tile (+ halo margins, run length, empty queue, empty bins): everything tile_staged_path reads from shared
// memory. h is valid in thread 0 only. The caller puts a barrier between this and tile_staged_path.
template <class TokT, bool STAGE_FROM_REGS, int NV>
__device__ __forceinline__ void tile_stage(TokT* ext, const uint4 (&v)[NV], const TileHalo<TokT>& h, bool use_bins,
                                           uint32_t* bin_key, uint32_t* bin_val, uint32_t* q_n, uint32_t* sh_runA) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    if (use_bins) for (int i = (int)threadIdx.x; i < MERGE_NBIN; i += THREADS) { bin_key[i] = EMPTY_KEY; bin_val[i] = 0; }
    if (STAGE_FROM_REGS) {
        uint4* dst = reinterpret_cast<uint4*>(ext + OFF);
#pragma unroll
        for (int k = 0; k < NV; k++) dst[k * THREADS + (int)threadIdx.x] = v[k];
    }
    if (threadIdx.x == 0) {
        ext[OFF - 2] = h.l2;
        ext[OFF - 1] = h.l1;
        ext[OFF + TILE + 0] = h.r0;
        ext[OFF + TILE + 1] = h.r1;
        ext[OFF + TILE + 2] = h.r2;
        *sh_runA = h.runA;
        *q_n = 0;
    } else if (threadIdx.x < 32) {
        const int i = (int)threadIdx.x - 1;  // the remaining margin slots are holes
        if (i < OFF - 2) ext[i] = H;
        else if (OFF + TILE + 3 + (i - (OFF - 2)) < EXT) ext[OFF + TILE + 3 + (i - (OFF - 2))] = H;
    }
}

template <class TokT, bool DELTAS, bool STAGE_FROM_REGS, int NV, bool PRESTAGED = false>
__device__ __forceinline__ void tile_staged_path(TokT* ext, const uint4 (&v)[NV], TokT* __restrict__ tok, size_t base,
                                                 const TileHalo<TokT>& h, uint32_t hitbits, uint32_t Au, uint32_t Bu, uint32_t Xu,
                                                 bool use_bins, bool direct, bool pair_filter, bool flush_bins,
                                                 uint32_t* bin_key, uint32_t* bin_val, uint16_t* q_pos, uint32_t* q_n,
                                                 uint32_t* sh_runA, uint32_t* __restrict__ cntL, uint32_t* __restrict__ cntR,
                                                 uint32_t& nAB, uint32_t& nXX) {
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NBIN = MERGE_NBIN;
    constexpr int QCAP = MERGE_QCAP;
    const TokT H = (TokT)TokTraits<TokT>::hole;
    const TokT A = (TokT)Au, B = (TokT)Bu, X = (TokT)Xu;
    const bool AEQB = (Au == Bu);
    // stage the tile (+ halo) so that neighbours can be looked up across thread boundaries
    if (!PRESTAGED) {
        tile_stage<TokT, STAGE_FROM_REGS, NV>(ext, v, h, use_bins, bin_key, bin_val, q_n, sh_runA);
        __syncthreads();
    }

    // a pre-staged tile is read back from shared memory, so that the 16 registers of v[] are free after the prologue
    auto getv = [&](int k) -> uint4 {
        return PRESTAGED ? reinterpret_cast<const uint4*>(ext + OFF)[k * THREADS + (int)threadIdx.x] : v[k];
    };
    // one occurrence candidate: the A at ext index s (A != B). Writes X / hole, emits the deltas.
    auto process_ab = [&](int s) {
        const int j = next_live(ext, s);
        if (j < 0 || ext[j] != B) return;
        tok[base + (size_t)(s - OFF)] = X;
        if (j < OFF + TILE) tok[base + (size_t)(j - OFF)] = H;
        nAB++;
        if (DELTAS) {
            // left side: always owned by this occurrence
            const int p = prev_live(ext, s);
            if (p >= 0) {
                const TokT tp = ext[p];
                bool merged_second = false;
                if (tp == B) { int pp = prev_live(ext, p); merged_second = (pp >= 0 && ext[pp] == A); }
                if (merged_second) nXX++;
                else if (use_bins) bin_add<NBIN>(bin_key, bin_val, (uint32_t)tp, cntL, cntR);
                else atomicAdd(&cntL[tp], 1u);
            }
            // right side: owned only if the next live token does not start another occurrence
            const int n = next_live(ext, j);
            if (n >= 0) {
                const TokT tn = ext[n];
                bool is_start = false;
                if (tn == A) { int nn = next_live(ext, n); is_start = (nn >= 0 && ext[nn] == B); }
                if (!is_start) {
                    if (use_bins) bin_add<NBIN>(bin_key, bin_val, 0x10000u | (uint32_t)tn, cntL, cntR);
                    else atomicAdd(&cntR[tn], 1u);
                }
            }
        }
    };
    if (!AEQB) {
        if (threadIdx.x == 0 && ext[OFF - 1] == A) {
            // head duty: the A just before this tile starts an occurrence iff our first live token is B
            int f = next_live(ext, OFF - 1);
            if (f >= 0 && f < OFF + TILE && ext[f] == B) tok[base + (size_t)(f - OFF)] = H;
        }
        if (direct) {
            // Sparse step (few occurrences per tile, the common case after the first few hundred merges: the pair is
            // rare although its first token is not). Every thread looks its own A's up in the staged tile — no
            // candidate masks, no queue, no shared-memory atomics, one barrier less. With a handful of occurrences per
            // tile the idle lanes cost less than filling and draining a queue does.
#pragma unroll
            for (int k = 0; k < NV; k++) {
                if (!((hitbits >> k) & 1u)) continue;
                const int s0 = OFF + (k * THREADS + (int)threadIdx.x) * VEC;
                // (the slot after the vector comes from the staged tile: the next thread's first slot, or the halo)
                uint32_t mask = pair_filter ? vec_pair_candidates<TokT>(getv(k), Au, Bu, (uint32_t)ext[s0 + VEC]) : vec_mask<TokT>(getv(k), Au);
                while (mask) {
                    const int bit = __ffs((int)mask) - 1;
                    mask &= mask - 1;
                    process_ab(s0 + mask_bit_to_slot<TokT>(bit));
                }
            }
        } else {
        // The A's are sparse and unevenly spread over the lanes, so each thread only queues the
        // positions of its A's; the CTA then works through the queue with all lanes busy.
#pragma unroll
        for (int k = 0; k < NV; k++) {
            // (the shuffle is executed by every lane; vectors k of neighbouring lanes are neighbours in the tile)
            const int s0 = OFF + (k * THREADS + (int)threadIdx.x) * VEC;
            uint32_t mask;
            if (pair_filter) {  // the slot after the vector comes from the staged tile
                if (!((hitbits >> k) & 1u)) continue;
                mask = vec_pair_candidates<TokT>(getv(k), Au, Bu, (uint32_t)ext[s0 + VEC]);
                if (!mask) continue;
            } else {
                const uint4 vk = getv(k);
                const uint32_t nf = __shfl_down_sync(0xffffffffu, vec_first_slot<TokT>(vk), 1);
                if (!((hitbits >> k) & 1u)) continue;
                mask = vec_candidates<TokT>(vk, Au, Bu, nf, (threadIdx.x & 31u) != 31u);
            }
            uint32_t at = atomicAdd(q_n, (uint32_t)__popc(mask));
            while (mask) {
                const int bit = __ffs((int)mask) - 1;
                mask &= mask - 1;
                const int s = s0 + mask_bit_to_slot<TokT>(bit);
                if (at < (uint32_t)QCAP) q_pos[at] = (uint16_t)s;
                else process_ab(s);  // queue full (very dense step): handle in place
                at++;
            }
        }
        __syncthreads();
        const uint32_t nq = *q_n < (uint32_t)QCAP ? *q_n : (uint32_t)QCAP;
        for (uint32_t i = threadIdx.x; i < nq; i += THREADS) process_ab((int)q_pos[i]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            if (!((hitbits >> k) & 1u)) continue;
            uint32_t mask = 0;  // A==B needs the slots in ascending order: bit = slot
            {
                uint32_t m = vec_mask<TokT>(getv(k), Au);
                while (m) { const int bit = __ffs((int)m) - 1; m &= m - 1; mask |= 1u << mask_bit_to_slot<TokT>(bit); }
            }
            const int s0 = OFF + (k * THREADS + (int)threadIdx.x) * VEC;
            uint32_t run = 0;       // consecutive live A's immediately before the current slot
            bool run_known = false;
            while (mask) {
                const int bit = __ffs((int)mask) - 1;
                mask &= mask - 1;
                const int s = s0 + bit;
                if (!run_known) {
                    // live A's immediately before s inside the tile; if the tile start is reached
                    // the run continues into earlier tiles (runA)
                    uint32_t c = 0;
                    bool stop = false;
                    for (int q = s - 1; q >= OFF; q--) {
                        TokT u = ext[q];
                        if (u == H) continue;
                        if (u == A) c++; else { stop = true; break; }
                    }
                    run = stop ? c : c + *sh_runA;
                    run_known = true;
                }
                const uint32_t off = run;
                // the next A of this vector continues the run only if nothing but holes lies between
                if (mask) {
                    const int nb = __ffs((int)mask) - 1;
                    bool contiguous = true;
                    for (int q = s + 1; q < s0 + nb; q++) if (ext[q] != H) { contiguous = false; break; }
                    run = contiguous ? off + 1 : 0;
                }
                if (off & 1u) {
                    tok[base + (size_t)(s - OFF)] = H;  // second element of the occurrence at off-1
                    continue;
                }
                const int j = next_live(ext, s);
                if (j < 0 || ext[j] != A) continue;
                tok[base + (size_t)(s - OFF)] = X;
                nAB++;
                if (DELTAS) {
                    const int p = prev_live(ext, s);
                    if (p >= 0) {
                        const TokT tp = ext[p];
                        if (tp == A) nXX++;  // same run, odd offset: second element of the previous occurrence
                        else if (use_bins) bin_add<NBIN>(bin_key, bin_val, (uint32_t)tp, cntL, cntR);
                        else atomicAdd(&cntL[tp], 1u);
                    }
                    const int n = next_live(ext, j);
                    if (n >= 0) {
                        const TokT tn = ext[n];
                        bool is_start = false;
                        if (tn == A) { int nn = next_live(ext, n); is_start = (nn >= 0 && ext[nn] == A); }
                        if (!is_start) {
                            if (use_bins) bin_add<NBIN>(bin_key, bin_val, 0x10000u | (uint32_t)tn, cntL, cntR);
                            else atomicAdd(&cntR[tn], 1u);
                        }
                    }
                }
            }
        }
    }
    if (use_bins && flush_bins) {
        __syncthreads();
        for (int i = (int)threadIdx.x; i < NBIN; i += THREADS) {
            const uint32_t k = bin_key[i];
            if (k != EMPTY_KEY) atomicAdd((k & 0x10000u) ? &cntR[k & 0xFFFFu] : &cntL[k], bin_val[i]);
        }
    }
}

#ifndef BPE_MERGE_MINBLOCKS
#define BPE_MERGE_MINBLOCKS 6
#endif
template <class TokT, bool DELTAS, bool FROMCTL, bool FILTER, bool PRESTAGE = false>
__global__ void __launch_bounds__(THREADS, BPE_MERGE_MINBLOCKS) merge_kernel(TokT* __restrict__ tok, const TileHalo<TokT>* __restrict__ halo,
                                                        const StepCtl* __restrict__ ctl, uint32_t* __restrict__ cntL,
                                                        uint32_t* __restrict__ cntR, uint32_t* nxx_out, uint32_t* nab_out,
                                                        uint32_t Au, uint32_t Bu, uint32_t Xu, uint32_t bins_min_count,
                                                        int backwards, uint32_t direct_max_count) {
    __shared__ __align__(16) TokT ext[EXT];
    // block-private bins for the neighbour deltas: key = token id (+ 0x10000 for the right side)
    __shared__ uint32_t bin_key[DELTAS ? MERGE_NBIN : 1];
    __shared__ uint32_t bin_val[DELTAS ? MERGE_NBIN : 1];
    __shared__ uint16_t q_pos[MERGE_QCAP];
    __shared__ uint32_t q_n;
    __shared__ uint32_t sh_runA;
    constexpr int VEC = 16 / (int)sizeof(TokT);   // slots per 16-byte vector
    constexpr int NV = TILE / VEC / THREADS;       // vectors per thread (interleaved for coalescing)
    static_assert(NV >= 1 && NV * VEC * THREADS == TILE, "tile geometry");
    const TokT H = (TokT)TokTraits<TokT>::hole;
    // odd steps walk the sequence backwards (`backwards`, chosen by the host from the step parity): the
    // tiles the previous pass touched last are still in the 126 MB L2 when this pass starts with them
    const bool pair_filter = !(backwards & 2);  // (bit 1 of `backwards`: direct path without the pair filter, for comparison)
    const uint32_t tile = (backwards & 1) ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const size_t base = (size_t)tile * TILE;

    // 1. stream the tile through registers. The loads are issued before the control block is read so
    // that the two latencies overlap.
    const uint4* src = reinterpret_cast<const uint4*>(tok + base);
    uint4 v[NV];
    bool any = false;
    TileHalo<TokT> h;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
    if (FROMCTL) {
        // train loop: launched as a programmatic dependent of the apply kernel. The tile itself is only ever written
        // by merge passes (all complete), so its loads are already in flight while that kernel finishes; the halos
        // and the control block are its output.
        BPE_GRID_DEP_WAIT();
        BPE_GRID_DEP_LAUNCH();
    }
    if (threadIdx.x == 0) h = halo[tile];
    bool use_bins = false;  // dense steps privatise the deltas per CTA; sparse steps go straight to global
    bool direct = !FROMCTL && direct_max_count != 0u;  // sparse steps: no candidate queue (tile_staged_path)
    if (FROMCTL) {
        if (ctl->halt) return;
        Au = ctl->A; Bu = ctl->B; Xu = ctl->X;
        const uint32_t mc = ctl->max_count;  // occurrences of (A,B) in the whole corpus
        use_bins = DELTAS && mc >= bins_min_count;
        direct = mc < direct_max_count;
        // tell the apply kernel of this step that the pass ran (its halo CTAs cannot look at `halt`: the selection
        // at the end of that very kernel may set it while they are still starting)
        if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<StepCtl*>(ctl)->pass_step = ctl->step + 1u;
    }
    const TokT A = (TokT)Au;
    uint32_t hitbits = 0;  // which of my vectors hold an A
#pragma unroll
    for (int k = 0; k < NV; k++) hitbits |= vec_has<TokT>(v[k], Au) ? (1u << k) : 0u;
    // The tile needs the staged path only if some A can start an occurrence (candidate filter above; a step whose A is
    // a frequent token but whose pair is rare leaves most tiles on the streaming path) or, for first == second, if it
    // holds an A at all (the run bookkeeping looks at every A).
    if (!FILTER || Au == Bu) any = hitbits != 0;
    else {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const uint32_t nf = __shfl_down_sync(0xffffffffu, vec_first_slot<TokT>(v[k]), 1);
            if (((hitbits >> k) & 1u) && vec_candidates<TokT>(v[k], Au, Bu, nf, (threadIdx.x & 31u) != 31u)) any = true;
        }
    }
    if (threadIdx.x == 0 && h.l1 == A) any = true;  // possible head duty
    // The tile goes to shared memory before it is known whether anybody needs it there: four 16-byte stores per thread
    // cost less than the second barrier a staged tile would otherwise wait at (the barrier below then also publishes the
    // staged tile).
    if (PRESTAGE) tile_stage<TokT, true, NV>(ext, v, h, use_bins, bin_key, bin_val, &q_n, &sh_runA);
    if (!__syncthreads_or(any ? 1 : 0)) return;
    uint32_t nAB = 0, nXX = 0;
    tile_staged_path<TokT, DELTAS, true, NV, PRESTAGE>(ext, v, tok, base, h, hitbits, Au, Bu, Xu, use_bins, direct, pair_filter, true, bin_key, bin_val, q_pos, &q_n, &sh_runA,
                                             cntL, cntR, nAB, nXX);
    if (nAB) atomicAdd(nab_out, nAB);
    if (DELTAS && nXX) atomicAdd(nxx_out, nXX);
}

#ifndef BPE_MERGE_LOOP_MINBLOCKS
#define BPE_MERGE_LOOP_MINBLOCKS 5  // 51 registers: the prefetched tile (16 registers) stays in registers while the staged tile is worked on
#endif
// The train loop's merge pass with resident CTAs (merge_loop = CTAs per SM): a CTA takes the tiles blockIdx.x,
// blockIdx.x + gridDim.x, ... The one-CTA-per-tile kernel above keeps only 57 % of the warp slots busy (a CTA lives for
// ~2 us: launch, load, test, exit), so here (1) the next tile's four vectors and its halo are requested as soon as the
// current tile is staged, and travel while the CTA waits at the barrier and works on the staged tile; (2) the control
// block is read once per CTA; (3) the delta bins live as long as the CTA and are flushed once.
// DBUF: two staging buffers, tiles alternate between them. A buffer is then rewritten only after the barrier of the tile
// in between, which every thread reaches after it has finished with the buffer: a tile of the queue-less path needs no
// barrier of its own any more (the queued path keeps it: its queue is single).
template <bool DBUF>
__global__ void __launch_bounds__(THREADS, BPE_MERGE_LOOP_MINBLOCKS) merge_loop_kernel(uint16_t* __restrict__ tok, const TileHalo<uint16_t>* __restrict__ halo,
                                                        const StepCtl* __restrict__ ctl, uint32_t* __restrict__ cntL,
                                                        uint32_t* __restrict__ cntR, uint32_t* nxx_out, uint32_t* nab_out,
                                                        uint32_t bins_min_count, int backwards, uint32_t direct_max_count, uint32_t ntiles) {
    using TokT = uint16_t;
    static_assert(sizeof(TileHalo<TokT>) == 16, "a halo travels as one 16-byte vector");
    __shared__ __align__(16) TokT ext_all[DBUF ? 2 * EXT : EXT];
    TokT* ext = ext_all;
    __shared__ uint32_t bin_key[MERGE_NBIN];
    __shared__ uint32_t bin_val[MERGE_NBIN];
    __shared__ uint16_t q_pos[MERGE_QCAP];
    __shared__ uint32_t q_n;
    __shared__ uint32_t sh_runA;
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    const bool pair_filter = !(backwards & 2);
    const bool back = (backwards & 1) != 0;
    uint32_t it = blockIdx.x;
    if (it >= ntiles) return;
    uint32_t tile = back ? ntiles - 1u - it : it;
    uint4 v[NV];
    __shared__ __align__(16) uint4 hraw[2];  // halo of the current / the next tile: {l2 | l1 << 16, r0 | r1 << 16, r2, runA}
    uint32_t hsel = 0;
    {
        const uint4* src = reinterpret_cast<const uint4*>(tok + (size_t)tile * TILE);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
    }
    BPE_GRID_DEP_WAIT();
    BPE_GRID_DEP_LAUNCH();
    if (threadIdx.x == 0) cp_async16(&hraw[0], halo + tile);  // (asynchronous: no register, nothing waits for it here)
    if (ctl->halt) { cp_async_wait_all(); return; }  // (no copy into shared memory may outlive the CTA)
    const uint32_t Au = ctl->A, Bu = ctl->B, Xu = ctl->X;
    const uint32_t mc = ctl->max_count;
    const bool use_bins = mc >= bins_min_count;
    const bool direct = mc < direct_max_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<StepCtl*>(ctl)->pass_step = ctl->step + 1u;
    if (use_bins) for (int i = (int)threadIdx.x; i < MERGE_NBIN; i += THREADS) { bin_key[i] = EMPTY_KEY; bin_val[i] = 0; }
    if (threadIdx.x >= 1 && threadIdx.x < 32) {  // the margin slots beyond the halo tokens stay holes for the CTA's lifetime
        const int i = (int)threadIdx.x - 1;
#pragma unroll
        for (int b = 0; b < (DBUF ? 2 : 1); b++) {
            if (i < OFF - 2) ext_all[b * EXT + i] = (TokT)0xFFFFu;
            else if (OFF + TILE + 3 + (i - (OFF - 2)) < EXT) ext_all[b * EXT + OFF + TILE + 3 + (i - (OFF - 2))] = (TokT)0xFFFFu;
        }
    }
    uint32_t nAB = 0, nXX = 0;
    const TileHalo<TokT> h_unused = {};
    while (true) {
        uint32_t hitbits = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) hitbits |= vec_has<TokT>(v[k], Au) ? (1u << k) : 0u;
        bool any = hitbits != 0;
        {   // stage the tile and its halo
            uint4* dst = reinterpret_cast<uint4*>(ext + OFF);
#pragma unroll
            for (int k = 0; k < NV; k++) dst[k * THREADS + (int)threadIdx.x] = v[k];
            if (threadIdx.x == 0) {
                cp_async_wait_all();
                const uint4 hq = hraw[hsel];
                if ((hq.x >> 16) == Au) any = true;  // possible head duty
                *reinterpret_cast<uint32_t*>(ext + OFF - 2) = hq.x;
                *reinterpret_cast<uint32_t*>(ext + OFF + TILE) = hq.y;
                ext[OFF + TILE + 2] = (TokT)hq.z;
                sh_runA = hq.w;
                q_n = 0;
            }
        }
        const size_t base = (size_t)tile * TILE;
        // the next tile's data: on its way while this tile is worked on
        const uint32_t it_next = it + gridDim.x;
        const bool more = it_next < ntiles;
        const uint32_t tile_next = more ? (back ? ntiles - 1u - it_next : it_next) : tile;
        if (more) {
            const uint4* src = reinterpret_cast<const uint4*>(tok + (size_t)tile_next * TILE);
#pragma unroll
            for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
            hsel ^= 1u;
            if (threadIdx.x == 0) cp_async16(&hraw[hsel], halo + tile_next);
        }
        if (__syncthreads_or(any ? 1 : 0)) {
            tile_staged_path<TokT, true, true, NV, true>(ext, v, tok, base, h_unused, hitbits, Au, Bu, Xu, use_bins, direct, pair_filter, false, bin_key,
                                                         bin_val, q_pos, &q_n, &sh_runA, cntL, cntR, nAB, nXX);
            // everybody has read the staged tile (and the queue) before the next one replaces it
            if (more && (!DBUF || !direct || Au == Bu)) __syncthreads();
        }
        if (!more) break;
        it = it_next;
        tile = tile_next;
        if (DBUF) ext = (ext == ext_all) ? ext_all + EXT : ext_all;
    }
    if (nAB) atomicAdd(nab_out, nAB);
    if (nXX) atomicAdd(nxx_out, nXX);
    if (use_bins) {
        __syncthreads();
        for (int i = (int)threadIdx.x; i < MERGE_NBIN; i += THREADS) {
            const uint32_t k = bin_key[i];
            if (k != EMPTY_KEY) atomicAdd((k & 0x10000u) ? &cntR[k & 0xFFFFu] : &cntL[k], bin_val[i]);
        }
    }
}

// =========================================================================================
// apply_kernel: fold the per-neighbour merge deltas into the pair table (and into the
// reference-home population on births / deaths). One thread per token id.
// =========================================================================================
__device__ __forceinline__ void tbl_sub(const PairTable& tbl, uint32_t key, uint32_t c, StepCtl* ctl, const ZigPop& z) {
    uint32_t s = tbl_find(tbl, key);
    if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_KEY_MISSING); return; }
    uint32_t old = atomicSub(&tbl.counts[s], c);
    if (old < c) { atomicOr(&ctl->err, (uint32_t)ERR_COUNT_UNDERFLOW); return; }
    if (old == c) {  // death
        atomicSub(&ctl->live_keys, 1u);
        zcnt_add(z, key, -1, ctl);
    }
}
__device__ __forceinline__ void tbl_add(const PairTable& tbl, uint32_t key, uint32_t c, StepCtl* ctl, const ZigPop& z,
                                        const HeavyList& hl) {
    uint32_t s = tbl_find_or_insert(tbl, key, &ctl->n_inserted);
    if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_TABLE_FULL); return; }
    uint32_t old = atomicAdd(&tbl.counts[s], c);
    if (hl.theta && old < hl.theta && old + c >= hl.theta) {  // newly heavy
        uint32_t idx = atomicAdd(&ctl->n_heavy, 1u);
        if (idx < hl.cap) hl.slots[idx] = s;
    }
    if (old == 0) {  // birth
        atomicAdd(&ctl->live_keys, 1u);
        zcnt_add(z, key, +1, ctl);
    }
}

// rebuild the reference-home population from the table (capacity change / table rebuild)
__global__ void zig_rebuild_kernel(PairTable tbl, ZigPop z, StepCtl* ctl) {
    size_t cap = (size_t)tbl.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x)
        if (tbl.counts[i] > 0) zcnt_add(z, tbl.keys[i], +1, ctl);
}

// =========================================================================================
// merge_tma_kernel: the same pass as merge_kernel, fed by a TMA ring. A persistent CTA owns every
// gridDim.x-th tile; one thread keeps RING_STAGES bulk copies (cp.async.bulk, 16 KB each) in flight
// into shared memory, each completing on its own mbarrier. The tile therefore arrives already
// staged: the streaming test reads it from shared memory, and a tile that needs work pays no extra
// copy and no extra load latency (the next tiles are already on their way). In-flight data per SM
// = 2 CTAs x RING_STAGES x 16 KB, independent of the register budget.
// =========================================================================================
#ifndef BPE_RING_STAGES
#define BPE_RING_STAGES 3
#endif
#ifndef BPE_RING_CTAS
#define BPE_RING_CTAS 4
#endif
constexpr int RING_STAGES = BPE_RING_STAGES;  // bulk copies in flight per CTA
constexpr int RING_CTAS_PER_SM = BPE_RING_CTAS;
template <class TokT> __host__ __device__ constexpr size_t ring_stage_bytes() { return (size_t)EXT * sizeof(TokT); }
template <class TokT> __host__ __device__ constexpr size_t ring_smem_bytes() { return RING_STAGES * ring_stage_bytes<TokT>(); }

template <class TokT, bool DELTAS, bool FROMCTL>
__global__ void __launch_bounds__(THREADS, RING_CTAS_PER_SM) merge_tma_kernel(TokT* __restrict__ tok, const TileHalo<TokT>* __restrict__ halo,
                                                               const StepCtl* __restrict__ ctl, uint32_t* __restrict__ cntL,
                                                               uint32_t* __restrict__ cntR, uint32_t* nxx_out, uint32_t* nab_out,
                                                               uint32_t Au, uint32_t Bu, uint32_t Xu, uint32_t bins_min_count,
                                                               uint32_t ntiles) {
    __shared__ uint32_t bin_key[DELTAS ? MERGE_NBIN : 1];
    __shared__ uint32_t bin_val[DELTAS ? MERGE_NBIN : 1];
    __shared__ uint16_t q_pos[MERGE_QCAP];
    __shared__ uint32_t q_n;
    __shared__ uint32_t sh_runA;
    __shared__ __align__(8) uint64_t bar[RING_STAGES];
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    constexpr uint32_t TILE_BYTES = (uint32_t)(TILE * sizeof(TokT));
    bool use_bins = false;
    bool backwards = false;
    if (FROMCTL) {
        if (ctl->halt) return;
        Au = ctl->A; Bu = ctl->B; Xu = ctl->X;
        use_bins = DELTAS && ctl->max_count >= bins_min_count;
        backwards = (ctl->step & 1u) != 0;  // odd steps walk backwards: the previous pass's tail is still in L2
    }
    const TokT A = (TokT)Au;
    unsigned char* ring = reinterpret_cast<unsigned char*>(bpe_dyn_smem());
    const uint32_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t n_my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    auto tile_of = [&](uint32_t j) { uint32_t t = first + j * stride; return backwards ? ntiles - 1u - t : t; };
    auto stage_ptr = [&](uint32_t s) { return reinterpret_cast<TokT*>(ring + (size_t)s * ring_stage_bytes<TokT>()); };
    auto issue = [&](uint32_t j) {  // thread 0 only
        const uint32_t s = j % RING_STAGES;
        mbar_arrive_expect_tx(&bar[s], TILE_BYTES);
        tma_load_1d(stage_ptr(s) + OFF, tok + (size_t)tile_of(j) * TILE, TILE_BYTES, &bar[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < RING_STAGES; s++) mbar_init(&bar[s], 1);
        mbar_fence_init();
        fence_proxy_async();
        for (uint32_t j = 0; j < n_my && j < (uint32_t)RING_STAGES; j++) issue(j);
    }
    __syncthreads();

    uint32_t nAB = 0, nXX = 0;
    for (uint32_t j = 0; j < n_my; j++) {
        const uint32_t s = j % RING_STAGES;
        const uint32_t tile = tile_of(j);
        TileHalo<TokT> h;
        if (threadIdx.x == 0) h = halo[tile];  // its latency hides behind the wait
        mbar_wait(&bar[s], (j / RING_STAGES) & 1u);
        TokT* ext = stage_ptr(s);
        const uint4* src = reinterpret_cast<const uint4*>(ext + OFF);
        uint4 v[NV];
        bool any = false;
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
        uint32_t hitbits = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) hitbits |= vec_has<TokT>(v[k], Au) ? (1u << k) : 0u;
        any = hitbits != 0;
        if (threadIdx.x == 0 && h.l1 == A) any = true;  // possible head duty
        if (__syncthreads_or(any ? 1 : 0)) {
            tile_staged_path<TokT, DELTAS, false, NV>(ext, v, tok, (size_t)tile * TILE, h, hitbits, Au, Bu, Xu, use_bins, false, false, true, bin_key, bin_val,
                                                      q_pos, &q_n, &sh_runA, cntL, cntR, nAB, nXX);
            __syncthreads();  // every thread is done with this stage
        }
        if (threadIdx.x == 0 && j + RING_STAGES < n_my) {
            fence_proxy_async();  // order the generic-proxy accesses to the stage before the async refill
            issue(j + RING_STAGES);
        }
    }
    if (nAB) atomicAdd(nab_out, nAB);
    if (DELTAS && nXX) atomicAdd(nxx_out, nXX);
}

// =========================================================================================
// Candidate-scan merge path (A != B, single GPU): three small steps instead of halo + tiled merge.
//   scan_kernel    streams the sequence once through registers (no shared memory, no barrier). An A
//                  whose successor inside the same 16-byte vector is a live token other than B can never
//                  start an occurrence and is dropped on the spot; only the others (successor is B, a
//                  hole, or lies in the next vector) are queued by position.
//   resolve_kernel one thread per queued A walks the (still unmodified) sequence in global memory,
//                  decides the occurrence, emits the neighbour deltas and the token writes.
//   write_kernel   applies the writes (X over A, hole over the consumed B).
// Nothing is modified before every decision has been taken, so no tiles, halos or staging are needed.
// =========================================================================================
struct TokWrite { uint32_t pos; uint32_t val; };

template <class TokT>
__global__ void __launch_bounds__(THREADS, BPE_MERGE_MINBLOCKS) scan_kernel(const TokT* __restrict__ tok, StepCtl* ctl,
                                                                           uint32_t* __restrict__ cand, uint32_t cand_cap, int backwards) {
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    const uint32_t H = TokTraits<TokT>::hole;
    const uint32_t tile = backwards ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const size_t base = (size_t)tile * TILE;
    const uint4* src = reinterpret_cast<const uint4*>(tok + base);
    uint4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
    if (ctl->halt) return;
    const uint32_t Au = ctl->A, Bu = ctl->B;
    if (Au == Bu) {  // runs need parity from the run start: the tiled pass handles them
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->halt = H_CLASSIC;
        return;
    }
#pragma unroll
    for (int k = 0; k < NV; k++) {
        if (!vec_has<TokT>(v[k], Au)) continue;
        uint32_t tv[VEC];
        unpack_vec<TokT>(v[k], tv);
        const uint32_t p0 = (uint32_t)(base + (size_t)(k * THREADS + (int)threadIdx.x) * VEC);
#pragma unroll
        for (int i = 0; i < VEC; i++) {
            if (tv[i] != Au) continue;
            bool c = true;  // last slot of the vector: the successor is out of reach
            if (i + 1 < VEC) c = (tv[i + 1] == Bu) || (tv[i + 1] == H);
            if (c) {
                const uint32_t at = atomicAdd(&ctl->cand_n, 1u);
                if (at < cand_cap) cand[at] = p0 + (uint32_t)i;
            }
        }
    }
}

template <class TokT>
__global__ void resolve_kernel(const TokT* __restrict__ tok, size_t n_slots, StepCtl* ctl, const uint32_t* __restrict__ cand,
                               uint32_t cand_cap, TokWrite* __restrict__ wr, uint32_t* __restrict__ cntL, uint32_t* __restrict__ cntR,
                               uint32_t* nxx_out, uint32_t* nab_out) {
    const TokT H = (TokT)TokTraits<TokT>::hole;
    if (ctl->halt) return;
    const uint32_t n = ctl->cand_n;
    if (n > cand_cap) {  // too dense for the queue: nothing has been modified yet, the tiled pass redoes the step
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->halt = H_CLASSIC;
        return;
    }
    const TokT A = (TokT)ctl->A, B = (TokT)ctl->B;
    const uint32_t X = ctl->X;
    auto next_of = [&](size_t i) -> long long { for (size_t j = i + 1; j < n_slots; j++) if (tok[j] != H) return (long long)j; return -1; };
    auto prev_of = [&](size_t i) -> long long { for (size_t j = i; j > 0;) { --j; if (tok[j] != H) return (long long)j; } return -1; };
    uint32_t nAB = 0, nXX = 0;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const size_t s = cand[c];
        const long long j = next_of(s);
        if (j < 0 || tok[j] != B) continue;
        const uint32_t w = atomicAdd(&ctl->w_n, 2u);
        wr[w].pos = (uint32_t)s; wr[w].val = X;
        wr[w + 1].pos = (uint32_t)j; wr[w + 1].val = (uint32_t)H;
        nAB++;
        // left side: always owned by this occurrence
        const long long p = prev_of(s);
        if (p >= 0) {
            const TokT tp = tok[p];
            bool merged_second = false;
            if (tp == B) { const long long pp = prev_of((size_t)p); merged_second = (pp >= 0 && tok[pp] == A); }
            if (merged_second) nXX++; else atomicAdd(&cntL[tp], 1u);
        }
        // right side: owned only if the next live token does not start another occurrence
        const long long nx = next_of((size_t)j);
        if (nx >= 0) {
            const TokT tn = tok[nx];
            bool is_start = false;
            if (tn == A) { const long long nn = next_of((size_t)nx); is_start = (nn >= 0 && tok[nn] == B); }
            if (!is_start) atomicAdd(&cntR[tn], 1u);
        }
    }
    if (nAB) atomicAdd(nab_out, nAB);
    if (nXX) atomicAdd(nxx_out, nXX);
}

template <class TokT>
__global__ void write_kernel(TokT* __restrict__ tok, StepCtl* ctl, const TokWrite* __restrict__ wr) {
    if (ctl->halt) return;
    const uint32_t n = ctl->w_n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) tok[wr[i].pos] = (TokT)wr[i].val;
}

// =========================================================================================
// apply_kernel: fold the merge deltas into the pair table (and the reference-home population on
// births / deaths), one thread per token id, then advance the device-side step counter.
// delta layout: [0,vcap) cntL, [vcap,2*vcap) cntR, [2*vcap] cntXX, [2*vcap+1] cntAB
// =========================================================================================
// Peer mode (ps_on): the deltas are the sum over the mailbox slots the peers filled; wait for their
// arrival flags first, and copy the gathered EdgeInfo slots to where the next halo pass reads them.
// Fused halo: the CTAs after the first `apply_blocks` gather the tile halos for the NEXT merge pass (they only need
// the sequence as the pass of this step left it, not the merge that is being chosen), so the per-step halo launch
// and its latency disappear behind the table updates. Steps with first == second additionally need run lengths
// that depend on the chosen token: select_body halts the loop for them (H_AEQB) and the host runs the classic
// halo_kernel for that step.
struct HaloArgs {
    const uint16_t* tok;     // nullptr: no fused halo
    size_t n_slots;
    uint32_t ntiles;
    TileHalo<uint16_t>* halo;
    uint32_t apply_blocks;
    uint32_t step1;          // 1 + index of the step this launch belongs to (compared with StepCtl::pass_step)
};
__global__ void __launch_bounds__(256) apply_kernel(PairTable tbl, uint32_t* __restrict__ delta, uint32_t vcap, StepCtl* ctl,
                             ZigPop z, uint32_t n_ids, HeavyList hl, MergeRec* rec, int fuse_select,
                             int ps_on, PeerSet ps, int rank, int world, uint32_t parity, uint32_t epoch, uint32_t edge_off,
                             HaloArgs ha, PushArgs pa) {
    BPE_GRID_DEP_WAIT();    // (launched as a programmatic dependent of the merge pass: everything below reads its output)
    BPE_GRID_DEP_LAUNCH();  // the next merge pass may start and load its tiles; it waits for this kernel before anything else
    if (ha.tok ? (ctl->pass_step != ha.step1) : (ctl->halt != 0u)) return;  // this step did not run (the loop was halted before it)
    if (ha.tok && blockIdx.x >= ha.apply_blocks) {
        const uint32_t t = (blockIdx.x - ha.apply_blocks) * blockDim.x + threadIdx.x;
        if (t >= ha.ntiles) return;
        EdgeSrc es;
        es.local = world > 1 ? reinterpret_cast<const EdgeInfo*>(delta + edge_off) : nullptr;
        es.mbox = nullptr; es.flags = nullptr; es.slot_words = 0; es.edge_off = edge_off; es.epoch = epoch; es.err = &ctl->err;
        if (ps_on) {  // peer path: the neighbours' shard ends of this step sit in the mailbox (apply block 0 copies them later)
            es.mbox = ps.mbox[rank] + 2 * ((size_t)parity * (size_t)world * ps.slot_words);
            es.flags = nullptr;
            es.slot_words = ps.slot_words;
        }
        halo_gather<uint16_t>(ha.tok, ha.n_slots, t, ha.halo, es, rank, world);
        return;
    }
    const uint32_t n_apply = ha.tok ? ha.apply_blocks : gridDim.x;
    const uint32_t* mb = nullptr;
    // thread t: token id p = t / 4, side = left/right neighbour, op = retire the old pair / credit the new one.
    // The two ops of a (p, side) sit in adjacent lanes so the four table round trips overlap.
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = t >> 2, side = (t >> 1) & 1u, op = t & 1u;
    const uint32_t t0 = 4u * n_ids;  // three more threads: adjacent occurrences, the merged pair itself, bookkeeping
    const uint32_t my_cell = (p < n_ids) ? (side ? vcap + p : p) : (t == t0 ? 2u * vcap : (t == t0 + 2u ? 2u * vcap + 1u : 0xFFFFFFFFu));
    if (ps_on) {
        // ---- my cells of this rank's deltas into slot `rank` of every peer's mailbox (tagged 8-byte stores, no fence) ----
        const size_t slot = ((size_t)parity * (size_t)world + (size_t)rank) * ps.slot_words;
        if (blockIdx.x == 0) {
            EdgeInfo* edges = reinterpret_cast<EdgeInfo*>(delta + edge_off);
            if (threadIdx.x < 32) edge_body<uint16_t>(pa.tok, pa.n_slots, pa.tail_hint, edges, rank, world, ctl, nullptr, 0, false);
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < 16u * (uint32_t)world; i += blockDim.x) {
                const uint32_t q = i / 16u, w = i % 16u;
                if ((int)q != rank) peer_put(ps.mbox[q], slot + edge_off + (uint32_t)rank * 16u + w, delta[edge_off + (uint32_t)rank * 16u + w], epoch);
            }
        }
        if (my_cell != 0xFFFFFFFFu && (op == 0u || p >= n_ids)) {
            const uint32_t c_local = delta[my_cell];
            for (int q = 0; q < world; q++) if (q != rank) peer_put(ps.mbox[q], slot + my_cell, c_local, epoch);
        }
        mb = ps.mbox[rank] + 2 * ((size_t)parity * (size_t)world * ps.slot_words);
        if (blockIdx.x == 0 && threadIdx.x < 16u * (uint32_t)world) {  // gathered shard ends for host-driven halo passes
            const uint32_t r = threadIdx.x / 16u, w = threadIdx.x % 16u;
            if ((int)r != rank) delta[edge_off + r * 16u + w] = peer_get(mb, (size_t)r * ps.slot_words + edge_off + r * 16u + w, epoch, &ctl->err);
        }
    }
    auto cell_value = [&](uint32_t idx) -> uint32_t {
        uint32_t sum = delta[idx];
        if (ps_on) {
            // the peers' cells, all requested at once; the ones that do not carry this step's tag yet are polled
            uint2 v[MAX_PEERS];
#pragma unroll
            for (int r = 0; r < MAX_PEERS; r++)
                if (r < world && r != rank) v[r] = __ldcv(reinterpret_cast<const uint2*>(mb) + ((size_t)r * ps.slot_words + idx));
#pragma unroll
            for (int r = 0; r < MAX_PEERS; r++)
                if (r < world && r != rank) sum += v[r].y == epoch ? v[r].x : peer_get(mb, (size_t)r * ps.slot_words + idx, epoch, &ctl->err);
        }
        return sum;
    };
    const uint32_t A = ctl->A, B = ctl->B, X = ctl->X;
    z.zmask = ctl->zcap - 1;
    hl.theta = ctl->theta;
    uint32_t c = 0;
    uint32_t cell = 0;
    if (p < n_ids && p <= X) {
        cell = side ? vcap + p : p;
        c = cell_value(cell);
    }
    __syncwarp();
    if (c) {
        if (op == 0) tbl_sub(tbl, side ? pair_key(B, p) : pair_key(p, A), c, ctl, z);
        else { delta[cell] = 0; tbl_add(tbl, side ? pair_key(X, p) : pair_key(p, X), c, ctl, z, hl); }
    }
    // (t0 is a multiple of 4, so the three scalar threads share a warp: all read the scalars, then one clears them)
    uint32_t scalar = 0;
    if (t == t0 || t == t0 + 1) scalar = cell_value(2 * vcap);
    else if (t == t0 + 2) scalar = cell_value(2 * vcap + 1);
    __syncwarp();
    if (t == t0) {
        if (scalar) tbl_sub(tbl, pair_key(B, A), scalar, ctl, z);
    } else if (t == t0 + 1) {
        if (scalar) tbl_add(tbl, pair_key(X, X), scalar, ctl, z, hl);
    } else if (t == t0 + 2) {
        if (ps_on) ctl->local_live -= delta[2 * vcap + 1];  // this shard's own merged occurrences (the sum is global)
        delta[2 * vcap] = 0;      // local scalars ready for the next merge pass
        delta[2 * vcap + 1] = 0;
        ctl->cand_n = 0;          // candidate-scan path: queues ready for the next step
        ctl->w_n = 0;
        if (scalar) tbl_sub(tbl, pair_key(A, B), scalar, ctl, z);
        ctl->last_merged = scalar;
        ctl->live_tokens -= scalar;
        ctl->step += 1;
    }
    if (!fuse_select) return;
    // the last CTA to finish chooses the next merge (saves a launch and its gap on every step)
    __shared__ uint32_t s_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(&ctl->apply_done, 1u);
        s_is_last = (prev == n_apply - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_is_last) return;
    if (threadIdx.x == 0) ctl->apply_done = 0;
    __threadfence();
    select_body(tbl, hl, ctl, rec, z);
}


// =========================================================================================
// pair_visit_kernel: one CTA per tile, calls op(pos, key) for every live adjacent pair whose left
// token lies in the tile (pos = slot index of the left token). Used by the replay path
// (first-occurrence positions) and by verify mode (full recount).
// =========================================================================================
struct FirstPosOp {
    PairTable tbl; uint32_t* firstpos; StepCtl* ctl;
    __device__ __forceinline__ void operator()(size_t pos, uint32_t key) const {
        uint32_t s = tbl_find(tbl, key);
        if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_KEY_MISSING); return; }
        uint32_t p = (uint32_t)pos;
        if (p < ((volatile uint32_t*)firstpos)[s]) atomicMin(&firstpos[s], p);
        if (p > ((volatile uint32_t*)&ctl->last_pair_pos)[0]) atomicMax(&ctl->last_pair_pos, p);
    }
};
struct RecountOp {
    PairTable tbl; uint32_t* recount; StepCtl* ctl;
    __device__ __forceinline__ void operator()(size_t, uint32_t key) const {
        uint32_t s = tbl_find(tbl, key);
        if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_KEY_MISSING); return; }
        atomicAdd(&recount[s], 1u);
    }
};

template <class TokT, class Op>
__global__ void __launch_bounds__(THREADS) pair_visit_kernel(const TokT* __restrict__ tok,
                                                             const TileHalo<TokT>* __restrict__ halo, Op op) {
    __shared__ __align__(16) TokT ext[EXT];
    const TokT H = (TokT)TokTraits<TokT>::hole;
    const size_t base = (size_t)blockIdx.x * TILE;
    stage_tile(ext, tok, base, halo[blockIdx.x]);
    __syncthreads();
    const int s0 = OFF + (int)threadIdx.x * SPT;
    for (int s = s0; s < s0 + SPT; s++) {
        TokT t = ext[s];
        if (t == H) continue;
        int j = next_live(ext, s);
        if (j < 0) continue;
        op(base + (size_t)(s - OFF), pair_key((uint32_t)t, (uint32_t)ext[j]));
    }
}

__global__ void verify_counts_kernel(PairTable tbl, const uint32_t* __restrict__ recount, StepCtl* ctl,
                                     uint32_t* live_out) {
    size_t cap = (size_t)tbl.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t c = tbl.counts[i];
        if (c != recount[i]) {
            uint32_t k = atomicAdd(&ctl->verify_mismatch, 1u);
            if (k < 8) printf("verify: key (%u,%u) table=%u recount=%u\n", tbl.keys[i] & 0xFFFFu, tbl.keys[i] >> 16, c, recount[i]);
        }
        if (c) atomicAdd(live_out, 1u);
    }
}

// =========================================================================================
// level_kernel (encode): apply a whole set of mutually independent merges in ONE pass. The host
// (build_encode_schedule) groups the merge list into levels such that, inside a level, no merge uses
// a token another one produces and no token is the second component of one pair and the first of
// another. Occurrences of different pairs of a level therefore never overlap and never create or
// destroy each other, so every slot can decide on its own: a live token that is the first component
// of a level pair and whose next live token completes it becomes the new id; a live token that is the
// second component of a level pair whose previous live token completes it becomes a hole. The result
// is the same as running the level's merges one after the other (src/basic_tokenizer.zig:71-88).
// One CTA per tile; the level's pairs live in shared memory (2-bit role map per id + open-addressing
// hash pair -> new id).
// =========================================================================================
struct LevelEntry { uint32_t key; uint32_t z; };
constexpr int LVL_MAX = 1024;        // pairs per pass (larger levels are split; any subset of a level is a level)
constexpr int LVL_HASH_MAX = 2048;   // hash slots for LVL_MAX pairs; smaller levels use fewer (power of two >= 2 * pairs)
constexpr int LVL_ROLE_WORDS = 65536 / 32;  // 1 bit per id: "first component of a pair of this level"
constexpr int LVL_QCAP = 2048;              // candidates of a tile worked off through the CTA-wide queue
constexpr int LVL_BYTE_IDS_MAX = 16384;     // MODE 1 keeps one role byte per id (ids below this bound)
// MODE 0: role_words = LVL_ROLE_WORDS (bitmap over all ids); MODE 1: role_words = ids / 4 (one byte per id);
// MODE 2: role_words = 2 * LVL_ROLE_WORDS (first-component bitmap + a 65,536-bit Bloom filter over the level's PAIRS)
__host__ __device__ __forceinline__ uint32_t lvl_pair_hash(uint32_t a, uint32_t b) { return ((a * 40503u) ^ (b * 30011u) ^ (b >> 5)) & 0xFFFFu; }
template <class TokT> __host__ __device__ constexpr size_t level_smem_bytes(int hash_slots, int role_words = LVL_ROLE_WORDS) {
    return (size_t)EXT * sizeof(TokT) + (size_t)role_words * 4 + (size_t)hash_slots * 4 + (size_t)hash_slots * 2;
}
__device__ __forceinline__ uint32_t lvl_find(const uint32_t* hkey, const uint16_t* hval, uint32_t key, uint32_t hash_shift,
                                             uint32_t hash_mask) {
    if (key == EMPTY_KEY) return EMPTY_KEY;  // (65535,65535): first == second pairs are never part of a level pass
    uint32_t s = (key * 0x9E3779B1u) >> hash_shift;
    for (uint32_t probe = 0; probe <= hash_mask; probe++) {
        const uint32_t k = hkey[s];
        if (k == key) return (uint32_t)hval[s];
        if (k == EMPTY_KEY) return EMPTY_KEY;
        s = (s + 1) & hash_mask;
    }
    return EMPTY_KEY;
}

// Work split: (1) every thread tests its 32 slots against the role bitmap (uniform, no divergence) and keeps a
// 32-bit candidate mask; (2) it walks its candidates with ONE copy of the look-up code, working in place on the
// staged tile: the slot of the first component becomes the new id and, if the second component lies in the same
// tile, that slot becomes a hole. In-place is safe inside a level: a slot is written only by the owner of the
// occurrence it belongs to, and whatever a concurrent reader can see instead of the original token (a new id, which
// no pair of the level mentions) leads to the same decision. A second component that lies in the NEXT tile is
// turned into a hole by that tile (its first live token against the previous tile's last one, from the halo
// snapshot). (3) the threads write their vectors back if they changed. CTAs take tiles round-robin so the level's
// tables are built once per CTA.
// MODE 1 (encode_filter = 1, not yet measured on a GPU): one role BYTE per id (bit 0 first, bit 1 second component;
// ids < role_words * 4 <= LVL_BYTE_IDS_MAX) and a register-level filter — a slot stays a candidate only if its
// in-vector successor is a second component of the level, a hole, or out of the vector — so that far fewer slots reach
// the hash look-up.
// MODE 2 (encode_filter = 2): the filter looks at the PAIR a slot forms with its successor slot (the next vector's first slot
// comes in by shuffle): one bit of a 65,536-bit Bloom filter over the level's pairs decides (false positives ~ pairs / 65,536),
// so that practically only real occurrences reach the hash look-up; a slot whose successor slot is a hole (or unknown) falls
// back to the first-component bitmap.
template <class TokT, int MODE>
__global__ void __launch_bounds__(THREADS, sizeof(TokT) == 2 ? 6 : 4) level_kernel(TokT* __restrict__ tok, const TileHalo<TokT>* __restrict__ halo,
                                                        const LevelEntry* __restrict__ ents, uint32_t n_ent, uint32_t* nab_out,
                                                        int backwards, uint32_t ntiles, uint32_t hash_log2, uint32_t role_words) {
    unsigned char* raw = reinterpret_cast<unsigned char*>(bpe_dyn_smem());
    TokT* ext = reinterpret_cast<TokT*>(raw);
    uint32_t* role = reinterpret_cast<uint32_t*>(raw + (size_t)EXT * sizeof(TokT));
    if (MODE == 0) role_words = LVL_ROLE_WORDS;
    if (MODE == 2) role_words = 2 * LVL_ROLE_WORDS;
    const unsigned char* role8 = reinterpret_cast<const unsigned char*>(role);
    uint32_t* hkey = role + role_words;
    const uint32_t hash_slots = 1u << hash_log2, hash_mask = hash_slots - 1u, hash_shift = 32u - hash_log2;
    uint16_t* hval = reinterpret_cast<uint16_t*>(hkey + hash_slots);
    __shared__ uint32_t sh_n;
    __shared__ uint32_t lq_n;
    __shared__ uint16_t lq_pos[LVL_QCAP];
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    static_assert(NV * VEC <= 32, "candidate mask is 32 bits");
    const uint32_t H = TokTraits<TokT>::hole;
    for (uint32_t i = threadIdx.x; i < role_words; i += THREADS) role[i] = 0u;
    for (uint32_t i = threadIdx.x; i < hash_slots; i += THREADS) hkey[i] = EMPTY_KEY;
    if (threadIdx.x == 0) { sh_n = 0u; lq_n = 0u; }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n_ent; e += THREADS) {
        const uint32_t key = ents[e].key, a = key & 0xFFFFu;
        if (MODE == 0) {
            atomicOr(&role[a >> 5], 1u << (a & 31u));
        } else if (MODE == 2) {
            const uint32_t h = lvl_pair_hash(a, key >> 16);
            atomicOr(&role[a >> 5], 1u << (a & 31u));
            atomicOr(&role[LVL_ROLE_WORDS + (h >> 5)], 1u << (h & 31u));
        } else {
            const uint32_t b = key >> 16;
            atomicOr(&role[a >> 2], 1u << ((a & 3u) * 8u));
            atomicOr(&role[b >> 2], 2u << ((b & 3u) * 8u));
        }
        uint32_t s = (key * 0x9E3779B1u) >> hash_shift;
        while (true) {
            const uint32_t old = atomicCAS(&hkey[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY) { hval[s] = (uint16_t)ents[e].z; break; }
            s = (s + 1) & hash_mask;
        }
    }
    uint32_t merged = 0;
    for (uint32_t it = blockIdx.x; it < ntiles; it += gridDim.x) {
        const uint32_t tile = backwards ? ntiles - 1u - it : it;
        uint4* gv = reinterpret_cast<uint4*>(tok + (size_t)tile * TILE);
        uint4 v[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = gv[k * THREADS + (int)threadIdx.x];
        uint4* xv = reinterpret_cast<uint4*>(ext + OFF);
        __syncthreads();  // the previous tile's write-back has read the stage (first round: the tables are complete)
#pragma unroll
        for (int k = 0; k < NV; k++) xv[k * THREADS + (int)threadIdx.x] = v[k];
        if (threadIdx.x == 0) {
            const TileHalo<TokT> h = halo[tile];
            for (int i = 0; i < OFF - 2; i++) ext[i] = (TokT)H;
            ext[OFF - 2] = h.l2; ext[OFF - 1] = h.l1;
            ext[OFF + TILE + 0] = h.r0; ext[OFF + TILE + 1] = h.r1; ext[OFF + TILE + 2] = h.r2;
            for (int i = OFF + TILE + 3; i < EXT; i++) ext[i] = (TokT)H;
        }
        // (1) candidates: my slots whose token is the first component of some pair of the level
        uint32_t cand = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t tv[VEC];
            unpack_vec<TokT>(v[k], tv);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const uint32_t t = tv[i] & 0xFFFFu;  // a hole maps to id 65535, whose bit is only set on u32 slots
                    const uint32_t bit = (role[t >> 5] >> (t & 31u)) & 1u;
                    cand |= ((sizeof(TokT) == 4 && tv[i] == H) ? 0u : bit) << (k * VEC + i);
                }
            } else if (MODE == 2) {
                uint32_t nf = __shfl_down_sync(0xffffffffu, tv[0], 1);  // first slot of the next vector (every lane takes part)
                if ((threadIdx.x & 31u) == 31u) nf = H;                    // unknown: treated like a hole
                uint32_t cm = 0;
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const uint32_t a = tv[i], b = (i + 1 < VEC) ? tv[i + 1] : nf;
                    const uint32_t ai = a & 0xFFFFu, h = lvl_pair_hash(ai, b & 0xFFFFu);
                    const bool pair_known = b != H;
                    const uint32_t w = pair_known ? (uint32_t)LVL_ROLE_WORDS + (h >> 5) : (ai >> 5), sh = pair_known ? (h & 31u) : (ai & 31u);
                    const uint32_t bit = (role[w] >> sh) & 1u;
                    cm |= (a != H ? bit : 0u) << i;
                }
                cand |= cm << (k * VEC);
            } else {
                uint32_t fm = 0, sm = 0;  // first-component slots; slots that are a second component or a hole
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const bool hole = tv[i] == H;
                    const uint32_t r = hole ? 2u : (uint32_t)role8[tv[i]];
                    fm |= (r & 1u) << i;
                    sm |= (r >> 1) << i;
                }
                cand |= (fm & ((sm >> 1) | (1u << (VEC - 1)))) << (k * VEC);
            }
        }
        // (u16 slots never hold id 65535 as a token, so its role bit is never set and holes drop out by themselves)
        // The candidates are unevenly spread over the lanes (a walk with 5 of 32 lanes busy was half of the kernel's
        // instructions, profiles/r02_level_kernel_and_tilenc_v5.csv): they go through a CTA-wide queue and are worked
        // off with all lanes busy; what does not fit the queue is handled in place by its own thread.
        auto process = [&](int s) {
            const int j = next_live(ext, s);
            if (j < 0) return;
            const uint32_t z = lvl_find(hkey, hval, pair_key((uint32_t)ext[s], (uint32_t)ext[j]), hash_shift, hash_mask);
            if (z == EMPTY_KEY) return;
            ext[s] = (TokT)z;
            if (j < OFF + TILE) ext[j] = (TokT)H;
            merged++;
        };
        uint32_t at = cand ? atomicAdd(&lq_n, (uint32_t)__popc(cand)) : 0u;
        uint32_t late = 0;  // my candidates that did not fit the queue
        while (cand) {
            const int c = __ffs((int)cand) - 1;
            cand &= cand - 1u;
            if (at < (uint32_t)LVL_QCAP) lq_pos[at] = (uint16_t)(OFF + ((c / VEC) * THREADS + (int)threadIdx.x) * VEC + (c % VEC));
            else late |= 1u << c;
            at++;
        }
        __syncthreads();
        // (2) the tile's first live token may be the second component of a pair that starts in the previous tile
        if (threadIdx.x == 0) {
            const uint32_t p = (uint32_t)ext[OFF - 1];
            const int f = next_live(ext, OFF - 1);
            if (p != H && f >= 0 && f < OFF + TILE && lvl_find(hkey, hval, pair_key(p, (uint32_t)ext[f]), hash_shift, hash_mask) != EMPTY_KEY)
                ext[f] = (TokT)H;
        }
        {
            const uint32_t nq = lq_n < (uint32_t)LVL_QCAP ? lq_n : (uint32_t)LVL_QCAP;
            for (uint32_t i = threadIdx.x; i < nq; i += THREADS) process((int)lq_pos[i]);
        }
        while (late) {
            const int c = __ffs((int)late) - 1;
            late &= late - 1u;
            process(OFF + ((c / VEC) * THREADS + (int)threadIdx.x) * VEC + (c % VEC));
        }
        __syncthreads();
        if (threadIdx.x == 0) lq_n = 0u;  // (the next tile's pushes come after the barrier at the top of the loop)
        // (3) write back what changed
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const uint4 o = xv[k * THREADS + (int)threadIdx.x];
            if (o.x != v[k].x || o.y != v[k].y || o.z != v[k].z || o.w != v[k].w) gv[k * THREADS + (int)threadIdx.x] = o;
        }
    }
    if (merged) atomicAdd(&sh_n, merged);
    __syncthreads();
    if (threadIdx.x == 0 && sh_n) atomicAdd(nab_out, sh_n);
}

// =========================================================================================
// compaction: squeeze the holes out. Two streaming passes over the sequence: per-tile live counts (tile_count),
// a scan of the counts (tile_scan, one CTA), and the scatter, which compacts a tile in shared memory and writes
// it out as one contiguous, 16-byte-vectorised range (compact_scatter). Both passes read the tile with the same
// interleaved 128-bit loads as merge_kernel.
// =========================================================================================
template <class TokT> __device__ __forceinline__ uint32_t vec_live_count(const uint4& v);
template <> __device__ __forceinline__ uint32_t vec_live_count<uint16_t>(const uint4& v) {
    // halfwords equal to 0xFFFF: exact zero-halfword test of ~w (see vec_mask)
    auto holes = [](uint32_t w) { const uint32_t x = ~w; return (uint32_t)__popc(~(((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) & 0x80008000u); };
    return 8u - holes(v.x) - holes(v.y) - holes(v.z) - holes(v.w);
}
template <> __device__ __forceinline__ uint32_t vec_live_count<uint32_t>(const uint4& v) {
    return (v.x != 0xFFFFFFFFu) + (v.y != 0xFFFFFFFFu) + (v.z != 0xFFFFFFFFu) + (v.w != 0xFFFFFFFFu);
}
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

template <class TokT>
__global__ void __launch_bounds__(THREADS) tile_count_kernel(const TokT* __restrict__ tok, uint32_t* __restrict__ tile_live) {
    __shared__ uint32_t sh[THREADS / 32];
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    const uint4* src = reinterpret_cast<const uint4*>(tok + (size_t)blockIdx.x * TILE);
    uint4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < NV; k++) c += vec_live_count<TokT>(v[k]);
    c = warp_inclusive_scan(c);
    if ((threadIdx.x & 31u) == 31u) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < THREADS / 32; w++) t += sh[w];
        tile_live[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of tile_live[0..ntiles) into tile_off (u64), total -> *total
__global__ void tile_scan_kernel(const uint32_t* __restrict__ tile_live, uint32_t ntiles,
                                 unsigned long long* __restrict__ tile_off, unsigned long long* total) {
    __shared__ unsigned long long part[THREADS];
    const uint32_t per = (ntiles + THREADS - 1) / THREADS;
    const uint32_t lo = threadIdx.x * per;
    unsigned long long s = 0;
    for (uint32_t k = 0; k < per; k++) if (lo + k < ntiles) s += tile_live[lo + k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < THREADS; i++) { unsigned long long v = part[i]; part[i] = acc; acc += v; }
        *total = acc;
    }
    __syncthreads();
    unsigned long long acc = part[threadIdx.x];
    for (uint32_t k = 0; k < per; k++) if (lo + k < ntiles) { tile_off[lo + k] = acc; acc += tile_live[lo + k]; }
}

template <class TokT, class OutT>
__global__ void __launch_bounds__(THREADS) compact_scatter_kernel(const TokT* __restrict__ tok,
                                                                  const unsigned long long* __restrict__ tile_off,
                                                                  OutT* __restrict__ dst) {
    constexpr int VEC = 16 / (int)sizeof(TokT);
    constexpr int NV = TILE / VEC / THREADS;
    constexpr int VO = 16 / (int)sizeof(OutT);  // output slots per 16-byte vector
    constexpr int NW = THREADS / 32;
    __shared__ __align__(16) OutT stage[TILE + VO];
    __shared__ uint32_t wsum[NV * NW];
    const uint32_t H = TokTraits<TokT>::hole;
    const uint4* src = reinterpret_cast<const uint4*>(tok + (size_t)blockIdx.x * TILE);
    uint4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = src[k * THREADS + (int)threadIdx.x];
    // slot order = (k, thread, slot in vector): scan the per-vector live counts in that order
    uint32_t incl[NV], cnt[NV];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        cnt[k] = vec_live_count<TokT>(v[k]);
        incl[k] = warp_inclusive_scan(cnt[k]);
        if (lane == 31u) wsum[k * NW + (int)warp] = incl[k];
    }
    __syncthreads();
    // the tile's output starts at dst + tile_off; stage it with the same offset inside a 16-byte vector
    OutT* out0 = dst + tile_off[blockIdx.x];
    const uint32_t a = (uint32_t)(((size_t)out0 & 15u) / sizeof(OutT));
    uint32_t run = 0, total = 0;
    uint32_t base[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        base[k] = 0;
        for (int w = 0; w < NW; w++) {
            const uint32_t sv = wsum[k * NW + w];
            if ((uint32_t)w == warp) base[k] = run;
            run += sv;
        }
    }
    total = run;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        uint32_t o = a + base[k] + incl[k] - cnt[k];
        uint32_t tv[VEC];
        unpack_vec<TokT>(v[k], tv);
#pragma unroll
        for (int i = 0; i < VEC; i++)
            if (tv[i] != H) stage[o++] = (OutT)tv[i];
    }
    __syncthreads();
    // stage[a, a + total) -> out0[0, total): scalar head up to the first 16-byte boundary, vectors, scalar tail
    const uint32_t end = a + total;
    const uint32_t body_lo = a ? (uint32_t)VO : 0u, body_hi = end / VO * VO;
    if (body_hi > body_lo) {
        const uint4* s4 = reinterpret_cast<const uint4*>(stage);
        uint4* d4 = reinterpret_cast<uint4*>(out0 - a);
        for (uint32_t i = body_lo / VO + threadIdx.x; i < body_hi / VO; i += THREADS) d4[i] = s4[i];
    }
    const uint32_t head_hi = body_hi > body_lo ? body_lo : end;  // no whole vector: everything is "head"
    for (uint32_t i = a + threadIdx.x; i < head_hi; i += THREADS) out0[i - a] = stage[i];
    if (body_hi > body_lo)
        for (uint32_t i = body_hi + threadIdx.x; i < end; i += THREADS) out0[i - a] = stage[i];
}

// =========================================================================================
// pair table rebuild (growth / dropping dead keys)
// =========================================================================================
__global__ void table_rehash_kernel(PairTable src, PairTable dst, StepCtl* ctl, int drop_dead) {
    size_t cap = (size_t)src.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t k = src.keys[i];
        if (k == EMPTY_KEY) continue;
        uint32_t c = src.counts[i];
        if (drop_dead && c == 0) continue;
        uint32_t s = tbl_find_or_insert(dst, k, &ctl->n_inserted);
        if (s == EMPTY_KEY) { atomicOr(&ctl->err, (uint32_t)ERR_TABLE_FULL); continue; }
        dst.counts[s] = c;
    }
}

// =========================================================================================
// decode (src/basic_tokenizer.zig:90-138): ids -> byte lengths -> offsets -> bytes.
// The host builds, once per merge list, for every id: its byte length (0: the id has no valid expansion), the
// offset of its flattened expansion in voc_bytes (VOC_NOT_FLAT for the rare ids whose expansion is too long to
// flatten) and its two components (for those: byte k of an id is found by descending through the merge DAG with
// the lengths, which needs no per-token limit at all).
// decode_len: per-tile byte totals (+ unknown-id detection). decode_scatter: a CTA decodes a tile of ids into shared
// memory and writes it out as one contiguous, 16-byte-vectorised range; tiles that do not fit the stage write
// straight to global memory.
// =========================================================================================
constexpr uint32_t VOC_NOT_FLAT = 0xFFFFFFFFu;
constexpr int DEC_STAGE = 40 * 1024;  // bytes of a tile's output staged in shared memory (+ 16 for alignment)
struct DecVocab {
    const uint32_t* len;    // [65536] bytes of the id's expansion, 0 = invalid id
    const uint32_t* off;    // [65536] offset in `bytes`, VOC_NOT_FLAT if not flattened
    const uint32_t* pair;   // [65536] first | second << 16
    const uint8_t* bytes;
};
__device__ __forceinline__ uint8_t dec_byte_by_descent(const DecVocab& v, uint32_t id, uint32_t k) {
    while (id >= 256u) {
        const uint32_t pr = v.pair[id], f = pr & 0xFFFFu;
        const uint32_t lf = v.len[f];
        if (k < lf) id = f; else { k -= lf; id = pr >> 16; }
    }
    return (uint8_t)id;
}

__global__ void __launch_bounds__(THREADS) decode_len_kernel(const uint16_t* __restrict__ toks, size_t n,
                                                             const uint32_t* __restrict__ voc_len,
                                                             unsigned long long* __restrict__ tile_bytes, StepCtl* ctl) {
    __shared__ unsigned long long sh[THREADS];
    const size_t base = (size_t)blockIdx.x * TILE;
    unsigned long long c = 0;
    for (int k = 0; k < SPT; k++) {
        const size_t i = base + (size_t)k * THREADS + threadIdx.x;  // interleaved: coalesced loads
        if (i < n) {
            const uint32_t l = voc_len[toks[i]];
            if (l == 0) atomicOr(&ctl->err, (uint32_t)ERR_KEY_MISSING);  // unknown id
            c += l;
        }
    }
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int off = THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_bytes[blockIdx.x] = sh[0];
}

// single CTA: exclusive scan of 64-bit per-tile totals
__global__ void tile_scan64_kernel(const unsigned long long* __restrict__ tile_bytes, uint32_t ntiles,
                                   unsigned long long* __restrict__ tile_off, unsigned long long* total) {
    __shared__ unsigned long long part[THREADS];
    const uint32_t per = (ntiles + THREADS - 1) / THREADS;
    const uint32_t lo = threadIdx.x * per;
    unsigned long long s = 0;
    for (uint32_t k = 0; k < per; k++) if (lo + k < ntiles) s += tile_bytes[lo + k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < THREADS; i++) { unsigned long long v = part[i]; part[i] = acc; acc += v; }
        *total = acc;
    }
    __syncthreads();
    unsigned long long acc = part[threadIdx.x];
    for (uint32_t k = 0; k < per; k++) if (lo + k < ntiles) { tile_off[lo + k] = acc; acc += tile_bytes[lo + k]; }
}

__global__ void __launch_bounds__(THREADS) decode_scatter_kernel(const uint16_t* __restrict__ toks, size_t n, DecVocab v,
                                                                 const unsigned long long* __restrict__ tile_off,
                                                                 const unsigned long long* __restrict__ tile_bytes,
                                                                 uint8_t* __restrict__ out, size_t cap) {
    __shared__ __align__(16) uint8_t stage[DEC_STAGE + 16];
    __shared__ unsigned long long sh[THREADS];
    // thread t owns the SPT consecutive ids [base + t * SPT, ...): their bytes are consecutive in the output
    const size_t base = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * SPT;
    unsigned long long c = 0;
    for (int k = 0; k < SPT; k++) { const size_t i = base + k; if (i < n) c += v.len[toks[i]]; }
    // exclusive scan over the threads (64-bit: a tile of long tokens can exceed 4 GB)
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < THREADS; off <<= 1) {
        const unsigned long long add = ((int)threadIdx.x >= off) ? sh[threadIdx.x - off] : 0ull;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    const unsigned long long my_off = sh[threadIdx.x] - c;
    const unsigned long long total = tile_bytes[blockIdx.x];
    const unsigned long long o0 = tile_off[blockIdx.x];
    uint8_t* dst = out + o0;
    const bool staged = total <= (unsigned long long)DEC_STAGE;
    const uint32_t a = (uint32_t)((size_t)dst & 15u);  // stage with the destination's offset inside a 16-byte vector
    uint8_t* w = staged ? stage + a + my_off : dst + my_off;
    unsigned long long room = o0 + my_off < cap ? cap - (o0 + my_off) : 0ull;  // (the caller checked total <= cap; belt and braces)
    for (int k = 0; k < SPT; k++) {
        const size_t i = base + k;
        if (i >= n) break;
        const uint32_t id = toks[i];
        const uint32_t l = v.len[id];
        if ((unsigned long long)l > room) break;
        const uint32_t fo = v.off[id];
        if (fo != VOC_NOT_FLAT) {
            const uint8_t* src = v.bytes + fo;
            for (uint32_t b = 0; b < l; b++) w[b] = src[b];
        } else {
            for (uint32_t b = 0; b < l; b++) w[b] = dec_byte_by_descent(v, id, b);
        }
        w += l;
        room -= l;
    }
    if (!staged) return;
    __syncthreads();
    const uint32_t end = a + (uint32_t)total;
    const uint32_t body_lo = a ? 16u : 0u, body_hi = end / 16u * 16u;
    if (body_hi > body_lo) {
        const uint4* s4 = reinterpret_cast<const uint4*>(stage);
        uint4* d4 = reinterpret_cast<uint4*>(dst - a);
        for (uint32_t i = body_lo / 16u + threadIdx.x; i < body_hi / 16u; i += THREADS) d4[i] = s4[i];
        for (uint32_t i = a + threadIdx.x; i < body_lo; i += THREADS) dst[i - a] = stage[i];
        for (uint32_t i = body_hi + threadIdx.x; i < end; i += THREADS) dst[i - a] = stage[i];
    } else {
        for (uint32_t i = a + threadIdx.x; i < end; i += THREADS) dst[i - a] = stage[i];
    }
}

}  // namespace bpe
