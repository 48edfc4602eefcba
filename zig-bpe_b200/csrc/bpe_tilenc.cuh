// bpe_tilenc.cuh — tile-resident encode (src/basic_tokenizer.zig:71-88 for merge lists a trained tokenizer can
// produce): ONE launch over the text. A CTA keeps a window of the text (a tile of up to TN_TMAX bytes plus TN_M
// bytes of margin on both sides) in shared memory from the first merge to the last, so the text is read from HBM
// once and the ids are written once; the level-scheduled passes re-read the whole sequence once per level.
//
// Why it is exact (regular lists: every new_token >= 256, differs from its components, is mentioned by no earlier
// merge, no pair repeats; the three facts are proved in bpe_segenc.cuh):
//  (1) levels (merge_levels) are a valid clock: "repeat: apply all present pairs of the lowest level" is the
//      reference's loop over the list; a pair that contains a token produced at level l has a level > l, and two
//      different pairs of one level never share a token.
//  (2) a token boundary of the final encoding is never crossed, so
//  (3) stitching: if P encodes bytes [p0,p1) and Q encodes [q0,q1), q0 < p1, and both hold the same token at the same
//      position c, then P[..c) ++ Q[c..) is the encoding of [p0,q1).
// Windows are processed in text order (a ticket gives every CTA the next window); a window publishes the tokens of its
// overlap with the next window, the next window's CTA finds the first common (position, token) there — the seam —
// and emits the previous window's tokens before it and its own after it. A seam that does not exist (margins too
// small for this text, e.g. periodic input), a token longer than 255 bytes or a run of more than 255 equal tokens
// raises *fail and the host runs the level-scheduled passes instead: the result never depends on a heuristic.
//
// Inside a window every live token sits at the slot of its first byte:
//   cell[s] = token | level << 16   level of the pair (token, next live token); TL_INF no merge, TL_DEAD not a token
//                                   start; bit 31 = claimed by a thread of the current round
//   nx[s]   = the id that pair merges into; len[s] = bytes of the token (next token = s + len[s]);
//   back[e] = distance from the last byte e of a token to its first (previous token = (s-1) - back[s-1])
// Rounds go through the levels in ascending order, a RANGE of up to 64 levels at a time: two passes over the window
// count the slots of every level of the range and lay their positions out level by level (`sorted`); pairs that
// come into being inside the range go into small per-level buckets (`late`) by the round that creates them (a bucket
// that overflows ends the range before its level; the next scan finds everything). Most
// levels hold a handful of occurrences per window, so ONE warp works through such sparse levels with nothing but
// __syncwarp between its steps while the other warps of the CTA wait at a barrier; levels with more than TN_SPARSE
// entries (the first few levels, where most of the merging happens) are taken by the whole CTA. Per round and entry:
// (a) an entry that is still valid decides whether it is an occurrence — for first == second pairs only every other
// token of a run is: a walk over the equal-level predecessors gives the offset in the run, even offsets merge
// (aaa -> Xa as the reference) — and claims its slot (atomicCAS, so a slot listed twice is taken once); (b) the
// claimed slots are rewritten (id, length, links); (c) the pairs left and right of every new token are looked up
// (dense table for byte pairs, hashed table else).
// Multi-GPU: a shard sees 512 bytes of its neighbours' text; around every shard boundary B both ranks encode the
// same bridge window [B-192, B+192) and stitch it between their edge tiles, the left rank with a seam <= B, the
// right rank with a seam >= B, so the token stream is cut at B consistently and nothing is exchanged between GPUs.
#pragma once
#include "bpe_common.cuh"
#include "bpe_segenc.cuh"

namespace bpe {

constexpr int TN_M = 64;          // margin on both sides of a tile
constexpr int TN_BR = 192;        // half width of a bridge window
constexpr int TN_ZONE = 256;      // slots a window publishes for its successor (>= TN_BR + TN_M, >= 2 * TN_M)
constexpr int TN_TMAX = 8192;     // largest tile
constexpr int TN_QCAP = 2048;     // positions laid out per range (`sorted`)
constexpr int TN_BUCKET = 64;     // pairs created inside a range that a level of the range can take (`late`), else the range ends there
constexpr int TN_RANGE = 32;      // levels per range
constexpr int TN_LCAP = TN_BUCKET * TN_RANGE;
constexpr int TN_SPARSE = 64;     // a level with at most this many entries is worked through by one warp
constexpr int TN_THREADS = 256;
constexpr uint32_t TL_INF = 0x7FFFu, TL_DEAD = 0x7FFEu, TL_MAXLVL = 0x7FF0u, TL_CLAIM = 0x80000000u;
constexpr uint32_t TN_MIN_SHARD = 1024;  // multi-GPU: smaller shards take the other encoders

struct TileGeom {
    long long n;         // bytes of this shard
    int hl, hr;          // bytes of the neighbouring shards available left / right (0 on a single GPU)
    int bwl, bwr;        // 1: a bridge window before the first / after the last tile
    long long ntile;     // tiles of this shard
    long long te;        // tile length (multiple of 16); the last two tiles share what is left evenly,
    long long last_start;  // so that no tile is shorter than te / 2 - 16 (start of the last tile; 0 when ntile == 1)
};

__host__ __device__ __forceinline__ long long tn_tile_start(const TileGeom& g, long long j) {
    return j >= g.ntile ? g.n : (j + 1 < g.ntile ? j * g.te : g.last_start);
}
inline TileGeom tn_geometry(long long n, int hl, int hr, bool left_text, bool right_text, int tile_max) {
    TileGeom g;
    g.n = n; g.hl = hl; g.hr = hr; g.bwl = left_text ? 1 : 0; g.bwr = right_text ? 1 : 0;
    g.te = tile_max;
    g.ntile = n <= tile_max ? 1 : (n + tile_max - 1) / tile_max;
    g.last_start = 0;
    if (g.ntile >= 2) {
        const long long rest = n - (g.ntile - 2) * g.te;  // te < rest <= 2 * te
        g.last_start = (g.ntile - 2) * g.te + ((rest / 2 + 15) / 16) * 16;
    }
    return g;
}

// window i of a shard: visible range [ws, we) in shard-local positions (may reach into the halos)
__host__ __device__ __forceinline__ void tn_window(const TileGeom& g, long long i, long long* ws, long long* we) {
    const long long vlo = -(long long)g.hl, vhi = g.n + g.hr;
    long long a, b;
    if (g.bwl && i == 0) { a = -TN_BR; b = TN_BR; }
    else if (g.bwr && i == g.bwl + g.ntile) { a = g.n - TN_BR; b = g.n + TN_BR; }
    else {
        const long long j = i - g.bwl;
        a = tn_tile_start(g, j) - TN_M;
        b = tn_tile_start(g, j + 1) + TN_M;
    }
    *ws = a < vlo ? vlo : a;
    *we = b > vhi ? vhi : b;
}

__host__ __device__ constexpr size_t tilenc_smem_bytes(int tile_max, uint32_t max_level) {
    const size_t W = (size_t)tile_max + 2 * TN_M;
    return W * 4 + W * 2 + W + W + (size_t)TN_QCAP * 2 + (size_t)TN_LCAP * 2 + (size_t)TN_RANGE * 4 * 4 + 8 +
           ((size_t)max_level / 32 + 2) * 4 + (size_t)TN_ZONE * 2 + 64;
}

__global__ void __launch_bounds__(TN_THREADS, 3) tilenc_kernel(const uint8_t* __restrict__ text, const uint8_t* __restrict__ halo_l,
                                                            const uint8_t* __restrict__ halo_r, TileGeom g, int tile_max, SegTab T,
                                                            uint32_t max_level, uint16_t* __restrict__ out_slots,
                                                            uint16_t* zone_buf, uint32_t* flags, uint32_t* ticket, uint32_t* fail, uint32_t* dbg) {
    constexpr int NT = TN_THREADS;
    const int WCAP = tile_max + 2 * TN_M;
    uint32_t* cell = bpe_dyn_smem();
    uint16_t* nx = reinterpret_cast<uint16_t*>(cell + WCAP);
    uint8_t* len = reinterpret_cast<uint8_t*>(nx + WCAP);
    uint8_t* back = len + WCAP;
    uint16_t* sorted = reinterpret_cast<uint16_t*>(back + WCAP);  // WCAP is a multiple of 16: still aligned
    uint16_t* late_pos = sorted + TN_QCAP;                              // [TN_RANGE][TN_BUCKET]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(late_pos + TN_LCAP);    // [TN_RANGE] slots per level of the range
    uint32_t* start = cnt + TN_RANGE;                                    // [TN_RANGE + 1] segment of the level in `sorted`
    uint32_t* fill = start + TN_RANGE + 1;                               // [TN_RANGE] placement cursors
    uint32_t* late_cnt = fill + TN_RANGE;                                // [TN_RANGE] entries in the level's bucket (may exceed TN_BUCKET)
    uint32_t* present = late_cnt + TN_RANGE + 1;
    const uint32_t LW = max_level / 32 + 2;
    uint16_t* zoneP = reinterpret_cast<uint16_t*>(present + LW);
    uint32_t* sc = reinterpret_cast<uint32_t*>(zoneP + TN_ZONE);  // scalars
    uint32_t& s_win = sc[0]; uint32_t& late_drop = sc[2]; uint32_t& s_rr = sc[3]; uint32_t& s_fail = sc[4];
    int& s_c = reinterpret_cast<int*>(sc)[5];
    uint32_t& s_hi = sc[6]; uint32_t& s_partial = sc[7]; uint32_t& s_progress = sc[8];
    const int t = (int)threadIdx.x;
    const uint32_t H16 = 0xFFFFu;

    if (t == 0) { s_win = atomicAdd(ticket, 1u); s_fail = 0; }
    for (uint32_t i = (uint32_t)t; i < LW; i += NT) present[i] = 0;
    __syncthreads();
    const long long win = (long long)s_win, nwin = g.bwl + g.ntile + g.bwr;
    long long ws, we;
    tn_window(g, win, &ws, &we);
    const int W = (int)(we - ws);  // 1 <= W <= WCAP

    // ---- bytes -> tokens; levels of the byte pairs ----
    {
        const bool inner = ws >= 0 && we <= g.n;
        for (int s = t; s < W; s += NT) {
            const long long p = ws + s;
            const uint32_t b = inner ? (uint32_t)text[p] : (p < 0 ? (uint32_t)halo_l[g.hl + p] : (p >= g.n ? (uint32_t)halo_r[p - g.n] : (uint32_t)text[p]));
            cell[s] = b;
        }
        __syncthreads();
        for (int s = t; s < W; s += NT) {
            const uint32_t a = cell[s] & 0xFFu;
            uint32_t lv = TL_INF, x = 0;
            if (s + 1 < W) {
                const uint32_t v = T.bp[a | ((cell[s + 1] & 0xFFu) << 8)];
                if (v != SEG_NONE) { lv = v & 0xFFFFu; x = v >> 16; atomicOr(&present[lv >> 5], 1u << (lv & 31u)); }
            }
            nx[s] = (uint16_t)x;
            len[s] = 1;
            back[s] = 0;
            cell[s] = a | (lv << 16);  // the level goes into the high half: neighbours read only the low byte
        }
        __syncthreads();
    }
    auto next_present = [&](uint32_t from) -> uint32_t {  // lowest level >= from that some slot has (had); > max_level: none
        if (from > max_level) return max_level + 1;
        uint32_t w = from >> 5;
        uint32_t bits = present[w] & (0xFFFFFFFFu << (from & 31u));
        while (!bits) { if (++w >= LW) return max_level + 1; bits = present[w]; }
        return w * 32u + (uint32_t)(__ffs((int)bits) - 1);
    };
    // Hashed look-up in two halves, so that a thread can have the first probes of several pairs in flight at once
    // (a round is a chain of dependent latencies; the table lives in L2)
    struct Probe { uint32_t key, s; uint2 e; };
    auto probe_start = [&](Probe& q, uint32_t a, uint32_t b) {
        q.key = pair_key(a, b);
        if ((a | b) < 256u) { q.e.x = q.key; q.e.y = T.bp[a | (b << 8)]; q.s = 0; return; }
        q.s = (q.key * 0x9E3779B1u) >> T.hshift;
        q.e = T.hk[q.s];
    };
    auto probe_finish = [&](Probe& q, uint32_t* x) -> uint32_t {  // level of the pair or TL_INF
        while (q.e.x != q.key) {
            if (q.e.x == EMPTY_KEY) return TL_INF;
            q.s = (q.s + 1u) & T.hmask;
            q.e = T.hk[q.s];
        }
        if (q.e.y == SEG_NONE) return TL_INF;
        *x = q.e.y >> 16;
        return q.e.y & 0xFFFFu;
    };
    // Up to two entries (slots p[0], p[1]) per thread of the group: the three steps of a round, a barrier of the group
    // between them. base / hi: the range whose pairs may join the late buckets.
    auto chunk = [&](uint32_t rr, uint32_t base, uint32_t hi, const int (&p)[2], const bool (&valid)[2], bool whole_cta) {
        // (a) decide + claim
        bool mine[2] = {false, false};
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!valid[e]) continue;
            const uint32_t c = cell[p[e]];
            if ((c >> 16) != rr) continue;  // stale entry, or already claimed through a duplicate
            uint32_t run = 0;
            int q = p[e];
            while (q > 0) {  // offset inside a run of equal pairs (only first == second pairs have equal-level neighbours)
                const int e1 = q - 1, pl = e1 - (int)back[e1];
                if (((cell[pl] >> 16) & 0x7FFFu) != rr) break;
                q = pl;
                if (++run > 255u) { s_fail = 1; break; }
            }
            if (!(run & 1u)) mine[e] = atomicCAS(&cell[p[e]], c, c | TL_CLAIM) == c;  // odd offset: consumed by the occurrence on its left
        }
        if (whole_cta) __syncthreads(); else __syncwarp();
        // (b) rewrite the claimed slots
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!mine[e]) continue;
            const int lp = len[p[e]], q = p[e] + lp, lq = len[q], nl = lp + lq;
            if (nl > 255) { s_fail = 1; mine[e] = false; continue; }
            cell[p[e]] = (uint32_t)nx[p[e]] | (TL_INF << 16);
            len[p[e]] = (uint8_t)nl;
            cell[q] = TL_DEAD << 16;
            back[q + lq - 1] = (uint8_t)(nl - 1);
            s_progress = 1;
        }
        if (whole_cta) __syncthreads(); else __syncwarp();
        // (c) levels of the pairs next to the new tokens: all first probes go out before any of them is waited for
        auto set_level = [&](int s, uint32_t tok, uint32_t lv, uint32_t x) {
            const uint32_t old = (cell[s] >> 16) & 0x7FFFu;
            cell[s] = tok | (lv << 16);
            nx[s] = (uint16_t)x;
            if (lv == TL_INF) return;
            if (lv <= rr) { s_fail = 1; return; }  // cannot happen for a regular list (fact 1)
            atomicOr(&present[lv >> 5], 1u << (lv & 31u));
            if (lv < hi && lv != old) {
                const uint32_t at = atomicAdd(&late_cnt[lv - base], 1u);
                if (at < (uint32_t)TN_BUCKET) late_pos[(lv - base) * TN_BUCKET + at] = (uint16_t)s;
                else atomicMin(&late_drop, lv);
            }
        };
        Probe pr[4];
        int slot[4];
        uint32_t tokv[4];
        bool on[4] = {false, false, false, false};
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!mine[e]) continue;
            const uint32_t X = cell[p[e]] & 0xFFFFu;
            const int n = p[e] + (int)len[p[e]];
            slot[2 * e] = p[e]; tokv[2 * e] = X;
            if (n < W) { on[2 * e] = true; probe_start(pr[2 * e], X, cell[n] & 0xFFFFu); }
            else set_level(p[e], X, TL_INF, 0u);
            if (p[e] > 0) {
                const int e1 = p[e] - 1, pl = e1 - (int)back[e1];
                const uint32_t tp = cell[pl] & 0xFFFFu;
                slot[2 * e + 1] = pl; tokv[2 * e + 1] = tp;
                on[2 * e + 1] = true;
                probe_start(pr[2 * e + 1], tp, X);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!on[i]) continue;
            uint32_t x = 0;
            const uint32_t lv = probe_finish(pr[i], &x);
            set_level(slot[i], tokv[i], lv, x);
        }
        if (whole_cta) __syncthreads(); else __syncwarp();
    };
    // all entries of level rr: its segment of `sorted`, then its late bucket; G threads, two entries each per chunk
    auto round = [&](uint32_t rr, uint32_t base, uint32_t hi, bool whole_cta) {
        const int G = whole_cta ? NT : 32, g = whole_cta ? t : (t & 31);
        const uint32_t li = rr - base;
        const uint32_t n_seg = cnt[li] < (uint32_t)TN_QCAP - start[li] ? cnt[li] : (uint32_t)TN_QCAP - start[li], s0 = start[li];
        const uint32_t n_late = late_cnt[li] < (uint32_t)TN_BUCKET ? late_cnt[li] : (uint32_t)TN_BUCKET;
        const uint32_t n_all = n_seg + n_late;
        for (uint32_t k = 0; k < n_all; k += 2u * (uint32_t)G) {
            int p[2];
            bool valid[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const uint32_t idx = k + (uint32_t)e * (uint32_t)G + (uint32_t)g;
                valid[e] = idx < n_all;
                p[e] = !valid[e] ? 0 : (idx < n_seg ? (int)sorted[s0 + idx] : (int)late_pos[li * TN_BUCKET + (idx - n_seg)]);
            }
            chunk(rr, base, hi, p, valid, whole_cta);
        }
    };

    // ---- rounds ----
    uint32_t cur = 0;  // all levels <= cur are done
    while (true) {
        const uint32_t base = next_present(cur + 1u);
        if (base > max_level || s_fail) break;
        uint32_t hi = base + (uint32_t)TN_RANGE;
        if (hi > max_level + 1u) hi = max_level + 1u;
        const uint32_t nl = hi - base;
        if (dbg && t == 0) atomicAdd(&dbg[0], 1u);
        __syncthreads();  // everybody has read the state of the previous range
        for (uint32_t i = (uint32_t)t; i < (uint32_t)TN_RANGE; i += NT) { cnt[i] = 0; fill[i] = 0; late_cnt[i] = 0; }
        if (t == 0) { late_drop = 0xFFFFu; s_progress = 0; }
        __syncthreads();
        const int W4 = W / 4;
        // After the first ranges most slots are dead or carry no pair: a group of four cells whose lowest level (packed
        // 16-bit minimum of cell - base << 16; the token halves never borrow) lies outside the range is skipped in 8
        // instructions. (No cell is claimed between rounds, so bit 31 is clear.)
        const uint32_t base_hi = base << 16;
        auto group_min = [&](const uint4& c4) -> uint32_t {
            return __vminu2(__vimin3_u16x2(c4.x - base_hi, c4.y - base_hi, c4.z - base_hi), c4.w - base_hi) >> 16;
        };
        for (int s4 = t; s4 < W4; s4 += NT) {  // pass 1: slots per level (four slots per load)
            const uint4 c4 = reinterpret_cast<const uint4*>(cell)[s4];
            if (group_min(c4) >= nl) continue;
            const uint32_t l0 = ((c4.x >> 16) & 0x7FFFu) - base, l1 = ((c4.y >> 16) & 0x7FFFu) - base, l2 = ((c4.z >> 16) & 0x7FFFu) - base,
                           l3 = ((c4.w >> 16) & 0x7FFFu) - base;
            if (l0 < nl) atomicAdd(&cnt[l0], 1u);
            if (l1 < nl) atomicAdd(&cnt[l1], 1u);
            if (l2 < nl) atomicAdd(&cnt[l2], 1u);
            if (l3 < nl) atomicAdd(&cnt[l3], 1u);
        }
        for (int s = W4 * 4 + t; s < W; s += NT) {
            const uint32_t lv = ((cell[s] >> 16) & 0x7FFFu) - base;
            if (lv < nl) atomicAdd(&cnt[lv], 1u);
        }
        __syncthreads();
        if (t == 0) {
            // segments; the range ends before the first level that does not fit (a first level that does not fit alone
            // is taken partially: any subset of a level's occurrences may go first, the next scan finds the rest)
            uint32_t acc = 0, h = hi, partial = 0;
            for (uint32_t i = 0; i < nl; i++) {
                start[i] = acc;
                if (acc + cnt[i] > (uint32_t)TN_QCAP) {
                    if (i == 0) { partial = 1; h = base + 1u; } else h = base + i;
                    break;
                }
                acc += cnt[i];
            }
            s_hi = h;
            s_partial = partial;
        }
        __syncthreads();
        hi = s_hi;
        const bool partial = s_partial != 0u;
        const uint32_t nh = hi - base;
        auto place = [&](uint32_t li, int s) {
            const uint32_t at = start[li] + atomicAdd(&fill[li], 1u);
            if (at < (uint32_t)TN_QCAP) sorted[at] = (uint16_t)s;
        };
        for (int s4 = t; s4 < W4; s4 += NT) {  // pass 2: positions, level by level
            const uint4 c4 = reinterpret_cast<const uint4*>(cell)[s4];
            if (group_min(c4) >= nh) continue;
            const uint32_t l0 = ((c4.x >> 16) & 0x7FFFu) - base, l1 = ((c4.y >> 16) & 0x7FFFu) - base, l2 = ((c4.z >> 16) & 0x7FFFu) - base,
                           l3 = ((c4.w >> 16) & 0x7FFFu) - base;
            if (l0 < nh) place(l0, s4 * 4);
            if (l1 < nh) place(l1, s4 * 4 + 1);
            if (l2 < nh) place(l2, s4 * 4 + 2);
            if (l3 < nh) place(l3, s4 * 4 + 3);
        }
        for (int s = W4 * 4 + t; s < W; s += NT) {
            const uint32_t lv = ((cell[s] >> 16) & 0x7FFFu) - base;
            if (lv < nh) place(lv, s);
        }
        __syncthreads();
        uint32_t rr = base;
        while (true) {
            if (t < 32) {
                // sparse levels: one warp, no CTA barrier
                while (rr < hi && rr < late_drop && !s_fail && cnt[rr - base] + late_cnt[rr - base] <= (uint32_t)TN_SPARSE) {
                    if (cnt[rr - base] + late_cnt[rr - base]) { round(rr, base, hi, false); if (dbg && t == 0) atomicAdd(&dbg[1], 1u); }
                    rr++;
                }
                if (t == 0) s_rr = rr;
            }
            __syncthreads();
            rr = s_rr;
            if (rr >= hi || rr >= late_drop || s_fail) break;
            round(rr, base, hi, true);  // a dense level: the whole CTA (ends with a barrier)
            if (dbg && t == 0) atomicAdd(&dbg[partial ? 3 : 2], 1u);
            rr++;
        }
        if (partial) {
            // the same level again; a partial round that merged nothing (all its entries were odd offsets of runs)
            // would repeat forever: give up on this path
            if (!s_progress) s_fail = 1;
        } else {
            const uint32_t lim = hi < late_drop ? hi : late_drop;
            cur = lim - 1u;
        }
    }
    __syncthreads();

    // ---- publish my overlap with the next window, then stitch with the previous one ----
    if (win + 1 < nwin) {
        long long nws, nwe;
        tn_window(g, win + 1, &nws, &nwe);
        uint16_t* zo = zone_buf + (size_t)win * TN_ZONE;
        const int z0 = (int)(nws - ws), zl = (int)(we - nws);  // zl <= TN_ZONE
        for (int j = t; j < zl && j < TN_ZONE; j += NT) {
            const uint32_t c = cell[z0 + j];
            zo[j] = (((c >> 16) & 0x7FFFu) == TL_DEAD) ? (uint16_t)H16 : (uint16_t)(c & 0xFFFFu);
        }
        __threadfence();
        __syncthreads();
        if (t == 0) { atomicExch(&flags[win], 1u); }
    }
    int cstar = 0;  // relative to ws: tokens of this window are emitted from here on
    int zlen = 0;   // length of the overlap with the previous window
    if (win > 0) {
        long long pws, pwe;
        tn_window(g, win - 1, &pws, &pwe);
        zlen = (int)(pwe - ws);
        if (zlen > TN_ZONE) zlen = TN_ZONE;
        if (t == 0) {
            uint32_t spins = 0;
            while (((volatile uint32_t*)flags)[win - 1] == 0u) { if (++spins > (1u << 26)) { s_fail = 1; break; } }
        }
        __syncthreads();
        __threadfence();
        const uint16_t* zi = zone_buf + (size_t)(win - 1) * TN_ZONE;
        for (int j = t; j < zlen; j += NT) zoneP[j] = ((volatile const uint16_t*)zi)[j];
        __syncthreads();
        if (t == 0) {
            // first common (position, token); the bridge windows cut the stream at the shard boundary
            int lo = 0, hi_ok = zlen;
            if (g.bwl && win == 1) lo = (int)(0 - ws);                       // seam at or after the boundary (local 0)
            if (g.bwr && win == nwin - 1) hi_ok = (int)(g.n - ws) + 1;       // seam at or before the boundary (local n)
            int c = -1;
            for (int j = lo < 0 ? 0 : lo; j < zlen && j < hi_ok; j++) {
                const uint32_t p = zoneP[j];
                if (p == H16) continue;
                const uint32_t q = cell[j];
                if (((q >> 16) & 0x7FFFu) != TL_DEAD && (q & 0xFFFFu) == p) { c = j; break; }
            }
            if (c < 0) { s_fail = 1; c = zlen < hi_ok ? zlen : hi_ok; }
            s_c = c;
        }
        __syncthreads();
        cstar = s_c;
    }
    if (s_fail && t == 0) atomicAdd(fail, 1u);
    // ---- emit [ws, start of the next window) clipped to the shard: the previous window's tokens before the seam,
    // mine after it; slots that start no token become holes ----
    long long ehi = we;
    if (win + 1 < nwin) { long long nws, nwe; tn_window(g, win + 1, &nws, &nwe); ehi = nws; }
    long long elo = ws;
    if (elo < 0) elo = 0;
    if (ehi > g.n) ehi = g.n;
    uint16_t* tile = nx;  // the pair targets are dead: their memory becomes the output tile
    __syncthreads();
    for (int s = t; s < W; s += NT) {
        uint32_t v;
        if (s < cstar) v = s < zlen ? (uint32_t)zoneP[s] : H16;
        else { const uint32_t c = cell[s]; v = (((c >> 16) & 0x7FFFu) == TL_DEAD) ? H16 : (c & 0xFFFFu); }
        tile[s] = (uint16_t)v;
    }
    __syncthreads();
    if (ehi > elo) {
        const int r0 = (int)(elo - ws), cntw = (int)(ehi - elo);
        uint16_t* dst = out_slots + elo;
        // 16-byte vectors where source and destination are aligned alike, scalars elsewhere
        const int mis = (int)((((size_t)dst) & 15u) / 2u);
        const int head = mis ? 8 - mis : 0;
        if (((r0 + head) & 7) == 0 && cntw > head + 8) {
            for (int i = t; i < head; i += NT) dst[i] = tile[r0 + i];
            const int nv = (cntw - head) / 8;
            const uint4* s4 = reinterpret_cast<const uint4*>(tile + r0 + head);
            uint4* d4 = reinterpret_cast<uint4*>(dst + head);
            for (int i = t; i < nv; i += NT) d4[i] = s4[i];
            for (int i = head + nv * 8 + t; i < cntw; i += NT) dst[i] = tile[r0 + i];
        } else {
            for (int i = t; i < cntw; i += NT) dst[i] = tile[r0 + i];
        }
    }
}

}  // namespace bpe
