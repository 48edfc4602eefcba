// bpe_segenc.cuh — segment-resident encode (src/basic_tokenizer.zig:71-88 for merge lists a trained
// tokenizer can produce), one launch over the text instead of one HBM sweep per level.
//
// Why it is exact. Call a merge list *regular* when every new_token is >= 256, differs from its two
// components, is mentioned by no earlier merge, and no pair repeats. For such a list:
//  (1) the level of a merge (build_encode_schedule: 1 + the largest level of the earlier merges it does not
//      commute with) is a valid clock: merges of one level commute, and a pair that contains a token
//      produced at level l has a level > l. "Repeat: apply all present pairs of the lowest level" is
//      therefore the reference's loop over the list.
//  (2) Cuts are never crossed: if position c is a token boundary of the final encoding of a text, the two
//      sides evolve exactly as if they were encoded alone (a pass only joins neighbours; the greedy
//      left-to-right rule for first == second restarts at a boundary). Hence every token-aligned
//      sub-sequence of an encoding is the encoding of its own bytes, and
//  (3) Stitching: if P is the encoding of bytes [p0,p1) and Q of [q0,q1), q0 < p1, and both contain the
//      same token at the same position [c,d), then P[..c) ++ Q[c..) is the encoding of [p0,q1): up to
//      the first pass that would join across c both sides evolve as in P and Q, where that pass does not
//      join them (run parity included, because P has a boundary wherever Q's window starts a run).
// So every thread encodes a small window (core of C bytes + M bytes of margin on both sides) on its
// own in shared memory, neighbouring windows look for a common (position, token) in their 2M-byte overlap
// — `seam` — and each window emits the tokens between its two seams. A window pair without a common token
// (margin too small for this text, e.g. periodic input) raises *fail and the host runs the
// level-scheduled passes instead: the result never depends on the margin being "large enough".
// Cores lie on a grid of global byte offsets (multiples of C), so two GPUs that hold neighbouring shards
// (plus 512 bytes of each other's text) compute identical windows around the shard boundary and cut the
// token stream at the same place.
#pragma once
#include "bpe_common.cuh"

namespace bpe {

constexpr uint32_t SEG_NONE = 0xFFFFFFFFu;
constexpr uint32_t LV_INF = 0xFFFFu;    // no merge for this pair (levels are <= 65,280)
constexpr uint32_t LV_DIRTY = 0xFFFEu;  // a neighbour changed: look the pair up again
constexpr uint32_t LV_DEAD = 0xFFFDu;   // slot consumed by the merge on its left
constexpr int SEG_HALO = 512;           // bytes of neighbouring shards a rank needs (>= 3C + M)

struct SegTab {
    const uint32_t* bp;   // [65536] byte pair first | second << 8 -> level | new_token << 16 (SEG_NONE: no merge)
    const uint2* hk;      // open addressing, .x = pair_key (EMPTY_KEY = free), .y = level | new_token << 16
    uint32_t hmask, hshift;
    const uint16_t* len;  // [65536] bytes every id expands to (clamped to 65535)
};

__device__ __forceinline__ uint32_t seg_lookup(const SegTab& T, uint32_t a, uint32_t b) {
    if ((a | b) < 256u) return T.bp[a | (b << 8)];
    const uint32_t key = pair_key(a, b);
    uint32_t s = (key * 0x9E3779B1u) >> T.hshift;
    while (true) {
        const uint2 e = T.hk[s];
        if (e.x == key) return e.y;
        if (e.x == EMPTY_KEY) return SEG_NONE;
        s = (s + 1u) & T.hmask;
    }
}

template <int C, int M, int NT> __host__ __device__ constexpr size_t segenc_smem_bytes() {
    // columns [L][NT] u32 + max(hit/dirty list [L][NT] u8, output tile built in its place) + per-thread seam scratch
    const size_t lst = (size_t)(C + 2 * M) * NT, tile = (size_t)(NT - 2) * C * 2;
    return (size_t)(C + 2 * M) * NT * 4 + ((lst > tile ? lst : tile) + 15) / 16 * 16 + (size_t)NT * 3 * 4 + 64;
}

// Geometry: thread t of a CTA owns core k = k_tile + t - 1 (threads 0 and NT-1 repeat the neighbouring
// CTAs' edge windows, so that every seam is decided from the same data on both sides without any
// exchange between CTAs). All positions are shard-local (may be negative / >= n inside the halos).
template <int C, int M, int NT>
__global__ void __launch_bounds__(NT) segenc_kernel(const uint8_t* __restrict__ text, long long n,
                                                    const uint8_t* __restrict__ halo_l, int hl,
                                                    const uint8_t* __restrict__ halo_r, int hr, long long g_lo,
                                                    long long k_first, long long k_last, SegTab T,
                                                    uint16_t* __restrict__ out_slots, uint32_t* fail) {
    constexpr int L = C + 2 * M;
    static_assert(L <= 256, "positions inside a window are kept in bytes");
    static_assert(C >= 2 * M, "the seam zones of one window must not overlap");
    static_assert(3 * C + M <= SEG_HALO, "halo too small for this geometry");
    static_assert(NT % 32 == 0, "whole warps");
    constexpr int TILE_SLOTS = (NT - 2) * C;
    constexpr int REGION = ((L * NT > TILE_SLOTS * 2 ? L * NT : TILE_SLOTS * 2) + 15) / 16 * 16;  // lists, then the output tile
    uint32_t* col = bpe_dyn_smem();                         // [L][NT] token | level << 16
    uint8_t* lst = reinterpret_cast<uint8_t*>(col + L * NT);  // [L][NT] hits from the front, dirty slots from the back
    uint16_t* tile = reinterpret_cast<uint16_t*>(lst);      // after the rounds: the CTA's output tile
    int* s_cnt = reinterpret_cast<int*>(lst + REGION);
    int* s_ws = s_cnt + NT;
    int* s_c = s_ws + NT;
    const int t = (int)threadIdx.x;
#define COL(s) col[(s) * NT + t]
#define LST(i) lst[(i) * NT + t]
    const long long k_tile = k_first + (long long)blockIdx.x * (NT - 2);  // core of thread 1
    const long long origin = (k_tile - 1) * C - g_lo - M;                 // local position of relative offset 0
    const long long vlo = -(long long)hl, vhi = n + hr;                   // bytes this rank can see
    const long long k = k_tile + t - 1;
    const bool active_core = k >= k_first && k <= k_last;
    // window, relative to origin
    int ws = t * C, we = t * C + L;
    if (origin + ws < vlo) ws = (int)(vlo - origin);
    if (origin + we > vhi) we = (int)(vhi - origin);
    const bool active = active_core && we > ws;

    int cnt = 0;
    if (active) {
        // ---- initial tokens + levels of the byte pairs (bytes straight from global memory: neighbouring lanes
        // read neighbouring 64-byte pieces, every line is used in full through L1) ----
        cnt = we - ws;
        uint32_t m = LV_INF;
        int nh = 0;
        const long long p0 = origin + ws;
        const bool inner = p0 >= 0 && p0 + cnt <= n;
        auto fetch = [&](long long p) -> uint32_t {
            return p < 0 ? halo_l[hl + p] : (p >= n ? halo_r[p - n] : text[p]);
        };
        uint32_t prev = inner ? (uint32_t)text[p0] : fetch(p0);
        for (int s = 0; s < cnt; s++) {
            uint32_t lv = LV_INF, nxt = 0;
            if (s + 1 < cnt) {
                nxt = inner ? (uint32_t)text[p0 + s + 1] : fetch(p0 + s + 1);
                lv = T.bp[prev | (nxt << 8)] & 0xFFFFu;
            }
            COL(s) = prev | (lv << 16);
            if (lv < m) { m = lv; nh = 0; }
            if (lv == m && lv != LV_INF) LST(nh++) = (uint8_t)s;
            prev = nxt;
        }
        // ---- rounds: all pairs of the window's lowest level, then the next level, ... ----
        while (m != LV_INF) {
            // B: the hits, left to right. Two hits at adjacent positions can only be the same pair (A,A)
            // (pairs of one level never share a token otherwise): the left one wins, as in the reference.
            int last = -2;
            for (int q = 0; q < nh; q++) {
                const int p = LST(q);
                if (p == last + 1) continue;
                last = p;
                const uint32_t a = COL(p) & 0xFFFFu, b = COL(p + 1) & 0xFFFFu;
                const uint32_t x = seg_lookup(T, a, b) >> 16;
                COL(p) = x | (LV_DIRTY << 16);
                COL(p + 1) = LV_DEAD << 16;
                if (p > 0) {
                    const uint32_t u = COL(p - 1);
                    if ((u >> 16) != LV_DEAD) COL(p - 1) = (u & 0xFFFFu) | (LV_DIRTY << 16);
                }
            }
            // C: squeeze the dead slots out; list the slots whose pair changed; hits of the next level among the rest
            int w = 0, nd = 0;
            uint32_t m2 = LV_INF;
            nh = 0;
            {
                const uint32_t* rp = col + t;
                uint32_t* wp = col + t;
                uint8_t* hp = lst + t;
                uint8_t* dp = lst + (L - 1) * NT + t;
                for (int s = 0; s < cnt; s++, rp += NT) {
                    const uint32_t v = *rp, lv = v >> 16;
                    if (lv == LV_DEAD) continue;
                    *wp = v;
                    wp += NT;
                    if (lv < LV_DEAD) {
                        if (lv < m2) { m2 = lv; nh = 0; hp = lst + t; }
                        if (lv == m2) { *hp = (uint8_t)w; hp += NT; nh++; }
                    } else if (lv == LV_DIRTY) { *dp = (uint8_t)w; dp -= NT; nd++; }
                    w++;
                }
            }
            cnt = w;
            // D: levels of the changed pairs
            for (int q = 0; q < nd; q++) {
                const int p = LST(L - 1 - q);
                const uint32_t a = COL(p) & 0xFFFFu;
                uint32_t lv = LV_INF;
                if (p + 1 < cnt) lv = seg_lookup(T, a, COL(p + 1) & 0xFFFFu) & 0xFFFFu;
                COL(p) = a | (lv << 16);
                if (lv < m2) { m2 = lv; nh = 0; }
                if (lv == m2 && lv != LV_INF) LST(nh++) = (uint8_t)p;
            }
            m = m2;
        }
    }
    s_cnt[t] = active ? cnt : 0;
    s_ws[t] = ws;
    __syncthreads();  // every window is final

    // ---- seam with the window on my right: first common (position, token) in the overlap ----
    int cstar = we;
    if (active && t + 1 < NT && s_cnt[t + 1] > 0) {
        const int wsq = s_ws[t + 1], cntq = s_cnt[t + 1];
        int i = cnt, pos = we;  // walk back to my first token that starts inside the neighbour's window
        while (i > 0) {
            const int l = (int)T.len[COL(i - 1) & 0xFFFFu];
            if (pos - l < wsq) break;
            pos -= l;
            i--;
        }
        int j = 0, posq = wsq;
        cstar = -1;
        while (i < cnt && j < cntq) {
            const uint32_t tp = COL(i) & 0xFFFFu, tq = col[j * NT + t + 1] & 0xFFFFu;
            if (pos == posq) {
                if (tp == tq) { cstar = pos; break; }
                pos += (int)T.len[tp]; i++;
                posq += (int)T.len[tq]; j++;
            } else if (pos < posq) { pos += (int)T.len[tp]; i++; }
            else { posq += (int)T.len[tq]; j++; }
        }
        if (cstar < 0) { atomicAdd(fail, 1u); cstar = (t + 1) * C + M; }
    }
    s_c[t] = cstar;
    __syncthreads();  // the lists are dead: their memory becomes the output tile
    // the output tile: cores of threads 1..NT-2, clipped to the shard
    long long tlo = origin + C + M, thi = tlo + TILE_SLOTS;
    if (tlo < 0) tlo = 0;
    if (thi > n) thi = n;
    const int rlo = (int)(tlo - origin), rhi = (int)(thi - origin);  // relative
    const uint16_t H16 = 0xFFFFu;
    for (int i = t; i < TILE_SLOTS; i += NT) tile[i] = H16;
    __syncthreads();
    if (active && rhi > rlo) {
        int lo = ws, hi = s_c[t];
        if (t > 0 && s_cnt[t - 1] > 0) lo = s_c[t - 1];
        if (lo < rlo) lo = rlo;
        if (hi > rhi) hi = rhi;
        int pos = ws;
        for (int i = 0; i < cnt && pos < hi; i++) {
            const uint32_t tk = COL(i) & 0xFFFFu;
            if (pos >= lo) tile[pos - rlo] = (uint16_t)tk;
            pos += (int)T.len[tk];
        }
    }
    __syncthreads();
    if (rhi > rlo) {
        uint16_t* dst = out_slots + tlo;
        const int cntw = rhi - rlo;
        if ((((size_t)dst) & 15) == 0) {
            const uint4* src4 = reinterpret_cast<const uint4*>(tile);
            uint4* dst4 = reinterpret_cast<uint4*>(dst);
            const int nv = cntw / 8;
            for (int i = t; i < nv; i += NT) dst4[i] = src4[i];
            for (int i = nv * 8 + t; i < cntw; i += NT) dst[i] = tile[i];
        } else {
            for (int i = t; i < cntw; i += NT) dst[i] = tile[i];
        }
    }
#undef COL
#undef LST
}

}  // namespace bpe
