// bpe_groupenc.cuh — level passes of encode (src/basic_tokenizer.zig:71-88) fused in groups: up to GRP_MAXLV consecutive
// levels of the schedule are applied during ONE residency of a tile in shared memory (option "encode_fuse", off by default:
// written after the round's last GPU lease, verified against the oracle under the CPU emulation only, not yet timed).
//
// Why: a level pass executes ~1 warp instruction per slot, a third of it the per-slot role look-up that finds the candidates
// of that one level, and C3's ~190 passes re-read the sequence ~190 times. Here a slot's role byte
// says in which of the group's levels its token is a first component: one look-up per slot per GROUP puts the slot into the
// queues of those levels, and a level then costs only its own queue. A token that a merge of the group produces enters the
// queues of the later levels it can start a pair in, so the queues follow the tokens as they change.
//
// Why it is exact. Pairs of one level never share a token (merge_levels), so whether (t_i, t_i+1) merges at a level depends
// on these two tokens alone. A tile is staged with its GRP_H nearest live tokens on either side (deep_halo_kernel, the
// sequence as the previous pass left it), laid out next to it without holes, and they go through the levels like the tile's
// own tokens. What lies beyond them is unknown: at every level the outermost halo token may be taken by a neighbour the CTA
// does not see. The CTA therefore keeps, per side, the index up to which the stage may differ from the truth ("taint"):
// after every level it moves on to the next live token (if the outermost trusted token P pairs with its inner neighbour at
// this level, that merge is certain — no pair of the same level can take P from the other side — otherwise P may be gone;
// either way everything up to and including P's slot, and up to the next live token if P's partner slot was emptied by a
// tainted token, is no longer trusted). Halo tokens merge with each other, so the number of levels a halo of GRP_H tokens
// covers depends on the text: when the taint reaches a tile slot the CTA raises `fail`. The pass writes OUT OF PLACE (the
// sequence's second buffer), so a failed group costs one wasted pass: the host keeps the input buffer, runs the group's
// levels one by one and halves the group size. A tile at an end of the sequence has nothing unknown beyond its halo.
// Both tiles next to a boundary simulate the same merges around it; each writes only its own slots (the slot of a first
// component becomes the new id, the slot of a second component a hole), so the two agree without talking to each other.
// Only regular level steps are grouped (u16 ids below LVL_BYTE_IDS_MAX, first != second, single GPU); everything else
// takes the existing passes.
#pragma once
#include "bpe_kernels.cuh"

namespace bpe {

constexpr int GRP_MAXLV = 8;                // levels per group (one bit each in a role byte)
constexpr int GRP_H = 30;                   // live tokens staged on either side of a tile
constexpr int GRP_OFF = 32;                 // first tile slot in the stage (>= GRP_H, 16-byte aligned)
constexpr int GRP_EXT = GRP_OFF + TILE + 32;
#ifdef BPE_EMUL
constexpr int GRP_QCAP = 12;                // (emulation: tiny queues, so that the rescan of an overflowed level is exercised)
#else
constexpr int GRP_QCAP = 1024;              // queued slots per level and tile; a level that overflows is rescanned instead
#endif
constexpr int GRP_HASH_MAX = 8192;          // hash slots of a group (power of two >= 2 * pairs)
constexpr int GRP_PAIRS_MAX = GRP_HASH_MAX / 2;

// l[0] / r[0] nearest to the tile; 0xFFFF: no such token. lend / rend: the sequence ends inside the halo (nothing unknown beyond)
struct DeepHalo { uint16_t l[GRP_H]; uint16_t r[GRP_H]; uint16_t lend, rend; };
struct GroupResult { uint32_t merged, fail; };
struct GroupDesc { uint32_t nlev; uint32_t end[GRP_MAXLV]; };  // level k of the group = ents[end[k-1], end[k])

__host__ __device__ constexpr size_t group_smem_bytes(int hash_slots, int role_bytes) {
    return (size_t)GRP_EXT * 2 + (size_t)GRP_EXT + (size_t)role_bytes + (size_t)hash_slots * 4 + (size_t)hash_slots * 2 + (size_t)hash_slots +
           (size_t)GRP_MAXLV * GRP_QCAP * 2 + 64;
}

// one thread per tile: the GRP_H nearest live tokens before and after it
__global__ void __launch_bounds__(HALO_THREADS) deep_halo_kernel(const uint16_t* __restrict__ tok, size_t n_slots, uint32_t ntiles,
                                                                  DeepHalo* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    DeepHalo h;
    // (16-byte loads: ~5 dependent round trips per side instead of ~35; positions stay multiples of 8)
    const uint4* tv = reinterpret_cast<const uint4*>(tok);
    int k = 0;
    size_t p = (size_t)t * TILE;
    while (k < GRP_H && p > 0) {
        p -= 8;
        uint32_t w[8];
        unpack_vec<uint16_t>(tv[p / 8], w);
        for (int i = 7; i >= 0 && k < GRP_H; i--) if (w[i] != 0xFFFFu) h.l[k++] = (uint16_t)w[i];
    }
    h.lend = (k < GRP_H) ? 1 : 0;  // ran into the start of the sequence (a full halo may still have live tokens before it)
    for (; k < GRP_H; k++) h.l[k] = 0xFFFFu;
    k = 0;
    p = ((size_t)t + 1) * TILE;
    while (k < GRP_H && p < n_slots) {
        uint32_t w[8];
        unpack_vec<uint16_t>(tv[p / 8], w);
        p += 8;
        for (int i = 0; i < 8 && k < GRP_H; i++) if (w[i] != 0xFFFFu) h.r[k++] = (uint16_t)w[i];
    }
    h.rend = (k < GRP_H) ? 1 : 0;
    for (; k < GRP_H; k++) h.r[k] = 0xFFFFu;
    out[t] = h;
}

__global__ void __launch_bounds__(THREADS, 3) group_kernel(const uint16_t* __restrict__ tok, uint16_t* __restrict__ tok_out,
                                                           const DeepHalo* __restrict__ halo, const LevelEntry* __restrict__ ents,
                                                           GroupDesc gd, GroupResult* res, int backwards, uint32_t ntiles,
                                                           uint32_t hash_log2, uint32_t role_bytes) {
    unsigned char* raw = reinterpret_cast<unsigned char*>(bpe_dyn_smem());
    uint16_t* ext = reinterpret_cast<uint16_t*>(raw);                                   // [GRP_EXT]
    const uint32_t hash_slots = 1u << hash_log2, hash_mask = hash_slots - 1u, hash_shift = 32u - hash_log2;
    uint32_t* hkey = reinterpret_cast<uint32_t*>(raw + (size_t)GRP_EXT * 2);           // [hash_slots] pair key
    uint16_t* hval = reinterpret_cast<uint16_t*>(hkey + hash_slots);                   // [hash_slots] new id
    uint16_t* queue = hval + hash_slots;                                               // [GRP_MAXLV][GRP_QCAP] stage indices
    unsigned char* hlev = reinterpret_cast<unsigned char*>(queue + GRP_MAXLV * GRP_QCAP);  // [hash_slots] level inside the group
    unsigned char* role8 = hlev + hash_slots;                                          // [role_bytes] bit k: first component in level k
    unsigned char* qmask = role8 + role_bytes;                                         // [GRP_EXT] bit k: the slot is in queue k already
    __shared__ uint32_t qn[GRP_MAXLV];
    __shared__ uint32_t qover;  // bit k: queue k overflowed on this tile (the level is rescanned)
    __shared__ uint32_t sh_n;
    __shared__ int taint_l, taint_r;  // stage indices <= taint_l / >= taint_r may differ from the truth
    __shared__ uint32_t sh_fail;
    constexpr int VEC = 8;
    constexpr int NV = TILE / VEC / THREADS;
    const uint16_t H = 0xFFFFu;
    const int nlev = (int)gd.nlev;

    // ---- the group's tables, once per CTA ----
    for (uint32_t i = threadIdx.x; i < role_bytes / 4u; i += THREADS) reinterpret_cast<uint32_t*>(role8)[i] = 0u;
    for (uint32_t i = threadIdx.x; i < hash_slots; i += THREADS) hkey[i] = EMPTY_KEY;
    if (threadIdx.x == 0) { sh_n = 0u; sh_fail = 0u; }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < gd.end[nlev - 1]; e += THREADS) {
        int lv = 0;
        while (e >= gd.end[lv]) lv++;
        const uint32_t key = ents[e].key, a = key & 0xFFFFu;
        atomicOr(reinterpret_cast<uint32_t*>(role8) + (a >> 2), (1u << lv) << ((a & 3u) * 8u));
        uint32_t s = (key * 0x9E3779B1u) >> hash_shift;
        while (true) {
            const uint32_t old = atomicCAS(&hkey[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY) { hval[s] = (uint16_t)ents[e].z; hlev[s] = (unsigned char)lv; break; }
            s = (s + 1) & hash_mask;
        }
    }
    // (the first barrier of the tile loop publishes the tables)
    // A slot enters a queue at most once (qmask): two entries for one slot could be worked on by two threads at once, and the
    // second could then pair the old token with the token BEHIND the partner the first has just consumed.
    auto push = [&](int lv, int s) {
        const uint32_t bit = (1u << lv) << ((s & 3) * 8);
        if (atomicOr(reinterpret_cast<uint32_t*>(qmask) + (s >> 2), bit) & bit) return;
        const uint32_t at = atomicAdd(&qn[lv], 1u);
        if (at < (uint32_t)GRP_QCAP) queue[lv * GRP_QCAP + (int)at] = (uint16_t)s;
        else atomicOr(&qover, 1u << lv);
    };
    auto next_live_g = [&](int s) -> int {
        for (int j = s + 1; j < GRP_OFF + TILE + GRP_H; j++) if (ext[j] != H) return j;
        return -1;
    };
    uint32_t merged = 0;  // (counted at write-back)
    // the slot at stage index s, if it holds a first component of level lv whose next live token completes a pair of that level
    auto process = [&](int s, int lv) {
        const uint32_t t = ext[s];
        if (t == H || !((role8[t] >> lv) & 1u)) return;  // consumed, or the token changed since the slot was queued
        const int j = next_live_g(s);
        if (j < 0) return;
        const uint32_t key = pair_key(t, (uint32_t)ext[j]);
        uint32_t h = (key * 0x9E3779B1u) >> hash_shift;
        while (true) {
            const uint32_t k = hkey[h];
            if (k == key) break;
            if (k == EMPTY_KEY) return;
            h = (h + 1) & hash_mask;
        }
        if ((int)hlev[h] != lv) return;
        const uint32_t z = hval[h];
        ext[s] = (uint16_t)z;
        ext[j] = H;
        // the new token may start pairs of later levels of the group
        uint32_t later = (uint32_t)role8[z] & (0xFFu << (lv + 1)) & 0xFFu;
        while (later) {
            const int l2 = __ffs((int)later) - 1;
            later &= later - 1u;
            push(l2, s);
        }
    };
    for (uint32_t it = blockIdx.x; it < ntiles; it += gridDim.x) {
        const uint32_t tile = backwards ? ntiles - 1u - it : it;
        const uint4* gv = reinterpret_cast<const uint4*>(tok + (size_t)tile * TILE);
        uint4* gvo = reinterpret_cast<uint4*>(tok_out + (size_t)tile * TILE);
        uint4 v[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = gv[k * THREADS + (int)threadIdx.x];
        uint4* xv = reinterpret_cast<uint4*>(ext + GRP_OFF);
        __syncthreads();  // the previous tile has been written out of the stage (first round: the tables are complete)
#pragma unroll
        for (int k = 0; k < NV; k++) xv[k * THREADS + (int)threadIdx.x] = v[k];
        for (int i = (int)threadIdx.x; i < GRP_OFF; i += THREADS) {  // margins: halo tokens next to the tile, holes beyond them
            const DeepHalo* hp = halo + tile;
            ext[GRP_OFF - 1 - i] = i < GRP_H ? hp->l[i] : H;
            ext[GRP_OFF + TILE + i] = i < GRP_H ? hp->r[i] : H;
            if (i == 0) {  // nothing is unknown beyond a halo that ends with the sequence
                taint_l = hp->lend ? -(1 << 20) : GRP_OFF - 1 - GRP_H;
                taint_r = hp->rend ? (1 << 20) : GRP_OFF + TILE + GRP_H;
            }
        }
        if (threadIdx.x < GRP_MAXLV) qn[threadIdx.x] = 0u;
        if (threadIdx.x == GRP_MAXLV) qover = 0u;
        static_assert(GRP_EXT % 16 == 0, "the queue masks are cleared in 16-byte stores");
        for (int i = (int)threadIdx.x; i < GRP_EXT / 16; i += THREADS) reinterpret_cast<uint4*>(qmask)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        // ---- one look at every slot: queue it for the levels its token can start a pair in ----
        auto enqueue = [&](int s, uint32_t t) {
            if (t == H) return;
            uint32_t rb = role8[t];
            while (rb) {
                const int lv = __ffs((int)rb) - 1;
                rb &= rb - 1u;
                push(lv, s);
            }
        };
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t tv[VEC];
            unpack_vec<uint16_t>(v[k], tv);
            const int s0 = GRP_OFF + (k * THREADS + (int)threadIdx.x) * VEC;
#pragma unroll
            for (int i = 0; i < VEC; i++) enqueue(s0 + i, tv[i]);
        }
        for (int i = (int)threadIdx.x; i < 2 * GRP_H; i += THREADS) {
            const int s = i < GRP_H ? GRP_OFF - 1 - i : GRP_OFF + TILE + (i - GRP_H);
            enqueue(s, ext[s]);
        }
        __syncthreads();
        // ---- the levels, in order ----
        for (int lv = 0; lv < nlev; lv++) {
            if ((qover >> lv) & 1u) {
                // dense level on this tile: everybody rescans the own slots (+ the halo slots) instead of the queue
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const int s0 = GRP_OFF + (k * THREADS + (int)threadIdx.x) * VEC;
                    for (int i = 0; i < VEC; i++) process(s0 + i, lv);
                }
                for (int i = (int)threadIdx.x; i < 2 * GRP_H; i += THREADS)
                    process(i < GRP_H ? GRP_OFF - 1 - i : GRP_OFF + TILE + (i - GRP_H), lv);
            } else {
                const uint32_t nq = qn[lv] < (uint32_t)GRP_QCAP ? qn[lv] : (uint32_t)GRP_QCAP;
                for (uint32_t i = threadIdx.x; i < nq; i += THREADS) process((int)queue[lv * GRP_QCAP + (int)i], lv);
            }
            __syncthreads();  // level lv is complete (tokens, and the queues of the later levels) before the next one starts
            // the taint moves on to the next live token on either side
            if (threadIdx.x == 0 && taint_l >= 0) {
                int j = taint_l + 1;
                while (j < GRP_OFF + TILE && ext[j] == H) j++;
                taint_l = j;
                if (j >= GRP_OFF) sh_fail = 1u;
            }
            if (threadIdx.x == THREADS - 1 && taint_r < (1 << 20)) {
                int j = taint_r - 1;
                while (j >= GRP_OFF && ext[j] == H) j--;
                taint_r = j;
                if (j < GRP_OFF + TILE) sh_fail = 1u;
            }
            __syncthreads();  // (the scans read the stage as this level left it)
        }
        // ---- write the tile; a merge is counted where its second component was (every merge empties one slot) ----
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const uint4 o = xv[k * THREADS + (int)threadIdx.x];
            gvo[k * THREADS + (int)threadIdx.x] = o;  // out of place: every vector is written
            if (o.x != v[k].x || o.y != v[k].y || o.z != v[k].z || o.w != v[k].w)
                merged += (uint32_t)__popc(vec_mask<uint16_t>(o, 0xFFFFu) & ~vec_mask<uint16_t>(v[k], 0xFFFFu));
        }
    }
    if (merged) atomicAdd(&sh_n, merged);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh_n) atomicAdd(&res->merged, sh_n);
        if (sh_fail) atomicOr(&res->fail, 1u);
    }
}

}  // namespace bpe
