// tests/emul/vec_check.cpp — TEST INFRASTRUCTURE: the packed-halfword primitives of the merge pass (vec_has, vec_mask,
// vec_candidates, vec_pair_candidates in zig-bpe_b200/csrc/bpe_kernels.cuh) against their slot-by-slot definitions,
// on random vectors biased towards the values where the bit tricks could go wrong (A, B, holes, A ^ 0x8000, A +- 1, 0).
// Compiled with the CPU emulation header by tests/test_emul.py; prints "<cases> <bad>".
#define BPE_EMUL 1
#include "bpe_kernels.cuh"
#include <random>
using namespace bpe;
int main(int argc, char** argv) {
    std::mt19937 rng(argc > 1 ? (unsigned)atoi(argv[1]) : 1u);
    long bad = 0, n = 0;
    for (int it = 0; it < 2000000; it++) {
        uint32_t a = rng() % 5 == 0 ? rng() & 0xFFFFu : 10 + rng() % 3, b = rng() % 5 == 0 ? rng() & 0xFFFFu : 10 + rng() % 3;
        if (a == 0xFFFFu) a = 1;
        if (b == 0xFFFFu) b = 2;
        uint16_t t[9];
        for (int i = 0; i < 9; i++) {
            const int r = rng() % 8;
            t[i] = r < 2 ? a : r < 4 ? b : r < 6 ? 0xFFFF : (r == 6 ? (uint16_t)(a ^ 0x8000u) : (uint16_t)(rng() & 0xFFFFu));
            if (rng() % 50 == 0) t[i] = (uint16_t)(a - 1);
            if (rng() % 50 == 0) t[i] = (uint16_t)(a + 1);
            if (rng() % 60 == 0) t[i] = 0;
        }
        uint4 v;
        v.x = t[0] | (uint32_t)t[1] << 16; v.y = t[2] | (uint32_t)t[3] << 16; v.z = t[4] | (uint32_t)t[5] << 16; v.w = t[6] | (uint32_t)t[7] << 16;
        uint32_t want_pair = 0, want_a = 0;
        bool has = false;
        for (int i = 0; i < 8; i++) {
            const uint32_t bit = 1u << ((i >> 1) + 16 * (i & 1));  // vec_mask's layout: bit (i/2 + 16 * (i & 1)) <-> slot i
            if (t[i] == a) { has = true; want_a |= bit; }
            if (t[i] == a && (t[i + 1] == b || t[i + 1] == 0xFFFF)) want_pair |= bit;
        }
        bool ok = vec_pair_candidates<uint16_t>(v, a, b, t[8]) == want_pair && vec_has<uint16_t>(v, a) == has &&
                  vec_mask<uint16_t>(v, a) == want_a && vec_candidates<uint16_t>(v, a, b, t[8], true) == want_pair;
        for (int bitpos = 0; bitpos < 32 && ok; bitpos++)
            if ((want_a >> bitpos) & 1u) ok = t[mask_bit_to_slot<uint16_t>(bitpos)] == a;
        // the 32-bit instantiation (encode with id 65535 in the list)
        uint4 w; w.x = t[0]; w.y = t[1] == 0xFFFF ? 0xFFFFFFFFu : t[1]; w.z = t[2] == 0xFFFF ? 0xFFFFFFFFu : t[2]; w.w = t[3] == 0xFFFF ? 0xFFFFFFFFu : t[3];
        if (t[0] == 0xFFFF) w.x = 0xFFFFFFFFu;
        const uint32_t nf32 = t[4] == 0xFFFF ? 0xFFFFFFFFu : t[4];
        uint32_t want32 = 0;
        for (int i = 0; i < 4; i++) if (t[i] == a && (t[i + 1] == b || t[i + 1] == 0xFFFF)) want32 |= 1u << i;
        ok = ok && vec_pair_candidates<uint32_t>(w, a, b, nf32) == want32;
        if (!ok) bad++;
        n++;
    }
    printf("%ld %ld\n", n, bad);
    return bad ? 1 : 0;
}
