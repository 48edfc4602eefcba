/* tools/synthcorpus.c — "synthcorpus-v1": the deterministic synthetic corpora of SURVEY.md §8(d).
 *
 * The same generator feeds the oracle (CPU) and the CUDA path, so both see identical bytes.
 * The corpus is a pure function of (seed, variant, byte offset): it is defined as a sequence
 * of independent 1 MiB blocks, block b generated from its own splitmix64 stream, so any byte
 * range can be produced by any number of threads / ranks with identical results.
 *
 *   lexicon : W = 50,000 words, length 1..12 (geometric, p = 0.2, truncated), letters from a
 *             26-letter alphabet with 1/rank weights
 *   sampling: Zipf(s = 1.0) over lexicon ranks by inverse CDF
 *   seps    : " " 85 %, ", " 5 %, ". " 5 %, ".\n" 3 %, "\n\n" 2 %
 *   variant 0 (ascii): letters only
 *   variant 1 (utf8) : 10 % of lexicon entries are 1..4 code points from U+00C0-U+024F (2-byte)
 *                      or U+4E00-U+9FFF / U+AC00-U+D7A3 (3-byte); output is valid UTF-8 except
 *                      where a block boundary or the final truncation cuts a sequence
 *   variant 2 (byte) : variant 1 plus 1 % of lexicon entries that are raw random byte strings
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SC_W 50000
#define SC_BLOCK ((size_t)1 << 20)
#define SC_MAXWORD 40

typedef struct {
    uint8_t bytes[SC_W][SC_MAXWORD];
    uint8_t len[SC_W];
    double cdf[SC_W];
    uint32_t guide[4097]; /* guide[k] = first rank with cdf >= k/4096 */
} sc_lexicon;

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t* s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char SC_ALPHA[27] = "etaoinshrdlcumwfgypbvkjxqz";

static int put_utf8(uint8_t* dst, uint32_t cp) {
    if (cp < 0x800) {
        dst[0] = (uint8_t)(0xC0 | (cp >> 6));
        dst[1] = (uint8_t)(0x80 | (cp & 0x3F));
        return 2;
    }
    dst[0] = (uint8_t)(0xE0 | (cp >> 12));
    dst[1] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
    dst[2] = (uint8_t)(0x80 | (cp & 0x3F));
    return 3;
}

static sc_lexicon* build_lexicon(uint64_t seed, int variant) {
    sc_lexicon* lx = (sc_lexicon*)malloc(sizeof(sc_lexicon));
    if (!lx) return NULL;
    uint64_t s = seed ^ 0x6C657869636F6E00ULL; /* "lexicon" */
    double lw[26], lsum = 0;
    for (int i = 0; i < 26; i++) { lw[i] = 1.0 / (double)(i + 1); lsum += lw[i]; }
    double lcdf[26], acc = 0;
    for (int i = 0; i < 26; i++) { acc += lw[i] / lsum; lcdf[i] = acc; }
    lcdf[25] = 1.0;
    for (int w = 0; w < SC_W; w++) {
        int len = 1;
        while (len < 12 && u01(&s) >= 0.2) len++;
        double kind = u01(&s);
        int n = 0;
        if (variant >= 2 && kind < 0.01) {
            for (int i = 0; i < len; i++) lx->bytes[w][n++] = (uint8_t)(splitmix64(&s) & 0xFF);
        } else if (variant >= 1 && kind < 0.11) {
            int chars = 1 + (len - 1) / 3;
            uint64_t r = splitmix64(&s) % 3;
            for (int i = 0; i < chars; i++) {
                uint32_t cp;
                if (r == 0) cp = 0x00C0 + (uint32_t)(splitmix64(&s) % (0x024F - 0x00C0 + 1));
                else if (r == 1) cp = 0x4E00 + (uint32_t)(splitmix64(&s) % (0x9FFF - 0x4E00 + 1));
                else cp = 0xAC00 + (uint32_t)(splitmix64(&s) % (0xD7A3 - 0xAC00 + 1));
                n += put_utf8(&lx->bytes[w][n], cp);
            }
        } else {
            for (int i = 0; i < len; i++) {
                double u = u01(&s);
                int k = 0;
                while (k < 25 && u > lcdf[k]) k++;
                lx->bytes[w][n++] = (uint8_t)SC_ALPHA[k];
            }
        }
        lx->len[w] = (uint8_t)n;
    }
    double h = 0;
    for (int w = 0; w < SC_W; w++) h += 1.0 / (double)(w + 1);
    acc = 0;
    for (int w = 0; w < SC_W; w++) { acc += (1.0 / (double)(w + 1)) / h; lx->cdf[w] = acc; }
    lx->cdf[SC_W - 1] = 1.0;
    uint32_t r = 0;
    for (int k = 0; k <= 4096; k++) {
        double t = (double)k / 4096.0;
        while (r < SC_W - 1 && lx->cdf[r] < t) r++;
        lx->guide[k] = r;
    }
    return lx;
}

static inline uint32_t sample_rank(const sc_lexicon* lx, double u) {
    uint32_t k = (uint32_t)(u * 4096.0);
    uint32_t lo = lx->guide[k], hi = lx->guide[k + 1];
    while (lo < hi) { /* first rank with cdf >= u */
        uint32_t mid = (lo + hi) >> 1;
        if (lx->cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* fill one block (index b) completely into tmp[SC_BLOCK + slack] */
static void gen_block(const sc_lexicon* lx, uint64_t seed, uint64_t b, uint8_t* tmp) {
    uint64_t s = seed ^ ((b + 1) * 0xD1342543DE82EF95ULL);
    (void)splitmix64(&s);
    size_t n = 0;
    while (n < SC_BLOCK) {
        uint32_t r = sample_rank(lx, u01(&s));
        memcpy(tmp + n, lx->bytes[r], lx->len[r]);
        n += lx->len[r];
        double u = u01(&s);
        if (u < 0.85) { tmp[n++] = ' '; }
        else if (u < 0.90) { tmp[n++] = ','; tmp[n++] = ' '; }
        else if (u < 0.95) { tmp[n++] = '.'; tmp[n++] = ' '; }
        else if (u < 0.98) { tmp[n++] = '.'; tmp[n++] = '\n'; }
        else { tmp[n++] = '\n'; tmp[n++] = '\n'; }
    }
}

/* bytes [offset, offset+len) of corpus (seed, variant) into out. returns 0 on success */
int synth_corpus_range(uint8_t* out, uint64_t offset, uint64_t len, uint64_t seed, int variant, int nthreads) {
    if (len == 0) return 0;
    sc_lexicon* lx = build_lexicon(seed, variant);
    if (!lx) return 2;
    uint64_t b0 = offset / SC_BLOCK, b1 = (offset + len - 1) / SC_BLOCK;
    int fail = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        uint8_t* tmp = (uint8_t*)malloc(SC_BLOCK + 2 * SC_MAXWORD + 8);
        if (!tmp) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t b = (int64_t)b0; b <= (int64_t)b1; b++) {
                gen_block(lx, seed, (uint64_t)b, tmp);
                uint64_t bs = (uint64_t)b * SC_BLOCK, be = bs + SC_BLOCK;
                uint64_t lo = bs > offset ? bs : offset;
                uint64_t hi = be < offset + len ? be : offset + len;
                memcpy(out + (lo - offset), tmp + (lo - bs), (size_t)(hi - lo));
            }
            free(tmp);
        }
    }
    free(lx);
    return fail ? 2 : 0;
}

int synth_corpus(uint8_t* out, uint64_t n, uint64_t seed, int variant, int nthreads) {
    return synth_corpus_range(out, 0, n, seed, variant, nthreads);
}
