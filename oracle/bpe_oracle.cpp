// oracle/bpe_oracle.cpp — TEST INFRASTRUCTURE ONLY. See bpe_oracle.hpp for the scope
// statement and the parity-pinning note. Every function cites the reference lines it
// follows (paths relative to /root/reference).
#include "bpe_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <unordered_map>

namespace oracle {

// =====================================================================================
// std.hash.Wyhash, Zig 0.13.0 (lib/std/hash/wyhash.zig upstream)
// =====================================================================================
static const uint64_t WY_SECRET[4] = {0xa0761d6478bd642fULL, 0xe7037ed1a0b428dbULL,
                                      0x8ebc6af09c88c6e3ULL, 0x589965cc75374cc3ULL};

static inline void wy_mum(uint64_t* a, uint64_t* b) {
    __uint128_t x = (__uint128_t)(*a) * (__uint128_t)(*b);
    *a = (uint64_t)x;
    *b = (uint64_t)(x >> 64);
}
static inline uint64_t wy_mix(uint64_t a, uint64_t b) {
    wy_mum(&a, &b);
    return a ^ b;
}
static inline uint64_t wy_r8(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t wy_r4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

uint64_t wyhash(uint64_t seed, const uint8_t* in, size_t len) {
    uint64_t st[3];
    st[0] = seed ^ wy_mix(seed ^ WY_SECRET[0], WY_SECRET[1]);
    st[1] = st[0];
    st[2] = st[0];
    uint64_t a, b;
    if (len <= 16) {  // smallKey
        if (len >= 4) {
            size_t end = len - 4;
            size_t quarter = (len >> 3) << 2;
            a = (wy_r4(in) << 32) | wy_r4(in + quarter);
            b = (wy_r4(in + end) << 32) | wy_r4(in + end - quarter);
        } else if (len > 0) {
            a = ((uint64_t)in[0] << 16) | ((uint64_t)in[len >> 1] << 8) | (uint64_t)in[len - 1];
            b = 0;
        } else {
            a = 0;
            b = 0;
        }
    } else {
        size_t i = 0;
        if (len >= 48) {
            while (i + 48 < len) {  // round()
                for (int k = 0; k < 3; k++) {
                    uint64_t ra = wy_r8(in + i + 8 * (2 * k));
                    uint64_t rb = wy_r8(in + i + 8 * (2 * k + 1));
                    st[k] = wy_mix(ra ^ WY_SECRET[k + 1], rb ^ st[k]);
                }
                i += 48;
            }
            st[0] ^= st[1] ^ st[2];  // final0
        }
        // final1
        const uint8_t* rest = in + i;
        size_t rlen = len - i;
        size_t j = 0;
        while (j + 16 < rlen) {
            st[0] = wy_mix(wy_r8(rest + j) ^ WY_SECRET[1], wy_r8(rest + j + 8) ^ st[0]);
            j += 16;
        }
        a = wy_r8(in + len - 16);
        b = wy_r8(in + len - 8);
    }
    // final2
    a ^= WY_SECRET[1];
    b ^= st[0];
    wy_mum(&a, &b);
    return wy_mix(a ^ WY_SECRET[0] ^ (uint64_t)len, b ^ WY_SECRET[1]);
}

// getAutoHashFn(CharPair): the struct has a unique representation, so the 4 key bytes are
// hashed as-is (first LE, second LE) with seed 0 (Appendix A.2).
uint64_t hash_pair(CharPair k) {
    uint8_t bytes[4] = {(uint8_t)(k.first & 0xff), (uint8_t)(k.first >> 8),
                        (uint8_t)(k.second & 0xff), (uint8_t)(k.second >> 8)};
    return wyhash(0, bytes, 4);
}

// =====================================================================================
// std.HashMapUnmanaged model (Appendix A.3 / A.4): pow2 capacity (min 8), linear probing,
// max load 80 %, growIfNeeded(1) BEFORE the lookup, rehash in old-slot order.
// =====================================================================================
void ZigPairMap::grow(uint32_t new_cap) {
    if (new_cap < 8) new_cap = 8;  // minimal_capacity
    std::vector<uint8_t> nused(new_cap, 0);
    std::vector<CharPair> nkeys(new_cap);
    std::vector<size_t> nvals(new_cap);
    const uint32_t mask = new_cap - 1;
    uint32_t moved = 0;
    for (uint32_t i = 0; i < cap_ && moved < size_; i++) {  // old-slot order
        if (!used_[i]) continue;
        uint32_t idx = (uint32_t)(hash_pair(keys_[i]) & mask);
        while (nused[idx]) idx = (idx + 1) & mask;
        nused[idx] = 1;
        nkeys[idx] = keys_[i];
        nvals[idx] = vals_[i];
        moved++;
    }
    used_.swap(nused);
    keys_.swap(nkeys);
    vals_.swap(nvals);
    cap_ = new_cap;
    available_ = (uint32_t)(((uint64_t)new_cap * 80) / 100) - size_;
}

size_t* ZigPairMap::get_or_put(CharPair key, bool* found_existing) {
    // growIfNeeded(1): if (1 > available) grow(capacityForSize(load + 1))
    if (available_ < 1) {
        uint64_t load = size_;
        uint64_t want = (load + 1) * 100 / 80 + 1;
        uint32_t nc = 1;
        while (nc < want) nc <<= 1;  // ceilPowerOfTwo
        grow(nc);
    }
    const uint32_t mask = cap_ - 1;
    uint32_t idx = (uint32_t)(hash_pair(key) & mask);
    while (used_[idx]) {
        if (keys_[idx].first == key.first && keys_[idx].second == key.second) {
            *found_existing = true;
            return &vals_[idx];
        }
        idx = (idx + 1) & mask;
    }
    used_[idx] = 1;
    keys_[idx] = key;
    vals_[idx] = 0;
    size_++;
    available_--;
    *found_existing = false;
    return &vals_[idx];
}

// =====================================================================================
// BasicTokenizer
// =====================================================================================
using Clock = std::chrono::steady_clock;
static inline double ms_since(Clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

// generateCodePointPairs (:234-255): materialise the N-1 adjacent pairs.
static void generate_pairs(const std::vector<uint16_t>& tokens, std::vector<CharPair>& pairs) {
    pairs.clear();
    if (tokens.empty()) return;  // reference underflows here (:246); defined as "no pairs"
    pairs.reserve(tokens.size() - 1);
    for (size_t i = 0; i + 1 < tokens.size(); i++) pairs.push_back({tokens[i], tokens[i + 1]});
}

// countCodePointPairs (:257-278): AutoHashMap getOrPut per pair (counts overlapping pairs).
static void count_pairs(const std::vector<CharPair>& pairs, ZigPairMap& map) {
    for (const CharPair& p : pairs) {
        bool found;
        size_t* v = map.get_or_put(p, &found);
        if (!found) *v = 1; else *v += 1;
    }
}

// sortCodePointPairs (:280-306): copy in iteration (slot) order, stable sort desc by count.
static void sort_pairs(const ZigPairMap& map, std::vector<PairCount>& sorted) {
    sorted.clear();
    sorted.reserve(map.count());
    map.for_each([&](uint32_t, CharPair k, size_t v) { sorted.push_back({k, v}); });
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const PairCount& a, const PairCount& b) { return a.count > b.count; });
}

// replaceTopPairWithNewToken (:207-232): in-place two-pointer, i += 2 on match.
static void replace_pair(std::vector<uint16_t>& t, CharPair pair, uint16_t new_token) {
    if (t.empty()) return;
    size_t i = 0, j = 0;
    while (i < t.size() - 1) {
        if (t[i] == pair.first && t[i + 1] == pair.second) {
            t[j] = new_token;
            i += 2;
        } else {
            t[j] = t[i];
            i += 1;
        }
        j += 1;
    }
    if (i < t.size()) {
        t[j] = t[i];
        j += 1;
    }
    t.resize(j);
}

static void print_merge_info(unsigned idx, unsigned vocab, CharPair p, size_t count) {  // :308-317
    fprintf(stderr, "merge %u/%u: (%u,%u) -> %u had %zu occurrences\n", idx - 256 + 1, vocab - 256,
            (unsigned)p.first, (unsigned)p.second, idx, count);
}

static inline uint32_t pkey(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }

int BasicTokenizer::train(const uint8_t* text, size_t n, unsigned vocab_size, bool verbose,
                          long max_steps, bool fast) {
    if (vocab_size < 256) return ERR_INVALID_VOCAB;  // :147-149
    if (vocab_size > 65535) return ERR_INVALID_VOCAB;  // vocabSize is u16 (:140)
    // generateInitialTokens (:155-170) + the copy in expandVocabulary (:173-175)
    std::vector<uint16_t> cur(n);
    for (size_t i = 0; i < n; i++) cur[i] = text[i];

    long steps = 0;
    if (!fast) {
        std::vector<CharPair> pairs;
        std::vector<PairCount> sorted;
        for (unsigned idx = 256; idx < vocab_size; idx++) {  // :183
            if (max_steps >= 0 && steps >= max_steps) break;
            auto t0 = Clock::now();
            generate_pairs(cur, pairs);
            stats.generate_pairs += ms_since(t0); stats.generate_calls++;
            t0 = Clock::now();
            ZigPairMap map;
            count_pairs(pairs, map);
            stats.just_count_pairs += ms_since(t0); stats.count_calls++;
            t0 = Clock::now();
            sort_pairs(map, sorted);
            stats.sort_pairs += ms_since(t0); stats.sort_calls++;
            if (sorted.empty()) {  // :188-191
                if (verbose) fprintf(stderr, "No more pairs to merge. Stopping early.\n");
                break;
            }
            PairCount top = sorted[0];  // :193
            if (verbose) print_merge_info(idx, vocab_size, top.pair, top.count);
            merges.push_back({top.pair, (uint16_t)idx});  // :199
            merge_counts.push_back(top.count);
            t0 = Clock::now();
            replace_pair(cur, top.pair, (uint16_t)idx);  // :201
            stats.replace_pair += ms_since(t0); stats.replace_calls++;
            steps++;
        }
        return OK;
    }

    // ---- fast mode: identical results, incremental counts; Zig-map replay on tie steps ----
    std::unordered_map<uint32_t, uint64_t> counts;
    counts.reserve(1 << 16);
    for (size_t i = 0; i + 1 < cur.size(); i++) counts[pkey(cur[i], cur[i + 1])]++;
    for (unsigned idx = 256; idx < vocab_size; idx++) {
        if (max_steps >= 0 && steps >= max_steps) break;
        uint64_t best = 0;
        size_t ntied = 0;
        uint32_t best_key = 0;
        for (auto& kv : counts) {
            if (kv.second > best) { best = kv.second; ntied = 1; best_key = kv.first; }
            else if (kv.second == best && best > 0) ntied++;
        }
        if (best == 0) {
            if (verbose) fprintf(stderr, "No more pairs to merge. Stopping early.\n");
            break;
        }
        CharPair win{(uint16_t)(best_key & 0xffff), (uint16_t)(best_key >> 16)};
        if (ntied > 1) {
            // replay the reference's table for this step: insert pairs in sequence order
            ZigPairMap map;
            for (size_t i = 0; i + 1 < cur.size(); i++) {
                bool found;
                size_t* v = map.get_or_put({cur[i], cur[i + 1]}, &found);
                if (!found) *v = 1; else *v += 1;
            }
            bool have = false;
            map.for_each([&](uint32_t, CharPair k, size_t v) {
                if (!have && v == best) { win = k; have = true; }
            });
        }
        if (verbose) print_merge_info(idx, vocab_size, win, best);
        merges.push_back({win, (uint16_t)idx});
        merge_counts.push_back(best);
        // replace (:207-232) with incremental count deltas
        const uint16_t A = win.first, B = win.second, X = (uint16_t)idx;
        auto dec = [&](uint16_t a, uint16_t b) {
            auto it = counts.find(pkey(a, b));
            if (--(it->second) == 0) counts.erase(it);
        };
        auto inc = [&](uint16_t a, uint16_t b) { counts[pkey(a, b)]++; };
        std::vector<uint16_t>& t = cur;
        const size_t len = t.size();
        size_t i = 0, j = 0;
        bool prev_merged = false;   // token written at j-1 came from a merge
        uint16_t prev_pre = 0;      // pre-merge token at i-1
        while (len > 0 && i < len - 1) {
            if (t[i] == A && t[i + 1] == B) {
                dec(A, B);
                if (i > 0) {
                    if (prev_merged) { dec(prev_pre, A); inc(X, X); }
                    else { dec(prev_pre, A); inc(prev_pre, X); }
                }
                if (i + 2 < len) {
                    uint16_t nx = t[i + 2];
                    bool next_is_start = (nx == A && i + 3 < len && t[i + 3] == B);
                    if (!next_is_start) { dec(B, nx); inc(X, nx); }
                }
                prev_pre = B;
                prev_merged = true;
                t[j] = X;
                i += 2;
            } else {
                prev_pre = t[i];
                prev_merged = false;
                t[j] = t[i];
                i += 1;
            }
            j += 1;
        }
        if (i < len) { t[j] = t[i]; j += 1; }
        t.resize(j);
        steps++;
    }
    return OK;
}

int BasicTokenizer::encode(const uint8_t* text, size_t n, std::vector<uint16_t>& out, bool linear) const {
    out.resize(n);
    for (size_t i = 0; i < n; i++) out[i] = text[i];  // generateInitialTokens (:72)
    if (!linear) {
        // verbatim (:75-85): replace then orderedRemove(i+1); i is not advanced after a hit
        for (const Merge& m : merges) {
            size_t i = 0;
            while (i < out.size()) {
                if (i + 1 < out.size() && out[i] == m.pair.first && out[i + 1] == m.pair.second) {
                    out[i] = m.new_token;
                    out.erase(out.begin() + (long)(i + 1));
                } else {
                    i += 1;
                }
            }
        }
        return OK;
    }
    // linear restatement: one two-pointer pass per merge. "i is not advanced" is kept:
    // after a hit the freshly written token is compared again with what follows.
    for (const Merge& m : merges) {
        const uint16_t A = m.pair.first, B = m.pair.second, X = m.new_token;
        size_t len = out.size();
        if (len == 0) break;
        size_t j = 0;          // write cursor; out[j] is the token under the reference's `i`
        size_t r = 1;          // next unread token
        // invariant: out[0..j] are final for this pass except out[j] which is "current"
        while (true) {
            if (r < len && out[j] == A && out[r] == B) {
                out[j] = X;    // tokens[i] = new_token; orderedRemove(i+1)
                r++;
            } else {
                if (r >= len) break;
                j++;
                out[j] = out[r];
                r++;
            }
        }
        out.resize(j + 1);
    }
    return OK;
}

int BasicTokenizer::decode(const uint16_t* toks, size_t n, std::vector<uint8_t>& out) const {
    out.clear();
    // findMerge (:109-116): first merge whose new_token matches; decodeMerge (:118-138) recursive.
    struct Rec {
        const std::vector<Merge>& ms;
        std::vector<uint8_t>& o;
        int find(uint16_t tok) const {
            for (size_t i = 0; i < ms.size(); i++) if (ms[i].new_token == tok) return (int)i;
            return -1;
        }
        int expand(const Merge& m, int depth) {
            if (depth > 100000) return ERR_INVALID_TOKEN;  // reference would overflow its stack
            uint16_t parts[2] = {m.pair.first, m.pair.second};
            for (uint16_t p : parts) {
                if (p < 256) o.push_back((uint8_t)p);
                else {
                    int k = find(p);
                    if (k < 0) return ERR_INVALID_TOKEN;
                    int rc = expand(ms[(size_t)k], depth + 1);
                    if (rc) return rc;
                }
            }
            return OK;
        }
    } rec{merges, out};
    for (size_t i = 0; i < n; i++) {
        uint16_t t = toks[i];
        if (t < 256) out.push_back((uint8_t)t);
        else {
            int k = rec.find(t);
            if (k < 0) return ERR_INVALID_TOKEN;
            int rc = rec.expand(merges[(size_t)k], 0);
            if (rc) return rc;
        }
    }
    return OK;
}

std::string BasicTokenizer::serialize_to_string() const {  // :319-330 "{d},{d},{d}\n"
    std::string s;
    char buf[64];
    for (const Merge& m : merges) {
        snprintf(buf, sizeof buf, "%u,%u,%u\n", (unsigned)m.pair.first, (unsigned)m.pair.second,
                 (unsigned)m.new_token);
        s += buf;
    }
    return s;
}

int BasicTokenizer::serialize(const std::string& path) const {
    FILE* f = fopen(path.c_str(), "wb");  // createFile truncates (:320)
    if (!f) return -10;
    std::string s = serialize_to_string();
    size_t w = fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    return w == s.size() ? OK : -11;
}

// std.fmt.parseInt(u16, s, 10): optional sign, '_' separators (not first/last), digits.
// returns 0 ok, -20 InvalidCharacter, -21 Overflow
static int zig_parse_u16(const char* s, size_t len, uint16_t* out) {
    if (len == 0) return -20;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = (s[0] == '-'); s++; len--; }
    if (len == 0) return -20;
    if (s[0] == '_' || s[len - 1] == '_') return -20;
    uint32_t acc = 0;
    for (size_t i = 0; i < len; i++) {
        char c = s[i];
        if (c == '_') continue;
        if (c < '0' || c > '9') return -20;
        uint32_t d = (uint32_t)(c - '0');
        if (neg) {
            // accumulates downward in an unsigned type: only zero survives
            if (acc != 0 || d != 0) return -21;
        } else {
            acc = acc * 10 + d;
            if (acc > 65535) return -21;
        }
    }
    *out = (uint16_t)acc;
    return 0;
}

int BasicTokenizer::deserialize(const std::string& path) {  // :332-348
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return -10;  // FileNotFound etc.
    int rc = OK;
    while (true) {
        // readUntilDelimiterOrEof into a 100-byte buffer: delimiter must be among the first
        // 100 bytes read, else error.StreamTooLong
        char line[101];
        size_t len = 0;
        bool got_delim = false, eof = false;
        for (int k = 0; k < 100; k++) {
            int c = fgetc(f);
            if (c == EOF) { eof = true; break; }
            if (c == '\n') { got_delim = true; break; }
            line[len++] = (char)c;
        }
        if (!got_delim && !eof) { rc = -22; break; }  // StreamTooLong
        if (eof && len == 0) break;                   // null -> loop ends
        // std.mem.split(u8, line, ","): three fields required, extras ignored
        const char* fld[3];
        size_t flen[3];
        size_t pos = 0;
        int nf = 0;
        while (nf < 3) {
            size_t start = pos;
            while (pos < len && line[pos] != ',') pos++;
            fld[nf] = line + start;
            flen[nf] = pos - start;
            nf++;
            if (pos >= len) break;
            pos++;  // skip ','
            if (pos == len && nf < 3) {  // trailing comma yields one more empty field
                fld[nf] = line + pos; flen[nf] = 0; nf++;
                break;
            }
        }
        uint16_t v[3];
        bool bad = false;
        for (int k = 0; k < 3; k++) {
            if (k >= nf) { rc = ERR_FORMAT; bad = true; break; }  // error.InvalidFormat
            int e = zig_parse_u16(fld[k], flen[k], &v[k]);
            if (e) { rc = e; bad = true; break; }
        }
        if (bad) break;
        merges.push_back({{v[0], v[1]}, v[2]});  // appends, never clears (:346)
        if (eof) break;
    }
    fclose(f);
    return rc;
}

}  // namespace oracle

// =====================================================================================
// C API
// =====================================================================================
using namespace oracle;

extern "C" {

uint64_t oracle_wyhash(uint64_t seed, const uint8_t* p, size_t len) { return wyhash(seed, p, len); }
uint64_t oracle_hash_pair(uint16_t first, uint16_t second) { return hash_pair({first, second}); }

int oracle_train(const uint8_t* text, size_t n, unsigned vocab_size, long max_steps, int fast,
                 uint16_t* merges_out, uint64_t* counts_out, size_t* n_merges_out, double* stats8_out) {
    BasicTokenizer tk;
    int rc = tk.train(text, n, vocab_size, false, max_steps, fast != 0);
    if (rc) return rc;
    for (size_t i = 0; i < tk.merges.size(); i++) {
        merges_out[3 * i + 0] = tk.merges[i].pair.first;
        merges_out[3 * i + 1] = tk.merges[i].pair.second;
        merges_out[3 * i + 2] = tk.merges[i].new_token;
        if (counts_out) counts_out[i] = tk.merge_counts[i];
    }
    *n_merges_out = tk.merges.size();
    if (stats8_out) {
        stats8_out[0] = tk.stats.sort_pairs; stats8_out[1] = (double)tk.stats.sort_calls;
        stats8_out[2] = tk.stats.replace_pair; stats8_out[3] = (double)tk.stats.replace_calls;
        stats8_out[4] = tk.stats.generate_pairs; stats8_out[5] = (double)tk.stats.generate_calls;
        stats8_out[6] = tk.stats.just_count_pairs; stats8_out[7] = (double)tk.stats.count_calls;
    }
    return OK;
}

static void load_merges(BasicTokenizer& tk, const uint16_t* merges, size_t n) {
    tk.merges.resize(n);
    for (size_t i = 0; i < n; i++) tk.merges[i] = {{merges[3 * i], merges[3 * i + 1]}, merges[3 * i + 2]};
}

int oracle_encode(const uint8_t* text, size_t n, const uint16_t* merges, size_t n_merges, int linear,
                  uint16_t* out, size_t* n_out) {
    BasicTokenizer tk;
    load_merges(tk, merges, n_merges);
    std::vector<uint16_t> v;
    int rc = tk.encode(text, n, v, linear != 0);
    if (rc) return rc;
    memcpy(out, v.data(), v.size() * sizeof(uint16_t));
    *n_out = v.size();
    return OK;
}

int oracle_decode(const uint16_t* toks, size_t n, const uint16_t* merges, size_t n_merges,
                  uint8_t* out, size_t cap, size_t* n_out) {
    BasicTokenizer tk;
    load_merges(tk, merges, n_merges);
    std::vector<uint8_t> v;
    int rc = tk.decode(toks, n, v);
    if (rc) return rc;
    *n_out = v.size();
    if (v.size() > cap) return ERR_OOM;
    memcpy(out, v.data(), v.size());
    return OK;
}

int oracle_serialize(const uint16_t* merges, size_t n_merges, const char* path) {
    BasicTokenizer tk;
    load_merges(tk, merges, n_merges);
    return tk.serialize(path);
}

long oracle_deserialize(const char* path, uint16_t* merges_out, size_t cap) {
    BasicTokenizer tk;
    int rc = tk.deserialize(path);
    size_t k = std::min(cap, tk.merges.size());
    for (size_t i = 0; i < k; i++) {
        merges_out[3 * i + 0] = tk.merges[i].pair.first;
        merges_out[3 * i + 1] = tk.merges[i].pair.second;
        merges_out[3 * i + 2] = tk.merges[i].new_token;
    }
    if (rc != OK) return rc < 0 ? (long)rc : -(long)(100 + rc);
    return (long)tk.merges.size();
}

uint32_t oracle_zigmap_slots(const uint16_t* pairs, size_t n_pairs, uint32_t* slots_out) {
    ZigPairMap map;
    for (size_t i = 0; i < n_pairs; i++) {
        bool f;
        map.get_or_put({pairs[2 * i], pairs[2 * i + 1]}, &f);
    }
    if (slots_out) {
        std::unordered_map<uint32_t, uint32_t> slot_of;
        map.for_each([&](uint32_t s, CharPair k, size_t) { slot_of[pkey(k.first, k.second)] = s; });
        for (size_t i = 0; i < n_pairs; i++) slots_out[i] = slot_of[pkey(pairs[2 * i], pairs[2 * i + 1])];
    }
    return map.capacity();
}

}  // extern "C"

#ifdef ORACLE_MAIN
// CLI used for quick checks and CPU-baseline timing:
//   bpe_oracle train <text-file> <vocab> <merges-out> [max_steps] [fast] [verbose]
// verbose=1 prints the reference's per-merge line (:309) to stderr as each merge is learned, so a long
// run (tools/make_golden_big.py) leaves a usable prefix behind if it is stopped early.
#include <fstream>
#include <iterator>
int main(int argc, char** argv) {
    if (argc < 5 || std::string(argv[1]) != "train") {
        fprintf(stderr, "usage: %s train <text> <vocab> <merges-out> [max_steps] [fast]\n", argv[0]);
        return 2;
    }
    std::ifstream f(argv[2], std::ios::binary);
    std::vector<uint8_t> text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    long max_steps = argc > 5 ? atol(argv[5]) : -1;
    bool fast = argc > 6 && atoi(argv[6]) != 0;
    bool verbose = argc > 7 && atoi(argv[7]) != 0;
    BasicTokenizer tk;
    auto t0 = Clock::now();
    int rc = tk.train(text.data(), text.size(), (unsigned)atoi(argv[3]), verbose, max_steps, fast);
    double ms = ms_since(t0);
    if (rc) { fprintf(stderr, "train failed: %d\n", rc); return 1; }
    tk.serialize(argv[4]);
    fprintf(stderr, "%zu merges in %.1f ms (%.2f merges/s)\n", tk.merges.size(), ms,
            tk.merges.size() / (ms / 1000.0));
    return 0;
}
#endif
