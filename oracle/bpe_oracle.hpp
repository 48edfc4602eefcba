// oracle/bpe_oracle.hpp — TEST INFRASTRUCTURE ONLY (not shipped, not measured as the product).
//
// CPU restatement of dbtreasure/zig-bpe's BasicTokenizer (reference:
// src/basic_tokenizer.zig) plus the slice of the Zig 0.13.0 standard library
// the result depends on (std.hash.Wyhash, std.HashMapUnmanaged, stable sort).
// Zig 0.13.0 is a third-party dependency that is absent from /root/reference
// (pinned only in prose, README.md:17-19); its behaviour is restated from
// SURVEY.md Appendix A.
//
// Parity pinning: the restatement reproduces the reference's committed golden
// merges.txt byte for byte (44 merges, one tie event) and the reference's five
// in-file unit tests (src/basic_tokenizer.zig:351-461); tie-breaks beyond that
// one event rest on the Appendix-A model of std.AutoHashMap ("parity pinned by
// golden merges.txt; tie order beyond step 38 pinned by the Zig std model only").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may use this code.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

// ---- std.hash.Wyhash (Zig 0.13) -------------------------------------------------
uint64_t wyhash(uint64_t seed, const uint8_t* p, size_t len);

// ---- reference types (src/basic_tokenizer.zig:12-48) ----------------------------
struct CharPair { uint16_t first, second; };
struct Merge { CharPair pair; uint16_t new_token; };
struct PairCount { CharPair pair; size_t count; };

// getAutoHashFn(CharPair): Wyhash(0) over the 4 key bytes, first LE then second LE.
uint64_t hash_pair(CharPair k);

// ---- std.AutoHashMap(CharPair, usize) model (Appendix A.3/A.4) -------------------
class ZigPairMap {
public:
    ZigPairMap() = default;
    // getOrPut: returns pointer to the value, sets found_existing.
    size_t* get_or_put(CharPair key, bool* found_existing);
    uint32_t capacity() const { return cap_; }
    uint32_t count() const { return size_; }
    // iteration in ascending slot order (iterator(), :291-297)
    template <class F> void for_each(F&& f) const {
        for (uint32_t i = 0; i < cap_; i++) if (used_[i]) f(i, keys_[i], vals_[i]);
    }
private:
    void grow(uint32_t new_cap);
    uint32_t cap_ = 0, size_ = 0, available_ = 0;
    std::vector<uint8_t> used_;
    std::vector<CharPair> keys_;
    std::vector<size_t> vals_;
};

enum Status { OK = 0, ERR_INVALID_VOCAB = 1, ERR_OOM = 2, ERR_INVALID_TOKEN = 3, ERR_FORMAT = 6 };

struct TimeStatsMs {  // utils/time_statistics.zig:4-34 (ms buckets + call counts)
    double sort_pairs = 0, replace_pair = 0, generate_pairs = 0, just_count_pairs = 0;
    size_t sort_calls = 0, replace_calls = 0, generate_calls = 0, count_calls = 0;
};

class BasicTokenizer {
public:
    std::vector<Merge> merges;
    std::vector<size_t> merge_counts;  // count of the winning pair per learned merge (verbose line :309)
    TimeStatsMs stats;

    // train (:140-153) with expandVocabulary (:172-205). max_steps < 0 = no limit
    // (a limit is used only to time the first K steps for the CPU baseline).
    // fast=false: verbatim per-step pipeline (materialised pairs, Zig map, stable sort).
    // fast=true : same results; counts kept incrementally, the Zig map is replayed
    //             only on steps whose top count is tied (checked against fast=false in tests).
    int train(const uint8_t* text, size_t n, unsigned vocab_size, bool verbose, long max_steps = -1, bool fast = false);
    // encode (:71-88). linear=false is the verbatim orderedRemove loop (quadratic);
    // linear=true is a two-pointer pass per merge producing identical output.
    int encode(const uint8_t* text, size_t n, std::vector<uint16_t>& out, bool linear = true) const;
    // decode (:90-138)
    int decode(const uint16_t* toks, size_t n, std::vector<uint8_t>& out) const;
    // serializeMerges / deserializeMerges (:319-348)
    int serialize(const std::string& path) const;
    int deserialize(const std::string& path);
    std::string serialize_to_string() const;
};

}  // namespace oracle

// ---- C API for ctypes (tests / bench cpu_baseline) ---------------------------------
extern "C" {
uint64_t oracle_wyhash(uint64_t seed, const uint8_t* p, size_t len);
uint64_t oracle_hash_pair(uint16_t first, uint16_t second);
// merges_out: triples (first, second, new_token) as uint16; counts_out nullable.
int oracle_train(const uint8_t* text, size_t n, unsigned vocab_size, long max_steps, int fast,
                 uint16_t* merges_out, uint64_t* counts_out, size_t* n_merges_out, double* stats8_out);
int oracle_encode(const uint8_t* text, size_t n, const uint16_t* merges, size_t n_merges, int linear,
                  uint16_t* out, size_t* n_out);
int oracle_decode(const uint16_t* toks, size_t n, const uint16_t* merges, size_t n_merges,
                  uint8_t* out, size_t cap, size_t* n_out);
int oracle_serialize(const uint16_t* merges, size_t n_merges, const char* path);
// returns number of merges appended (<= cap) or negative error
long oracle_deserialize(const char* path, uint16_t* merges_out, size_t cap);
// Zig map probe: insert keys in order, return slot of each key and final capacity.
uint32_t oracle_zigmap_slots(const uint16_t* pairs, size_t n_pairs, uint32_t* slots_out);
}
