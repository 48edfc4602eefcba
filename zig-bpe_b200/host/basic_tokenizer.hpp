// basic_tokenizer.hpp — C++ host mirror of the reference's BasicTokenizer
// (/root/reference/src/basic_tokenizer.zig:52-349) over the C ABI of include/bpe_b200.h.
// The struct, merge list, merges.txt (de)serialisation (:319-348) and the stderr text stay on the host;
// train / encode / decode run in CUDA. There is no CPU fallback: construction throws without a GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "bpe_b200.h"

namespace zigbpe {

struct TrainError : std::runtime_error { using std::runtime_error::runtime_error; };  // TrainError (:6-10)
struct InvalidVocabSize : TrainError { InvalidVocabSize() : TrainError("InvalidVocabSize") {} };
struct OutOfMemory : TrainError { OutOfMemory(const std::string& m) : TrainError("OutOfMemory: " + m) {} };
struct InvalidToken : std::runtime_error { InvalidToken() : std::runtime_error("InvalidToken") {} };
struct FormatError : std::runtime_error { using std::runtime_error::runtime_error; };  // InvalidFormat / InvalidCharacter / Overflow / StreamTooLong

struct CharPair { uint16_t first, second; };
struct Merge { CharPair pair; uint16_t new_token; };

struct TimeStats {  // utils/time_statistics.zig:4-34
    int64_t sort_pairs_time = 0, replace_pair_time = 0, generate_pairs_time = 0, just_count_pairs_time = 0;
    size_t sort_pairs_calls = 0, replace_pair_calls = 0, generate_pairs_calls = 0, just_count_pairs_calls = 0;
};

inline void printTimeStats(const TimeStats& s, int64_t total) {  // :36-60
    auto line = [](const char* name, int64_t t, size_t calls) {
        std::fprintf(stderr, "%s: %.3fs total, %zu calls, %.3fs avg\n", name, t / 1000.0, calls, calls ? t / (calls * 1000.0) : NAN);
    };
    std::fprintf(stderr, "\nTime statistics:\n");
    line("sortCodePointPairs", s.sort_pairs_time, s.sort_pairs_calls);
    line("replaceTopPairWithIndex", s.replace_pair_time, s.replace_pair_calls);
    line("generateCodePointPairs", s.generate_pairs_time, s.generate_pairs_calls);
    line("countPointPairs", s.just_count_pairs_time, s.just_count_pairs_calls);
    int64_t other = total - s.sort_pairs_time - s.replace_pair_time - s.generate_pairs_time - s.just_count_pairs_time;
    std::fprintf(stderr, "Other operations: %.3fs\n", other / 1000.0);
}

class BasicTokenizer {
public:
    std::vector<Merge> merges;  // Merges.merges (:17-38)
    TimeStats timeStats;

    explicit BasicTokenizer(int device = 0) {  // init (:57-64)
        if (bpe_ctx_create(&ctx_, device) != BPE_OK) throw OutOfMemory(bpe_last_error(nullptr));
        bpe_ctx_set_option(ctx_, "time_phases", 1);
    }
    ~BasicTokenizer() { bpe_ctx_destroy(ctx_); }  // deinit (:66-69)
    BasicTokenizer(const BasicTokenizer&) = delete;
    BasicTokenizer& operator=(const BasicTokenizer&) = delete;

    void put(CharPair pair, uint16_t new_token) { merges.push_back({pair, new_token}); }  // Merges.put (:32-37)

    static std::vector<uint16_t> generateInitialTokens(const std::string& text) {  // :155-170
        return std::vector<uint16_t>(reinterpret_cast<const uint8_t*>(text.data()), reinterpret_cast<const uint8_t*>(text.data()) + text.size());
    }

    void train(const std::string& text, uint16_t vocabSize, bool verbose) {  // :140-153
        if (vocabSize < 256) { printTimeStats(timeStats, 0); throw InvalidVocabSize(); }
        const size_t cap = (size_t)vocabSize - 256;
        std::vector<bpe_merge_t> out(cap ? cap : 1);
        std::vector<uint64_t> counts(cap ? cap : 1);
        size_t n = 0;
        bpe_stats_t st{};
        int rc = bpe_train(ctx_, reinterpret_cast<const uint8_t*>(text.data()), text.size(), vocabSize, out.data(), counts.data(), &n, &st);
        timeStats.sort_pairs_time += (int64_t)st.sort_pairs_ms; timeStats.sort_pairs_calls += st.sort_pairs_calls;
        timeStats.replace_pair_time += (int64_t)st.replace_pair_ms; timeStats.replace_pair_calls += st.replace_pair_calls;
        timeStats.just_count_pairs_time += (int64_t)st.just_count_pairs_ms; timeStats.just_count_pairs_calls += st.just_count_pairs_calls;
        // generateInitialTokens' own line (:156-160); the widening is the load kernel here
        std::fprintf(stderr, "generateInitialTokens runtime: %.3f seconds\n", st.kernel_ms[0] / 1000.0);
        if (rc == BPE_ERR_INVALID_VOCAB) { printTimeStats(timeStats, (int64_t)st.total_ms); throw InvalidVocabSize(); }
        if (rc != BPE_OK) { printTimeStats(timeStats, (int64_t)st.total_ms); throw OutOfMemory(bpe_last_error(ctx_)); }
        for (size_t i = 0; i < n; i++) {
            if (verbose)  // printMergeInfo (:308-317)
                std::fprintf(stderr, "merge %zu/%u: (%u,%u) -> %u had %llu occurrences\n", i + 1, (unsigned)(vocabSize - 256),
                             (unsigned)out[i].first, (unsigned)out[i].second, (unsigned)out[i].new_token, (unsigned long long)counts[i]);
            put({out[i].first, out[i].second}, out[i].new_token);
        }
        if (n < cap) std::fprintf(stderr, "No more pairs to merge. Stopping early.\n");  // :189
        printTimeStats(timeStats, (int64_t)st.total_ms);  // the deferred print of :142-145
    }

    std::vector<uint16_t> encode(const std::string& text) {  // :71-88
        std::vector<uint16_t> out(text.size() ? text.size() : 1);
        std::vector<bpe_merge_t> ms = abi();
        size_t n = 0;
        int rc = bpe_encode(ctx_, reinterpret_cast<const uint8_t*>(text.data()), text.size(), ms.data(), ms.size(), out.data(), &n, nullptr);
        if (rc != BPE_OK) throw OutOfMemory(bpe_last_error(ctx_));
        std::fprintf(stderr, "generateInitialTokens runtime: 0.000 seconds\n");  // encode calls it too (:72); fused away here
        out.resize(n);
        return out;
    }

    std::string decode(const std::vector<uint16_t>& tokens) {  // :90-138
        std::vector<bpe_merge_t> ms = abi();
        size_t need = 0, n = 0;
        int rc = bpe_decode_size(ctx_, tokens.data(), tokens.size(), ms.data(), ms.size(), &need);
        if (rc == BPE_ERR_INVALID_TOKEN) throw InvalidToken();
        if (rc != BPE_OK) throw OutOfMemory(bpe_last_error(ctx_));
        std::string out(need, '\0');
        rc = bpe_decode(ctx_, tokens.data(), tokens.size(), ms.data(), ms.size(), reinterpret_cast<uint8_t*>(&out[0]), need, &n, nullptr);
        if (rc == BPE_ERR_INVALID_TOKEN) throw InvalidToken();
        if (rc != BPE_OK) throw OutOfMemory(bpe_last_error(ctx_));
        std::fprintf(stderr, "generateInitialTokens runtime: 0.000 seconds\n");  // encode calls it too (:72); fused away here
        out.resize(n);
        return out;
    }

    void serializeMerges(const std::string& path) const {  // :319-330
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot create " + path);
        for (const Merge& m : merges) std::fprintf(f, "%u,%u,%u\n", (unsigned)m.pair.first, (unsigned)m.pair.second, (unsigned)m.new_token);
        std::fclose(f);
    }

    void deserializeMerges(const std::string& path) { readMergesFile(path, merges); }  // :332-348

    // the reader itself (no GPU involved): 100-byte line buffer, parseInt(u16), appends; merges parsed before a bad
    // line stay appended, as in the reference
    static void readMergesFile(const std::string& path, std::vector<Merge>& into) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("FileNotFound: " + path);
        while (true) {
            char line[101];
            size_t len = 0;
            bool delim = false, eof = false;
            for (int k = 0; k < 100; k++) {
                int c = std::fgetc(f);
                if (c == EOF) { eof = true; break; }
                if (c == '\n') { delim = true; break; }
                line[len++] = (char)c;
            }
            if (!delim && !eof) { std::fclose(f); throw FormatError("StreamTooLong"); }
            if (eof && len == 0) break;
            uint16_t v[3];
            size_t pos = 0;
            for (int k = 0; k < 3; k++) {
                if (pos > len) { std::fclose(f); throw FormatError("InvalidFormat"); }
                size_t start = pos;
                while (pos < len && line[pos] != ',') pos++;
                try { v[k] = parse_u16(line + start, pos - start); } catch (...) { std::fclose(f); throw; }
                pos++;  // skip ',' (or step past the end: no further field)
            }
            into.push_back({{v[0], v[1]}, v[2]});
            if (eof) break;
        }
        std::fclose(f);
    }

private:
    static uint16_t parse_u16(const char* s, size_t len) {  // std.fmt.parseInt(u16, s, 10)
        if (len == 0) throw FormatError("InvalidCharacter");
        bool neg = false;
        if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; s++; len--; }
        if (len == 0 || s[0] == '_' || s[len - 1] == '_') throw FormatError("InvalidCharacter");
        uint32_t acc = 0;
        for (size_t i = 0; i < len; i++) {
            if (s[i] == '_') continue;
            if (s[i] < '0' || s[i] > '9') throw FormatError("InvalidCharacter");
            uint32_t d = (uint32_t)(s[i] - '0');
            if (neg) { if (acc || d) throw FormatError("Overflow"); }
            else { acc = acc * 10 + d; if (acc > 65535) throw FormatError("Overflow"); }
        }
        return (uint16_t)acc;
    }
    std::vector<bpe_merge_t> abi() const {
        std::vector<bpe_merge_t> ms(merges.size());
        for (size_t i = 0; i < merges.size(); i++) ms[i] = {merges[i].pair.first, merges[i].pair.second, merges[i].new_token};
        return ms;
    }
    bpe_ctx* ctx_ = nullptr;
};

}  // namespace zigbpe
