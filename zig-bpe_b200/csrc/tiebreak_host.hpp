// tiebreak_host.hpp — exact replay of the reference's pair table for one merge step.
//
// Host-side control logic of the train loop (not a CPU fallback for the data path): on the rare
// tie steps that the GPU occupancy test cannot settle (zig_resolve_kernel -> TIE_NEED_REPLAY),
// the distinct live pairs are replayed, in order of first occurrence, through a model of Zig
// 0.13's std.AutoHashMap exactly as countCodePointPairs (src/basic_tokenizer.zig:257-278) would
// fill it: power-of-two capacity (min 8), linear probing from hash & (cap-1), growth *before*
// the lookup once 80 % full, rehash in old-slot order. The winner is
// the tied pair in the lowest slot (iteration order + stable sort, :291-303, then [0] at :193).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "bpe_common.cuh"

namespace bpe {

struct ReplayKey {
    uint32_t key;
    uint32_t home_hash;  // low 32 bits of the Wyhash (enough for any capacity <= 2^32)
    uint64_t first_pos;  // global position of the first occurrence
};

class ZigTableReplay {
public:
    // keys: all live pairs with their first-occurrence position; last_pair_pos: position of the
    // last adjacent pair of the sequence. Returns the slot of every key in `slots_out` order of
    // `keys` after sorting by first_pos (keys is sorted in place).
    void run(std::vector<ReplayKey>& keys, uint64_t last_pair_pos) {
        std::sort(keys.begin(), keys.end(),
                  [](const ReplayKey& a, const ReplayKey& b) { return a.first_pos < b.first_pos; });
        cap_ = 0; size_ = 0; available_ = 0;
        slots_.clear();
        for (size_t i = 0; i < keys.size(); i++) {
            if (available_ == 0) grow(keys);
            insert((uint32_t)i, keys[i].home_hash);
            size_++;
            available_--;
        }
        // one more getOrPut follows the last new key unless its first occurrence is the last pair
        if (!keys.empty() && available_ == 0 && keys.back().first_pos != last_pair_pos) grow(keys);
    }
    uint32_t capacity() const { return cap_; }
    // slot index of key i (index into the sorted keys vector)
    uint32_t slot_of(const std::vector<ReplayKey>& keys, uint32_t key) const {
        for (uint32_t s = 0; s < cap_; s++)
            if (slots_[s] != NONE && keys[slots_[s]].key == key) return s;
        return NONE;
    }
    // lowest-slot key among `tied`
    uint32_t winner(const std::vector<ReplayKey>& keys, const std::vector<uint32_t>& tied) const {
        std::vector<uint32_t> t(tied);
        std::sort(t.begin(), t.end());
        for (uint32_t s = 0; s < cap_; s++) {
            if (slots_[s] == NONE) continue;
            uint32_t k = keys[slots_[s]].key;
            if (std::binary_search(t.begin(), t.end(), k)) return k;
        }
        return NONE;
    }
    static constexpr uint32_t NONE = 0xFFFFFFFFu;

private:
    void insert(uint32_t idx, uint32_t hash) {
        uint32_t s = hash & (cap_ - 1);
        while (slots_[s] != NONE) s = (s + 1) & (cap_ - 1);
        slots_[s] = idx;
    }
    void grow(const std::vector<ReplayKey>& keys) {
        uint64_t want = ((uint64_t)size_ + 1) * 100 / 80 + 1;
        uint32_t nc = 8;
        while (nc < want) nc <<= 1;
        std::vector<uint32_t> old;
        old.swap(slots_);
        uint32_t old_cap = cap_;
        slots_.assign(nc, NONE);
        cap_ = nc;
        for (uint32_t s = 0; s < old_cap; s++)
            if (old[s] != NONE) insert(old[s], keys[old[s]].home_hash);
        available_ = zig_max_load(nc) - size_;
    }
    std::vector<uint32_t> slots_;
    uint32_t cap_ = 0, size_ = 0, available_ = 0;
};

}  // namespace bpe
