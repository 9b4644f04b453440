// ORACLE — test infrastructure only (CPU restatement of milli's query-time path).
// Nothing under oracle/ is linked, imported or executed by the product library.
//
// Bitmap: stands in for `roaring::RoaringBitmap` (roaring 0.10.12, Cargo.lock:5994).
// Only set semantics matter for parity, so it is a two-mode set of u32:
// sorted array (small) or dense words (large).  CBO decoding follows
// crates/milli/src/heed_codec/roaring_bitmap/cbo_roaring_bitmap_codec.rs:53-85 and the
// portable roaring layout restated in roaring_bitmap_len_codec.rs:9-54.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace orc {

struct Bitmap {
    // invariant: exactly one representation is active
    bool dense = false;
    std::vector<uint32_t> arr;    // sorted unique, when !dense
    std::vector<uint64_t> words;  // when dense
    uint64_t card = 0;            // valid when dense

    static constexpr size_t DENSE_MIN = 4096;  // switch to dense above this many elements

    bool is_empty() const { return dense ? card == 0 : arr.empty(); }
    uint64_t len() const { return dense ? card : arr.size(); }
    void clear() {
        dense = false;
        arr.clear();
        words.clear();
        card = 0;
    }
    bool contains(uint32_t x) const {
        if (dense) {
            size_t w = x >> 6;
            return w < words.size() && ((words[w] >> (x & 63)) & 1);
        }
        return std::binary_search(arr.begin(), arr.end(), x);
    }
    static Bitmap from_sorted(std::vector<uint32_t> v) {
        Bitmap b;
        b.arr = std::move(v);
        b.normalize();
        return b;
    }
    void recount() {
        card = 0;
        for (auto w : words) card += (uint64_t)__builtin_popcountll(w);
    }
    void to_dense() {
        if (dense) return;
        size_t nw = arr.empty() ? 0 : (arr.back() >> 6) + 1;
        words.assign(nw, 0);
        for (auto x : arr) words[x >> 6] |= 1ull << (x & 63);
        card = arr.size();
        arr.clear();
        arr.shrink_to_fit();
        dense = true;
    }
    void to_array() {
        if (!dense) return;
        arr.clear();
        arr.reserve(card);
        for (size_t w = 0; w < words.size(); w++) {
            uint64_t bits = words[w];
            while (bits) {
                arr.push_back((uint32_t)(w * 64 + __builtin_ctzll(bits)));
                bits &= bits - 1;
            }
        }
        words.clear();
        words.shrink_to_fit();
        dense = false;
        card = 0;
    }
    void normalize() {
        if (dense) {
            if (card < DENSE_MIN / 2) to_array();
        } else if (arr.size() > DENSE_MIN)
            to_dense();
    }
    template <class F>
    void for_each(F f) const {
        if (dense) {
            for (size_t w = 0; w < words.size(); w++) {
                uint64_t bits = words[w];
                while (bits) {
                    f((uint32_t)(w * 64 + __builtin_ctzll(bits)));
                    bits &= bits - 1;
                }
            }
        } else
            for (auto x : arr) f(x);
    }
    std::vector<uint32_t> to_vec() const {
        std::vector<uint32_t> v;
        v.reserve(len());
        for_each([&](uint32_t x) { v.push_back(x); });
        return v;
    }
    void insert(uint32_t x) {
        if (dense) {
            size_t w = x >> 6;
            if (w >= words.size()) words.resize(w + 1, 0);
            if (!((words[w] >> (x & 63)) & 1)) {
                words[w] |= 1ull << (x & 63);
                card++;
            }
        } else {
            auto it = std::lower_bound(arr.begin(), arr.end(), x);
            if (it == arr.end() || *it != x) arr.insert(it, x);
            if (arr.size() > DENSE_MIN) to_dense();
        }
    }

    // self |= o
    void or_with(const Bitmap &o) {
        if (o.is_empty()) return;
        if (!dense && !o.dense) {
            std::vector<uint32_t> r;
            r.reserve(arr.size() + o.arr.size());
            std::set_union(arr.begin(), arr.end(), o.arr.begin(), o.arr.end(), std::back_inserter(r));
            arr.swap(r);
            if (arr.size() > DENSE_MIN) to_dense();
            return;
        }
        to_dense();
        if (o.dense) {
            if (o.words.size() > words.size()) words.resize(o.words.size(), 0);
            for (size_t i = 0; i < o.words.size(); i++) words[i] |= o.words[i];
            recount();
        } else {
            for (auto x : o.arr) insert(x);
        }
    }
    // self &= o
    void and_with(const Bitmap &o) {
        if (is_empty()) return;
        if (o.is_empty()) {
            clear();
            return;
        }
        if (!dense) {
            size_t k = 0;
            for (auto x : arr)
                if (o.contains(x)) arr[k++] = x;
            arr.resize(k);
            return;
        }
        if (o.dense) {
            size_t n = std::min(words.size(), o.words.size());
            words.resize(n);
            for (size_t i = 0; i < n; i++) words[i] &= o.words[i];
            recount();
            normalize();
        } else {
            std::vector<uint32_t> r;
            for (auto x : o.arr)
                if (contains(x)) r.push_back(x);
            clear();
            arr.swap(r);
        }
    }
    // self -= o
    void sub(const Bitmap &o) {
        if (is_empty() || o.is_empty()) return;
        if (!dense) {
            size_t k = 0;
            for (auto x : arr)
                if (!o.contains(x)) arr[k++] = x;
            arr.resize(k);
            return;
        }
        if (o.dense) {
            size_t n = std::min(words.size(), o.words.size());
            for (size_t i = 0; i < n; i++) words[i] &= ~o.words[i];
            recount();
        } else {
            for (auto x : o.arr) {
                size_t w = x >> 6;
                if (w < words.size() && ((words[w] >> (x & 63)) & 1)) {
                    words[w] &= ~(1ull << (x & 63));
                    card--;
                }
            }
        }
        normalize();
    }
    bool is_disjoint(const Bitmap &o) const {
        if (is_empty() || o.is_empty()) return true;
        if (!dense) {
            for (auto x : arr)
                if (o.contains(x)) return false;
            return true;
        }
        if (!o.dense) return o.is_disjoint(*this);
        size_t n = std::min(words.size(), o.words.size());
        for (size_t i = 0; i < n; i++)
            if (words[i] & o.words[i]) return false;
        return true;
    }
    bool is_superset(const Bitmap &o) const {
        bool ok = true;
        o.for_each([&](uint32_t x) {
            if (!contains(x)) ok = false;
        });
        return ok;
    }
    bool equals(const Bitmap &o) const { return len() == o.len() && is_superset(o); }
};

inline Bitmap bm_and(const Bitmap &a, const Bitmap &b) {
    Bitmap r = a.len() <= b.len() ? a : b;
    r.and_with(a.len() <= b.len() ? b : a);
    return r;
}
inline Bitmap bm_or(const Bitmap &a, const Bitmap &b) {
    Bitmap r = a;
    r.or_with(b);
    return r;
}

struct Span {
    const uint8_t *p = nullptr;
    size_t n = 0;
    bool some = false;
};

// Iterate the docids of a CBO value, in ascending order.
template <class F>
inline void cbo_for_each(const uint8_t *p, size_t n, F f) {
    if (n <= 7 * 4) {  // THRESHOLD * size_of::<u32>()  (cbo_roaring_bitmap_codec.rs:54)
        for (size_t i = 0; i + 4 <= n; i += 4) {
            uint32_t v;
            memcpy(&v, p + i, 4);
            f(v);
        }
        return;
    }
    uint32_t cookie, nc;
    memcpy(&cookie, p, 4);
    memcpy(&nc, p + 4, 4);
    if (cookie != 12346) throw std::runtime_error("cbo: run containers / bad cookie unsupported");
    const uint8_t *desc = p + 8;
    const uint8_t *data = p + 8 + 8 * (size_t)nc;  // descriptive header + offset header
    for (uint32_t c = 0; c < nc; c++) {
        uint16_t key, cm1;
        memcpy(&key, desc + 4 * c, 2);
        memcpy(&cm1, desc + 4 * c + 2, 2);
        uint32_t card = (uint32_t)cm1 + 1;
        uint32_t hi = (uint32_t)key << 16;
        if (card <= 4096) {
            for (uint32_t i = 0; i < card; i++) {
                uint16_t lo;
                memcpy(&lo, data + 2 * i, 2);
                f(hi | lo);
            }
            data += 2 * (size_t)card;
        } else {
            for (uint32_t w = 0; w < 1024; w++) {
                uint64_t bits;
                memcpy(&bits, data + 8 * w, 8);
                while (bits) {
                    f(hi | (w * 64 + (uint32_t)__builtin_ctzll(bits)));
                    bits &= bits - 1;
                }
            }
            data += 8192;
        }
    }
}

// CboRoaringBitmapCodec::deserialize_from
inline Bitmap cbo_decode(const uint8_t *p, size_t n) {
    std::vector<uint32_t> v;
    cbo_for_each(p, n, [&](uint32_t x) { v.push_back(x); });
    std::sort(v.begin(), v.end());  // raw-u32 form is written ascending, but `insert` semantics tolerate any order
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return Bitmap::from_sorted(std::move(v));
}
// CboRoaringBitmapCodec::intersection_with_serialized (:76-85)
inline Bitmap cbo_intersect(const uint8_t *p, size_t n, const Bitmap &universe) {
    std::vector<uint32_t> v;
    cbo_for_each(p, n, [&](uint32_t x) {
        if (universe.contains(x)) v.push_back(x);
    });
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return Bitmap::from_sorted(std::move(v));
}
// CboRoaringBitmapLenCodec (roaring_bitmap_len_codec.rs)
inline uint64_t cbo_len(const uint8_t *p, size_t n) {
    if (n <= 28) return n / 4;
    uint32_t nc;
    memcpy(&nc, p + 4, 4);
    uint64_t total = 0;
    for (uint32_t c = 0; c < nc; c++) {
        uint16_t cm1;
        memcpy(&cm1, p + 8 + 4 * c + 2, 2);
        total += (uint64_t)cm1 + 1;
    }
    return total;
}

}  // namespace orc
