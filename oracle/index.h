// ORACLE — test infrastructure only. See bitmap.h.
// In-memory stand-in for the LMDB environment the reference reads through heed
// (crates/milli/src/index.rs:97-124): sorted byte keys + CBO values, looked up by binary
// search (the B-tree descent) and scanned by prefix (heed `prefix_iter`).
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "bitmap.h"

namespace orc {

enum DbId {
    DB_WORD_DOCIDS = 0,
    DB_EXACT_WORD_DOCIDS = 1,
    DB_WORD_PREFIX_DOCIDS = 2,
    DB_EXACT_WORD_PREFIX_DOCIDS = 3,
    DB_WORD_PAIR_PROXIMITY_DOCIDS = 4,
    DB_WORD_POSITION_DOCIDS = 5,
    DB_WORD_FID_DOCIDS = 6,
    DB_WORD_PREFIX_POSITION_DOCIDS = 7,
    DB_WORD_PREFIX_FID_DOCIDS = 8,
    DB_FIELD_ID_WORD_COUNT_DOCIDS = 9,
    DB_COUNT = 10
};

struct Db {
    uint64_t n = 0;
    std::vector<uint8_t> keys, vals;
    std::vector<uint64_t> koff, voff;

    int cmp_key(uint64_t i, const uint8_t *k, size_t kn) const {
        size_t n_i = koff[i + 1] - koff[i];
        int c = memcmp(keys.data() + koff[i], k, std::min(n_i, kn));
        if (c) return c;
        return n_i < kn ? -1 : (n_i > kn ? 1 : 0);
    }
    uint64_t lower_bound(const uint8_t *k, size_t kn) const {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (cmp_key(mid, k, kn) < 0)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    }
    Span get(const std::string &k) const {
        uint64_t i = lower_bound((const uint8_t *)k.data(), k.size());
        Span s;
        if (i < n && cmp_key(i, (const uint8_t *)k.data(), k.size()) == 0) {
            s.p = vals.data() + voff[i];
            s.n = voff[i + 1] - voff[i];
            s.some = true;
        }
        return s;
    }
    bool has_prefix(uint64_t i, const std::string &p) const {
        size_t n_i = koff[i + 1] - koff[i];
        return n_i >= p.size() && memcmp(keys.data() + koff[i], p.data(), p.size()) == 0;
    }
    std::string key(uint64_t i) const { return std::string((const char *)keys.data() + koff[i], koff[i + 1] - koff[i]); }
    Span val(uint64_t i) const {
        Span s;
        s.p = vals.data() + voff[i];
        s.n = voff[i + 1] - voff[i];
        s.some = true;
        return s;
    }
};

enum Criterion { C_WORDS = 0, C_TYPO, C_PROXIMITY, C_ATTRIBUTE, C_ATTRIBUTE_RANK, C_WORD_POSITION, C_SORT, C_EXACTNESS };
enum TermsMatchingStrategy { TMS_LAST = 0, TMS_ALL = 1, TMS_FREQUENCY = 2 };
enum ScoringStrategy { SCORING_SKIP = 0, SCORING_DETAILED = 1 };

struct Settings {
    uint32_t n_fields = 1;
    std::vector<uint16_t> weights;  // fid -> weight (searchable order)
    std::vector<int> criteria{C_WORDS, C_TYPO, C_PROXIMITY, C_ATTRIBUTE_RANK, C_SORT, C_WORD_POSITION, C_EXACTNESS};
    bool authorize_typos = true;
    uint32_t min_word_len_one_typo = 5;   // index.rs:46
    uint32_t min_word_len_two_typos = 9;  // index.rs:47
    std::set<std::string> exact_words;
    std::map<std::vector<std::string>, std::vector<std::vector<std::string>>> synonyms;
    bool prefix_search = true;
};

struct Index {
    // sorted dictionary (the words FST, index.rs:1238): word id = rank
    std::vector<uint8_t> dict_bytes;
    std::vector<uint64_t> dict_off;
    Db dbs[DB_COUNT];
    Bitmap documents_ids;
    Settings settings;
    // vector store: one embedding per row, cosine
    uint32_t dim = 0;
    std::vector<float> embeddings;
    std::vector<uint16_t> embeddings_f16;  // alternative storage (IEEE binary16 rows) for large stores; `embeddings` is then empty
    std::vector<float> emb_norms;
    std::vector<uint32_t> emb_docids;
    // Whole-universe nearest neighbours of the current batch's query vectors, computed by one blocked pass over the store
    // (batch_nns, rules.h) instead of one full scan per query: query vector pointer -> ascending (docid, distance).
    // Read-only while queries run.
    mutable std::map<const float *, std::vector<std::pair<uint32_t, float>>> nns_cache;
    float emb_at(size_t r, uint32_t i) const {
        if (!embeddings_f16.empty()) {
            _Float16 h;
            memcpy(&h, &embeddings_f16[r * dim + i], 2);
            return (float)h;
        }
        return embeddings[r * dim + i];
    }
    bool has_distribution = false;
    float dist_mean = 0, dist_sigma = 0;

    uint64_t n_words() const { return dict_off.empty() ? 0 : dict_off.size() - 1; }
    const uint8_t *word_ptr(uint64_t i) const { return dict_bytes.data() + dict_off[i]; }
    size_t word_len(uint64_t i) const { return dict_off[i + 1] - dict_off[i]; }
    std::string word(uint64_t i) const { return std::string((const char *)word_ptr(i), word_len(i)); }
    int cmp_word(uint64_t i, const uint8_t *k, size_t kn) const {
        size_t n_i = word_len(i);
        int c = memcmp(word_ptr(i), k, std::min(n_i, kn));
        if (c) return c;
        return n_i < kn ? -1 : (n_i > kn ? 1 : 0);
    }
    uint64_t dict_lower_bound(const uint8_t *k, size_t kn, uint64_t lo, uint64_t hi) const {
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (cmp_word(mid, k, kn) < 0)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    }
    bool contains_word(const std::string &w) const {
        uint64_t i = dict_lower_bound((const uint8_t *)w.data(), w.size(), 0, n_words());
        return i < n_words() && cmp_word(i, (const uint8_t *)w.data(), w.size()) == 0;
    }
    uint16_t max_searchable_weight() const {
        uint16_t m = 0;
        for (auto w : settings.weights) m = std::max(m, w);
        return m;
    }
};

}  // namespace orc
