// ORACLE — test infrastructure only. See bitmap.h.
// SearchContext + posting access (search/new/mod.rs:77-133, search/new/db_cache.rs:183-719),
// Levenshtein-automaton x dictionary walk (search/mod.rs:565-577, query_term/compute_derivations.rs:75-168),
// term construction (compute_derivations.rs:170-383, parse_query.rs:204-300).
//
// Third-party arithmetic restated from published semantics (crates are not vendored):
//   fst 0.4.7            -> dictionary enumerated in bytewise lexicographic order
//   levenshtein_automata 0.2.1, LevenshteinAutomatonBuilder::new(n, true)
//                        -> restricted Damerau-Levenshtein (transposition costs 1, adjacent
//                           transposed pair not edited again), distance over chars, prefix
//                           variant = min over prefixes of the candidate.
// ASCII only: char == byte.  Non-ASCII words are outside the pinned domain of this oracle.
#pragma once
#include <unordered_map>

#include "terms.h"

namespace orc {

struct Ctx {
    const Index &index;
    std::vector<std::string> words;
    std::unordered_map<std::string, uint32_t> word_ids;
    std::vector<Phrase> phrases;
    std::map<std::vector<int32_t>, uint32_t> phrase_ids;
    std::vector<QueryTerm> terms;
    std::map<uint32_t, Bitmap> phrase_docids;
    // accounting for the roofline's "algorithmic bytes" 
    uint64_t stat_fetches = 0, stat_fetch_bytes = 0;

    explicit Ctx(const Index &ix) : index(ix) {}

    uint32_t intern_word(const std::string &w) {
        auto it = word_ids.find(w);
        if (it != word_ids.end()) return it->second;
        uint32_t id = (uint32_t)words.size();
        words.push_back(w);
        word_ids.emplace(w, id);
        return id;
    }
    uint32_t intern_phrase(const Phrase &p) {
        auto it = phrase_ids.find(p.words);
        if (it != phrase_ids.end()) return it->second;
        uint32_t id = (uint32_t)phrases.size();
        phrases.push_back(p);
        phrase_ids.emplace(p.words, id);
        return id;
    }

    // DatabaseCache::get_value (db_cache.rs:50-84)
    bool get_value(int db, const std::string &key, const Bitmap *universe, Bitmap &out) {
        Span s = index.dbs[db].get(key);
        if (!s.some) return false;
        stat_fetches++;
        stat_fetch_bytes += s.n;
        out = universe ? cbo_intersect(s.p, s.n, *universe) : cbo_decode(s.p, s.n);
        return true;
    }
    static std::string key_u16(const std::string &w, uint16_t v) {
        std::string k = w;
        k.push_back(0);
        k.push_back((char)(v >> 8));
        k.push_back((char)(v & 0xff));
        return k;
    }
    static std::string key_pair(uint8_t prox, const std::string &w1, const std::string &w2) {
        std::string k;
        k.push_back((char)prox);
        k += w1;
        k.push_back(0);
        k += w2;
        return k;
    }
    // db_cache.rs:183-205
    bool word_docids(const Bitmap *universe, Word w, Bitmap &out) {
        const std::string &s = words[w.id];
        if (w.kind == W_ORIGINAL) {
            Bitmap e, t;
            bool he = get_value(DB_EXACT_WORD_DOCIDS, s, universe, e);
            bool ht = get_value(DB_WORD_DOCIDS, s, universe, t);
            if (!he && !ht) return false;
            if (he && ht) {
                e.or_with(t);
                out = std::move(e);
            } else
                out = he ? std::move(e) : std::move(t);
            return true;
        }
        return get_value(DB_WORD_DOCIDS, s, universe, out);
    }
    // db_cache.rs:272-294
    bool word_prefix_docids(const Bitmap *universe, Word w, Bitmap &out) {
        const std::string &s = words[w.id];
        if (w.kind == W_ORIGINAL) {
            Bitmap e, t;
            bool he = get_value(DB_EXACT_WORD_PREFIX_DOCIDS, s, universe, e);
            bool ht = get_value(DB_WORD_PREFIX_DOCIDS, s, universe, t);
            if (!he && !ht) return false;
            if (he && ht) {
                e.or_with(t);
                out = std::move(e);
            } else
                out = he ? std::move(e) : std::move(t);
            return true;
        }
        return get_value(DB_WORD_PREFIX_DOCIDS, s, universe, out);
    }
    // db_cache.rs:361-424 (ProximityPrecision::ByWord only)
    bool word_pair_proximity_docids(const Bitmap *universe, uint32_t w1, uint32_t w2, uint8_t prox, Bitmap &out) {
        return get_value(DB_WORD_PAIR_PROXIMITY_DOCIDS, key_pair(prox, words[w1], words[w2]), universe, out);
    }
    bool word_pair_proximity_docids_len(uint32_t w1, uint32_t w2, uint8_t prox, uint64_t &len) {
        Span s = index.dbs[DB_WORD_PAIR_PROXIMITY_DOCIDS].get(key_pair(prox, words[w1], words[w2]));
        if (!s.some) return false;
        len = cbo_len(s.p, s.n);
        return true;
    }
    // db_cache.rs:451-520: prefix iteration over (prox, w1, prefix2*) — the universe is NOT applied (ByWord branch)
    bool word_prefix_pair_proximity_docids(uint32_t w1, uint32_t prefix2, uint8_t prox, Bitmap &out) {
        const Db &db = index.dbs[DB_WORD_PAIR_PROXIMITY_DOCIDS];
        std::string p = key_pair(prox, words[w1], words[prefix2]);
        uint64_t i = db.lower_bound((const uint8_t *)p.data(), p.size());
        out.clear();
        for (; i < db.n && db.has_prefix(i, p); i++) {
            Span s = db.val(i);
            stat_fetches++;
            stat_fetch_bytes += s.n;
            out.or_with(cbo_decode(s.p, s.n));
        }
        return true;
    }
    bool word_fid_docids(const Bitmap *universe, uint32_t w, uint16_t fid, Bitmap &out) {
        return get_value(DB_WORD_FID_DOCIDS, key_u16(words[w], fid), universe, out);
    }
    bool word_prefix_fid_docids(const Bitmap *universe, uint32_t w, uint16_t fid, Bitmap &out) {
        return get_value(DB_WORD_PREFIX_FID_DOCIDS, key_u16(words[w], fid), universe, out);
    }
    bool word_position_docids(const Bitmap *universe, uint32_t w, uint16_t pos, Bitmap &out) {
        return get_value(DB_WORD_POSITION_DOCIDS, key_u16(words[w], pos), universe, out);
    }
    bool word_prefix_position_docids(const Bitmap *universe, uint32_t w, uint16_t pos, Bitmap &out) {
        return get_value(DB_WORD_PREFIX_POSITION_DOCIDS, key_u16(words[w], pos), universe, out);
    }
    // db_cache.rs:575-718: which (word, u16) keys exist
    std::vector<uint16_t> u16_keys_of(int db_id, uint32_t w) {
        const Db &db = index.dbs[db_id];
        std::string p = words[w];
        p.push_back(0);
        std::vector<uint16_t> out;
        uint64_t i = db.lower_bound((const uint8_t *)p.data(), p.size());
        for (; i < db.n && db.has_prefix(i, p); i++) {
            size_t kn = db.koff[i + 1] - db.koff[i];
            if (kn != p.size() + 2) continue;
            const uint8_t *k = db.keys.data() + db.koff[i] + p.size();
            out.push_back((uint16_t)((k[0] << 8) | k[1]));
        }
        return out;
    }
    std::vector<uint16_t> word_fids(uint32_t w) { return u16_keys_of(DB_WORD_FID_DOCIDS, w); }
    std::vector<uint16_t> word_prefix_fids(uint32_t w) { return u16_keys_of(DB_WORD_PREFIX_FID_DOCIDS, w); }
    std::vector<uint16_t> word_positions(uint32_t w) { return u16_keys_of(DB_WORD_POSITION_DOCIDS, w); }
    std::vector<uint16_t> word_prefix_positions(uint32_t w) { return u16_keys_of(DB_WORD_PREFIX_POSITION_DOCIDS, w); }
};

// ---------------------------------------------------------------- Levenshtein DFA x dictionary
// Walk the sorted dictionary as an implicit trie, carrying the DP rows of the restricted
// Damerau-Levenshtein distance between the query and the current prefix.  Equivalent to
// `fst.search_with_state(dfa)` streaming in lexicographic order.
// cb(word_id, distance, same_first_char) -> false to stop.
struct LevWalk {
    const Index &ix;
    const std::string &q;
    int k_same, k_diff;  // budget when the first char equals / differs (-1: branch excluded)
    bool prefix;
    std::function<bool(uint64_t, int, bool)> cb;
    bool stopped = false;
    uint64_t visited_nodes = 0;

    void run() {
        size_t m = q.size();
        std::vector<int> row0(m + 1);
        for (size_t i = 0; i <= m; i++) row0[i] = (int)i;
        std::vector<int> none;
        uint64_t lo = 0, hi = ix.n_words();
        // the empty word cannot be in the dictionary
        descend(lo, hi, 0, none, row0, 0, -1, (int)m, true);
    }
    uint64_t upper_of_byte(uint64_t lo, uint64_t hi, size_t depth, uint8_t c) {
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (ix.word_ptr(mid)[depth] <= c)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    }
    void descend(uint64_t lo, uint64_t hi, size_t depth, const std::vector<int> &prev2, const std::vector<int> &prev,
                 uint8_t prevc, int k, int best, bool same_first) {
        size_t m = q.size();
        if (lo < hi && ix.word_len(lo) == depth) {
            if (depth > 0) {
                int d = prefix ? best : prev[m];
                if (d <= k) {
                    if (!cb(lo, d, same_first)) {
                        stopped = true;
                        return;
                    }
                }
            }
            lo++;
        }
        while (lo < hi && !stopped) {
            uint8_t c = ix.word_ptr(lo)[depth];
            uint64_t up = upper_of_byte(lo, hi, depth, c);
            int kk = k;
            bool sf = same_first;
            if (depth == 0) {
                sf = ((uint8_t)q[0] == c);
                kk = sf ? k_same : k_diff;
            }
            if (kk >= 0) {
                visited_nodes++;
                std::vector<int> row(m + 1);
                size_t j = depth + 1;
                row[0] = (int)j;
                int mn = row[0];
                for (size_t i = 1; i <= m; i++) {
                    int v = std::min(prev[i] + 1, row[i - 1] + 1);
                    v = std::min(v, prev[i - 1] + ((uint8_t)q[i - 1] != c ? 1 : 0));
                    if (i > 1 && j > 1 && (uint8_t)q[i - 1] == prevc && (uint8_t)q[i - 2] == c) v = std::min(v, prev2[i - 2] + 1);
                    row[i] = v;
                    mn = std::min(mn, v);
                }
                int nbest = std::min(best, row[m]);
                int pmn = prev[0];
                for (size_t i = 1; i <= m; i++) pmn = std::min(pmn, prev[i]);
                bool alive = mn <= kk || pmn <= kk || (prefix && nbest <= kk);
                // distance-0 subtrees of a prefix automaton only yield distance 0, which every caller discards
                if (prefix && nbest == 0) alive = false;
                if (alive) descend(lo, up, depth + 1, prev, row, c, kk, nbest, sf);
            }
            lo = up;
        }
    }
};

// search/mod.rs:558-563
inline char get_first(const std::string &s) { return s[0]; }

// compute_derivations.rs:75-107
inline void find_one_typo_derivations(Ctx &ctx, uint32_t word_interned, bool is_prefix, std::set<uint32_t> &one_typo_words) {
    std::string word = ctx.words[word_interned];
    LevWalk w{ctx.index, word, 1, -1, is_prefix, nullptr};
    w.cb = [&](uint64_t id, int d, bool) {
        if (d == 1) {
            one_typo_words.insert(ctx.intern_word(ctx.index.word(id)));
            if (one_typo_words.size() >= 150) return false;  // limits::MAX_ONE_TYPO_COUNT
        }
        return true;
    };
    w.run();
}

// compute_derivations.rs:109-168
inline void find_one_two_typo_derivations(Ctx &ctx, uint32_t word_interned, bool is_prefix, std::set<uint32_t> &one_typo_words,
                                          std::set<uint32_t> &two_typo_words) {
    std::string word = ctx.words[word_interned];
    LevWalk w{ctx.index, word, 2, 1, is_prefix, nullptr};
    w.cb = [&](uint64_t id, int d, bool same_first) {
        bool finished_one = one_typo_words.size() >= 150;
        bool finished_two = two_typo_words.size() >= 50;  // limits::MAX_TWO_TYPOS_COUNT
        if (finished_one && finished_two) return false;
        if (!same_first && !finished_two) {
            two_typo_words.insert(ctx.intern_word(ctx.index.word(id)));
            return true;
        }
        // second_dfa.distance(..): for a different first char this is the (<=1) distance itself
        switch (d) {
            case 0: break;
            case 1:
                if (!finished_one) one_typo_words.insert(ctx.intern_word(ctx.index.word(id)));
                break;
            case 2:
                if (!finished_two) two_typo_words.insert(ctx.intern_word(ctx.index.word(id)));
                break;
        }
        return true;
    };
    w.run();
}

// parse_query.rs:204-225 — ASCII: chars().count() == len()
inline uint8_t number_of_typos_allowed(const Ctx &ctx, const std::string &word) {
    const Settings &s = ctx.index.settings;
    if (!s.authorize_typos || word.size() < s.min_word_len_one_typo || s.exact_words.count(word)) return 0;
    if (word.size() < s.min_word_len_two_typos) return 1;
    return 2;
}

// compute_derivations.rs:170-253
inline QueryTerm partially_initialized_term_from_word(Ctx &ctx, const std::string &word, uint8_t max_typo, bool is_prefix,
                                                      bool is_ngram) {
    uint32_t word_interned = ctx.intern_word(word);
    QueryTerm t;
    t.original = word_interned;
    if (word.size() > 250) {  // MAX_WORD_LENGTH
        t.one_init = t.two_init = true;
        return t;
    }
    bool use_prefix_db = is_prefix && (ctx.index.dbs[DB_WORD_PREFIX_DOCIDS].get(word).some ||
                                       (!is_ngram && ctx.index.dbs[DB_EXACT_WORD_PREFIX_DOCIDS].get(word).some));
    if (use_prefix_db) t.use_prefix_db = (int32_t)word_interned;
    if (ctx.index.contains_word(word)) t.exact = (int32_t)word_interned;
    if (is_prefix && !use_prefix_db) {
        // find_zero_typo_prefix_derivations :40-73 — merged prefix iteration over word_docids and exact_word_docids
        const Index &ix = ctx.index;
        uint64_t i = ix.dict_lower_bound((const uint8_t *)word.data(), word.size(), 0, ix.n_words());
        for (; i < ix.n_words(); i++) {
            if (ix.word_len(i) < word.size() || memcmp(ix.word_ptr(i), word.data(), word.size()) != 0) break;
            uint32_t d = ctx.intern_word(ix.word(i));
            if (d != word_interned) {
                t.prefix_of.insert(d);
                if (t.prefix_of.size() >= 1000) break;  // limits::MAX_PREFIX_COUNT
            }
        }
    }
    auto it = ctx.index.settings.synonyms.find(std::vector<std::string>{word});
    if (it != ctx.index.settings.synonyms.end()) {
        size_t synonym_word_count = 0, taken = 0;
        for (auto &syn : it->second) {
            if (taken++ >= 50) break;                           // MAX_SYNONYM_PHRASE_COUNT
            if (synonym_word_count + syn.size() > 100) continue;  // MAX_SYNONYM_WORD_COUNT
            synonym_word_count += syn.size();
            Phrase p;
            for (auto &w : syn) p.words.push_back((int32_t)ctx.intern_word(w));
            t.synonyms.insert(ctx.intern_phrase(p));
        }
    }
    t.max_levenshtein_distance = max_typo;
    t.is_prefix = is_prefix;
    return t;
}

// compute_derivations.rs:363-383
inline int32_t find_split_words(Ctx &ctx, const std::string &original) {
    bool have = false;
    uint64_t best = 0;
    uint32_t bl = 0, br = 0;
    for (size_t i = 1; i < original.size(); i++) {
        uint32_t left = ctx.intern_word(original.substr(0, i));
        uint32_t right = ctx.intern_word(original.substr(i));
        uint64_t freq;
        if (ctx.word_pair_proximity_docids_len(left, right, 1, freq)) {
            if (!have || freq > best) {
                have = true;
                best = freq;
                bl = left;
                br = right;
            }
        }
    }
    if (!have) return -1;
    Phrase p;
    p.words = {(int32_t)bl, (int32_t)br};
    return (int32_t)ctx.intern_phrase(p);
}

// compute_derivations.rs:21-37, 264-357
inline void compute_fully_if_needed(Ctx &ctx, uint32_t term_id) {
    QueryTerm &s = ctx.terms[term_id];
    if (s.max_levenshtein_distance <= 1 && !s.one_init) {
        std::set<uint32_t> one;
        if (s.max_levenshtein_distance > 0) find_one_typo_derivations(ctx, s.original, s.is_prefix, one);
        int32_t split = -1;
        if (ctx.terms[term_id].allows_split_words()) {
            std::string orig = ctx.words[ctx.terms[term_id].original];
            split = find_split_words(ctx, orig);
        }
        QueryTerm &t = ctx.terms[term_id];
        if (t.is_ngram && split >= 0) {
            // keep the split only if it differs from the ngram's own component words (:300-311)
            const Phrase &p = ctx.phrases[split];
            std::vector<uint32_t> pw;
            for (auto w : p.words)
                if (w >= 0) pw.push_back((uint32_t)w);
            if (pw == t.ngram_words) split = -1;
        }
        t.split_words = split;
        t.one_typo = std::move(one);
        t.one_init = true;
        t.two_init = true;
    } else if (s.max_levenshtein_distance > 1 && !s.two_init) {
        std::set<uint32_t> one, two;
        find_one_two_typo_derivations(ctx, s.original, s.is_prefix, one, two);
        std::string orig = ctx.words[ctx.terms[term_id].original];
        int32_t split = find_split_words(ctx, orig);
        QueryTerm &t = ctx.terms[term_id];
        t.one_typo = std::move(one);
        t.two_typos = std::move(two);
        t.split_words = split;
        t.one_init = t.two_init = true;
    }
}

// ---------------------------------------------------------------- QueryTermSubset accessors (query_term/mod.rs:124-406)
struct ExactTerm {
    bool some = false;
    bool is_phrase = false;
    uint32_t id = 0;
};
inline ExactTerm exact_term(const Ctx &ctx, const QueryTermSubset &s) {
    const QueryTerm &t = ctx.terms[s.original];
    ExactTerm e;
    if (t.is_ngram) return e;
    if (t.phrase >= 0) {
        if (s.zero.contains_phrase((uint32_t)t.phrase)) e = {true, true, (uint32_t)t.phrase};
    } else if (t.exact >= 0) {
        if (s.zero.contains_word((uint32_t)t.exact)) e = {true, false, (uint32_t)t.exact};
    }
    return e;
}
inline bool use_prefix_db(const Ctx &ctx, const QueryTermSubset &s, Word &out) {
    const QueryTerm &t = ctx.terms[s.original];
    if (t.use_prefix_db < 0) return false;
    uint32_t w = (uint32_t)t.use_prefix_db;
    bool ok = s.zero.kind == N_ALL || (s.zero.kind == N_SUBSET && s.zero.words.count(w));
    if (!ok) return false;
    out = Word{t.is_ngram ? W_DERIVED : W_ORIGINAL, w};
    return true;
}
inline std::set<Word> all_single_words_except_prefix_db(Ctx &ctx, const QueryTermSubset &s) {
    std::set<Word> result;
    if (!s.one.is_empty() || !s.two.is_empty()) compute_fully_if_needed(ctx, s.original);
    const QueryTerm &t = ctx.terms[s.original];
    int zk = t.is_ngram ? W_DERIVED : W_ORIGINAL;
    if (s.zero.kind == N_ALL) {
        if (t.exact >= 0) result.insert(Word{zk, (uint32_t)t.exact});
        for (auto w : t.prefix_of) result.insert(Word{zk, w});
    } else if (s.zero.kind == N_SUBSET) {
        if (t.exact >= 0 && s.zero.words.count((uint32_t)t.exact)) result.insert(Word{zk, (uint32_t)t.exact});
        for (auto w : t.prefix_of)
            if (s.zero.words.count(w)) result.insert(Word{zk, w});
    }
    if (s.one.kind == N_ALL) {
        for (auto w : t.one_typo) result.insert(Word{W_DERIVED, w});
    } else if (s.one.kind == N_SUBSET) {
        for (auto w : t.one_typo)
            if (s.one.words.count(w)) result.insert(Word{W_DERIVED, w});
    }
    if (s.two.kind == N_ALL) {
        for (auto w : t.two_typos) result.insert(Word{W_DERIVED, w});
    } else if (s.two.kind == N_SUBSET) {
        for (auto w : t.two_typos)
            if (s.two.words.count(w)) result.insert(Word{W_DERIVED, w});
    }
    return result;
}
inline std::set<uint32_t> all_phrases(Ctx &ctx, const QueryTermSubset &s) {
    std::set<uint32_t> result;
    if (!s.one.is_empty()) compute_fully_if_needed(ctx, s.original);
    const QueryTerm &t = ctx.terms[s.original];
    // NB: the zero-typo phrase and the synonyms are added regardless of zero_typo_subset (:301-304)
    if (t.phrase >= 0) result.insert((uint32_t)t.phrase);
    for (auto p : t.synonyms) result.insert(p);
    if (s.one.kind == N_ALL) {
        if (t.split_words >= 0) result.insert((uint32_t)t.split_words);
    } else if (s.one.kind == N_SUBSET) {
        if (t.split_words >= 0 && s.one.phrases.count((uint32_t)t.split_words)) result.insert((uint32_t)t.split_words);
    }
    return result;
}
inline int32_t original_phrase(const Ctx &ctx, const QueryTermSubset &s) {
    const QueryTerm &t = ctx.terms[s.original];
    if (t.phrase >= 0 && s.zero.contains_phrase((uint32_t)t.phrase)) return t.phrase;
    return -1;
}
inline uint8_t max_typo_cost(const Ctx &ctx, const QueryTermSubset &s) {
    const QueryTerm &t = ctx.terms[s.original];
    switch (t.max_levenshtein_distance) {
        case 0: return t.allows_split_words() ? 1 : 0;
        case 1: return s.one.is_empty() ? 0 : 1;
        default: return s.two.is_empty() ? (s.one.is_empty() ? 0 : 1) : 2;
    }
}
inline void keep_only_exact_term(const Ctx &ctx, QueryTermSubset &s) {
    ExactTerm e = exact_term(ctx, s);
    if (!e.some) return;
    s.zero = NTypoSubset{};
    s.zero.kind = N_SUBSET;
    if (e.is_phrase)
        s.zero.phrases.insert(e.id);
    else
        s.zero.words.insert(e.id);
    s.one = NTypoSubset{};
    s.two = NTypoSubset{};
}

}  // namespace orc
