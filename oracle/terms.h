// ORACLE — test infrastructure only. See bitmap.h.
// Query terms, typo/prefix/split derivations, query graph.
// Follows crates/milli/src/search/new/query_term/{mod.rs,ntypo_subset.rs,compute_derivations.rs,parse_query.rs},
// crates/milli/src/search/new/query_graph.rs, crates/milli/src/search/mod.rs:558-577.
#pragma once
#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "index.h"

namespace orc {

// ---------------------------------------------------------------- small dynamic bitset (SmallBitmap, small_bitmap.rs)
struct Bits {
    std::vector<uint64_t> w;
    uint32_t n = 0;
    Bits() {}
    explicit Bits(uint32_t n_) : w((n_ + 63) / 64, 0), n(n_) {}
    void insert(uint32_t i) { w[i >> 6] |= 1ull << (i & 63); }
    void remove(uint32_t i) { w[i >> 6] &= ~(1ull << (i & 63)); }
    bool contains(uint32_t i) const { return (i >> 6) < w.size() && ((w[i >> 6] >> (i & 63)) & 1); }
    void clear() { std::fill(w.begin(), w.end(), 0); }
    bool is_empty() const {
        for (auto x : w)
            if (x) return false;
        return true;
    }
    void union_with(const Bits &o) {
        for (size_t i = 0; i < w.size() && i < o.w.size(); i++) w[i] |= o.w[i];
    }
    bool intersects(const Bits &o) const {
        for (size_t i = 0; i < w.size() && i < o.w.size(); i++)
            if (w[i] & o.w[i]) return true;
        return false;
    }
    bool is_subset(const Bits &o) const {
        for (size_t i = 0; i < w.size(); i++) {
            uint64_t ow = i < o.w.size() ? o.w[i] : 0;
            if (w[i] & ~ow) return false;
        }
        return true;
    }
    std::vector<uint32_t> items() const {
        std::vector<uint32_t> v;
        for (size_t i = 0; i < w.size(); i++) {
            uint64_t b = w[i];
            while (b) {
                v.push_back((uint32_t)(i * 64 + __builtin_ctzll(b)));
                b &= b - 1;
            }
        }
        return v;
    }
};

// ---------------------------------------------------------------- interned data
struct Phrase {
    std::vector<int32_t> words;  // interned word id, -1 = stop word hole
    bool operator<(const Phrase &o) const { return words < o.words; }
};

enum NKind { N_ALL = 0, N_SUBSET = 1, N_NOTHING = 2 };
struct NTypoSubset {  // ntypo_subset.rs
    int kind = N_NOTHING;
    std::set<uint32_t> words, phrases;
    bool contains_word(uint32_t w) const { return kind == N_ALL || (kind == N_SUBSET && words.count(w)); }
    bool contains_phrase(uint32_t p) const { return kind == N_ALL || (kind == N_SUBSET && phrases.count(p)); }
    bool is_empty() const { return kind == N_NOTHING || (kind == N_SUBSET && words.empty() && phrases.empty()); }
    void union_with(const NTypoSubset &o) {
        if (kind == N_ALL) return;
        if (kind == N_SUBSET) {
            if (o.kind == N_ALL) {
                *this = NTypoSubset{N_ALL, {}, {}};
            } else if (o.kind == N_SUBSET) {
                words.insert(o.words.begin(), o.words.end());
                phrases.insert(o.phrases.begin(), o.phrases.end());
            }
            return;
        }
        *this = o;
    }
    void intersect(const NTypoSubset &o) {
        if (kind == N_ALL) {
            *this = o;
        } else if (kind == N_SUBSET) {
            if (o.kind == N_SUBSET) {
                std::set<uint32_t> ws, ps;
                for (auto x : words)
                    if (o.words.count(x)) ws.insert(x);
                for (auto x : phrases)
                    if (o.phrases.count(x)) ps.insert(x);
                words.swap(ws);
                phrases.swap(ps);
            } else if (o.kind == N_NOTHING)
                *this = NTypoSubset{};
        }
    }
    bool operator==(const NTypoSubset &o) const { return kind == o.kind && words == o.words && phrases == o.phrases; }
    void key(std::string &s) const {
        s += (char)('A' + kind);
        for (auto x : words) s += "w" + std::to_string(x);
        for (auto x : phrases) s += "p" + std::to_string(x);
        s += ';';
    }
};

struct QueryTerm {  // query_term/mod.rs:43-81
    uint32_t original = 0;  // interned word
    bool is_ngram = false;
    std::vector<uint32_t> ngram_words;
    uint8_t max_levenshtein_distance = 0;
    bool is_prefix = false;
    // ZeroTypoTerm
    int32_t phrase = -1;
    int32_t exact = -1;
    std::set<uint32_t> prefix_of;
    std::set<uint32_t> synonyms;  // phrases
    int32_t use_prefix_db = -1;
    // OneTypoTerm (lazy)
    bool one_init = false;
    int32_t split_words = -1;
    std::set<uint32_t> one_typo;
    // TwoTypoTerm (lazy)
    bool two_init = false;
    std::set<uint32_t> two_typos;

    bool allows_split_words() const { return phrase < 0; }
    bool is_empty() const {
        if (!one_init || !two_init) return false;
        return phrase < 0 && exact < 0 && prefix_of.empty() && synonyms.empty() && use_prefix_db < 0 && one_typo.empty() &&
               split_words < 0 && two_typos.empty();
    }
};

enum WordKind { W_ORIGINAL = 0, W_DERIVED = 1 };
struct Word {
    int kind;
    uint32_t id;
    bool operator<(const Word &o) const { return kind != o.kind ? kind < o.kind : id < o.id; }
};

struct QueryTermSubset {
    uint32_t original = 0;  // index in term store
    NTypoSubset zero, one, two;
    bool mandatory = false;
    static QueryTermSubset full(uint32_t t) {
        QueryTermSubset s;
        s.original = t;
        s.zero.kind = s.one.kind = s.two.kind = N_ALL;
        return s;
    }
    void intersect(const QueryTermSubset &o) {
        zero.intersect(o.zero);
        one.intersect(o.one);
        two.intersect(o.two);
    }
    bool operator==(const QueryTermSubset &o) const {
        return original == o.original && zero == o.zero && one == o.one && two == o.two && mandatory == o.mandatory;
    }
    void key(std::string &s) const {
        s += "T" + std::to_string(original) + (mandatory ? "!" : ".");
        zero.key(s);
        one.key(s);
        two.key(s);
    }
};

struct LocatedQueryTermSubset {
    QueryTermSubset term_subset;
    uint16_t pos_start = 0, pos_end = 0;
    uint8_t tid_start = 0, tid_end = 0;
    uint32_t term_ids_len() const { return (uint32_t)tid_end - tid_start + 1; }
    uint32_t positions_len() const { return (uint32_t)pos_end - pos_start + 1; }
    bool operator==(const LocatedQueryTermSubset &o) const {
        return term_subset == o.term_subset && pos_start == o.pos_start && pos_end == o.pos_end && tid_start == o.tid_start &&
               tid_end == o.tid_end;
    }
    std::string key() const {
        std::string s;
        term_subset.key(s);
        s += "@" + std::to_string(pos_start) + "-" + std::to_string(pos_end) + "#" + std::to_string(tid_start) + "-" +
             std::to_string(tid_end);
        return s;
    }
};

struct LocatedQueryTerm {
    uint32_t value;  // term id
    uint16_t pos_start, pos_end;
};

// ---------------------------------------------------------------- query graph
enum NodeKind { NODE_TERM = 0, NODE_DELETED = 1, NODE_START = 2, NODE_END = 3 };
struct QueryNode {
    int kind = NODE_DELETED;
    LocatedQueryTermSubset term;
    Bits predecessors, successors;
};
struct QueryGraph {
    uint32_t root_node = 0, end_node = 1;
    std::vector<QueryNode> nodes;
};

// query tokens, as charabia would hand them over (search/new/mod.rs:918-996)
enum TokenKind { TOK_WORD = 0, TOK_STOPWORD = 1, TOK_SEP_SOFT = 2, TOK_SEP_HARD = 3 };
struct Token {
    int kind;
    std::string lemma;
};

}  // namespace orc
