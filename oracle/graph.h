// ORACLE — test infrastructure only. See bitmap.h.
// Query parsing + query graph + term docids resolution.
// Follows query_term/parse_query.rs:28-300, query_graph.rs (all), resolve_query_graph.rs (all).
#pragma once
#include <deque>

#include "context.h"

namespace orc {

// ---------------------------------------------------------------- parse_query.rs:227-300
inline bool make_ngram(Ctx &ctx, const std::vector<LocatedQueryTerm> &terms, size_t from, size_t to /*inclusive*/,
                       LocatedQueryTerm &out) {
    for (size_t i = from; i <= to; i++)
        if (ctx.terms[terms[i].value].phrase >= 0) return false;
    for (size_t i = from; i < to; i++)
        if (terms[i].pos_end != (uint16_t)(terms[i + 1].pos_start - 1)) return false;
    std::vector<uint32_t> words_interned;
    for (size_t i = from; i <= to; i++) {
        const QueryTerm &t = ctx.terms[terms[i].value];
        if (t.is_ngram) return false;  // original_single_word
        words_interned.push_back(t.original);
    }
    std::vector<std::string> words;
    std::string ngram_str;
    for (auto w : words_interned) {
        words.push_back(ctx.words[w]);
        ngram_str += ctx.words[w];
    }
    uint16_t start = terms[from].pos_start, end = terms[to].pos_end;
    bool is_prefix = ctx.terms[terms[to].value].is_prefix;
    if (ngram_str.size() > 250) return false;
    uint32_t ngram_interned = ctx.intern_word(ngram_str);
    uint8_t n = number_of_typos_allowed(ctx, ngram_str);
    uint8_t dec = (uint8_t)(to - from);
    uint8_t max_nbr_typos = n > dec ? (uint8_t)(n - dec) : 0;
    QueryTerm term = partially_initialized_term_from_word(ctx, ngram_str, max_nbr_typos, is_prefix, true);
    auto it = ctx.index.settings.synonyms.find(words);
    if (it != ctx.index.settings.synonyms.end()) {
        for (auto &syn : it->second) {
            Phrase p;
            for (auto &w : syn) p.words.push_back((int32_t)ctx.intern_word(w));
            term.synonyms.insert(ctx.intern_phrase(p));
        }
    }
    term.original = ngram_interned;
    term.is_ngram = true;
    term.ngram_words = words_interned;
    term.is_prefix = is_prefix;
    term.max_levenshtein_distance = max_nbr_typos;
    term.one_init = term.two_init = false;
    ctx.terms.push_back(term);
    out = LocatedQueryTerm{(uint32_t)ctx.terms.size() - 1, start, end};
    return true;
}

// ---------------------------------------------------------------- query_graph.rs
inline void build_initial_edges(QueryGraph &g) {
    uint32_t n = (uint32_t)g.nodes.size();
    for (auto &node : g.nodes) {
        node.successors = Bits(n);
        node.predecessors = Bits(n);
    }
    for (uint32_t id = 0; id < n; id++) {
        int end_prev;
        const QueryNode &node = g.nodes[id];
        if (node.kind == NODE_TERM)
            end_prev = node.term.tid_end;
        else if (node.kind == NODE_START)
            end_prev = -1;
        else
            continue;
        Bits succ(n);
        int mn = 32767;
        for (uint32_t j = 0; j < n; j++) {
            const QueryNode &o = g.nodes[j];
            int start_next;
            if (o.kind == NODE_TERM)
                start_next = o.term.tid_start;
            else if (o.kind == NODE_END)
                start_next = 32767;
            else
                continue;
            if (start_next <= end_prev) continue;
            if (start_next < mn) {
                mn = start_next;
                succ.clear();
                succ.insert(j);
            } else if (start_next == mn)
                succ.insert(j);
        }
        g.nodes[id].successors = succ;
        for (auto s : succ.items()) g.nodes[s].predecessors.insert(id);
    }
}

inline QueryGraph graph_from_query(Ctx &ctx, std::vector<LocatedQueryTerm> &terms /* in: words; out: + ngrams */) {
    std::vector<LocatedQueryTerm> base = terms;
    QueryGraph g;
    auto add = [&](int kind, LocatedQueryTermSubset t) {
        QueryNode n;
        n.kind = kind;
        n.term = std::move(t);
        g.nodes.push_back(std::move(n));
        return (uint32_t)g.nodes.size() - 1;
    };
    add(NODE_START, {});
    add(NODE_END, {});
    bool prev1 = false, prev2 = false, prev0 = true;
    for (size_t i = 0; i < base.size(); i++) {
        LocatedQueryTermSubset t;
        t.term_subset = QueryTermSubset::full(base[i].value);
        t.pos_start = base[i].pos_start;
        t.pos_end = base[i].pos_end;
        t.tid_start = t.tid_end = (uint8_t)i;
        add(NODE_TERM, t);
        if (prev1) {
            LocatedQueryTerm ng;
            if (make_ngram(ctx, base, i - 1, i, ng)) {
                terms.push_back(ng);
                LocatedQueryTermSubset s;
                s.term_subset = QueryTermSubset::full(ng.value);
                s.pos_start = ng.pos_start;
                s.pos_end = ng.pos_end;
                s.tid_start = (uint8_t)(i - 1);
                s.tid_end = (uint8_t)i;
                add(NODE_TERM, s);
            }
        }
        if (prev2) {
            LocatedQueryTerm ng;
            if (make_ngram(ctx, base, i - 2, i, ng)) {
                terms.push_back(ng);
                LocatedQueryTermSubset s;
                s.term_subset = QueryTermSubset::full(ng.value);
                s.pos_start = ng.pos_start;
                s.pos_end = ng.pos_end;
                s.tid_start = (uint8_t)(i - 2);
                s.tid_end = (uint8_t)i;
                add(NODE_TERM, s);
            }
        }
        // (prev0, prev1, prev2) = (new_nodes, prev0, prev1): non-emptiness is all that is used
        prev2 = prev1;
        prev1 = prev0;
        prev0 = true;
    }
    build_initial_edges(g);
    return g;
}

inline void remove_nodes_keep_edges(QueryGraph &g, const std::vector<uint32_t> &nodes) {
    for (auto id : nodes) {
        Bits pred = g.nodes[id].predecessors, succ = g.nodes[id].successors;
        for (auto p : pred.items()) {
            g.nodes[p].successors.remove(id);
            g.nodes[p].successors.union_with(succ);
        }
        for (auto s : succ.items()) {
            g.nodes[s].predecessors.remove(id);
            g.nodes[s].predecessors.union_with(pred);
        }
        g.nodes[id].kind = NODE_DELETED;
        g.nodes[id].predecessors.clear();
        g.nodes[id].successors.clear();
    }
}

// query_graph.rs:379-406
inline std::vector<Bits> removal_order(const Ctx &ctx, const QueryGraph &g, const std::function<uint16_t(uint8_t)> &order) {
    std::map<uint16_t, Bits> to_remove;
    bool mandatory = false;
    for (uint32_t id = 0; id < g.nodes.size(); id++) {
        const QueryNode &n = g.nodes[id];
        if (n.kind != NODE_TERM) continue;
        if (original_phrase(ctx, n.term.term_subset) >= 0 || n.term.term_subset.mandatory) {
            mandatory = true;
            continue;
        }
        uint16_t cost = 0;
        for (int t = n.term.tid_start; t <= n.term.tid_end; t++) cost = std::max(cost, order((uint8_t)t));
        auto it = to_remove.find(cost);
        if (it == to_remove.end()) it = to_remove.emplace(cost, Bits((uint32_t)g.nodes.size())).first;
        it->second.insert(id);
    }
    std::vector<Bits> res;
    for (auto &kv : to_remove) res.push_back(kv.second);
    if (!mandatory && !res.empty()) res.pop_back();
    return res;
}
// query_graph.rs:346-377
inline std::vector<Bits> removal_order_last(const Ctx &ctx, const QueryGraph &g) {
    uint8_t first = 255, last = 0;
    for (auto &n : g.nodes)
        if (n.kind == NODE_TERM) {
            if (n.term.tid_end > last) last = n.term.tid_end;
            if (n.term.tid_start < first) first = n.term.tid_start;
        }
    if (first >= last) return {};
    return removal_order(ctx, g, [&](uint8_t t) { return (uint16_t)(1 + last - t); });
}

inline size_t words_in_phrases_count(const Ctx &ctx, const QueryGraph &g) {
    size_t c = 0;
    for (auto &n : g.nodes)
        if (n.kind == NODE_TERM) {
            int32_t p = original_phrase(ctx, n.term.term_subset);
            if (p < 0) continue;
            for (auto w : ctx.phrases[p].words)
                if (w >= 0) c++;
        }
    return c;
}

// ---------------------------------------------------------------- resolve_query_graph.rs
inline Bitmap compute_phrase_docids(Ctx &ctx, uint32_t phrase);
inline const Bitmap &get_phrase_docids(Ctx &ctx, uint32_t phrase) {
    auto it = ctx.phrase_docids.find(phrase);
    if (it != ctx.phrase_docids.end()) return it->second;
    Bitmap d = compute_phrase_docids(ctx, phrase);
    return ctx.phrase_docids.emplace(phrase, std::move(d)).first->second;
}

// :187-268
inline Bitmap compute_phrase_docids(Ctx &ctx, uint32_t phrase) {
    std::vector<int32_t> words = ctx.phrases[phrase].words;
    if (words.empty()) return Bitmap();
    bool have = false;
    Bitmap candidates;
    for (auto w : words) {
        if (w < 0) continue;
        Bitmap wd;
        if (ctx.word_docids(nullptr, Word{W_ORIGINAL, (uint32_t)w}, wd)) {
            if (have)
                candidates.and_with(wd);
            else {
                candidates = std::move(wd);
                have = true;
            }
        } else
            return Bitmap();
    }
    if (!have) return Bitmap();
    size_t winsize = std::min<size_t>(words.size(), 3);
    for (size_t ws = 0; ws + winsize <= words.size(); ws++) {
        std::vector<Bitmap> bitmaps;
        for (size_t offset = 0; offset < winsize; offset++) {
            if (words[ws + offset] < 0) continue;
            uint32_t s1 = (uint32_t)words[ws + offset];
            for (size_t k = offset + 1; k < winsize; k++) {
                if (words[ws + k] < 0) continue;
                uint32_t s2 = (uint32_t)words[ws + k];
                size_t dist = k - offset - 1;
                if (dist == 0) {
                    Bitmap m;
                    if (ctx.word_pair_proximity_docids(nullptr, s1, s2, 1, m))
                        bitmaps.push_back(std::move(m));
                    else
                        return Bitmap();
                } else {
                    Bitmap bitmap;
                    for (size_t d = 0; d <= dist; d++) {
                        Bitmap m;
                        if (ctx.word_pair_proximity_docids(nullptr, s1, s2, (uint8_t)(d + 1), m)) bitmap.or_with(m);
                    }
                    if (bitmap.is_empty()) return bitmap;
                    bitmaps.push_back(std::move(bitmap));
                }
            }
        }
        std::stable_sort(bitmaps.begin(), bitmaps.end(), [](const Bitmap &a, const Bitmap &b) { return a.len() < b.len(); });
        for (auto &bm : bitmaps) {
            candidates.and_with(bm);
            if (candidates.is_empty()) break;
        }
    }
    return candidates;
}

// :33-59
inline Bitmap compute_query_term_subset_docids(Ctx &ctx, const Bitmap *universe, const QueryTermSubset &term) {
    Bitmap docids;
    for (auto w : all_single_words_except_prefix_db(ctx, term)) {
        Bitmap wd;
        if (ctx.word_docids(universe, w, wd)) docids.or_with(wd);
    }
    for (auto p : all_phrases(ctx, term)) docids.or_with(get_phrase_docids(ctx, p));
    Word pw;
    if (use_prefix_db(ctx, term, pw)) {
        Bitmap pd;
        if (ctx.word_prefix_docids(universe, pw, pd)) docids.or_with(pd);
    }
    if (universe) docids.and_with(*universe);
    return docids;
}
// :61-93
inline Bitmap compute_query_term_subset_docids_within_field_id(Ctx &ctx, const Bitmap *universe, const QueryTermSubset &term,
                                                               uint16_t fid) {
    Bitmap docids;
    for (auto w : all_single_words_except_prefix_db(ctx, term)) {
        Bitmap wd;
        if (ctx.word_fid_docids(universe, w.id, fid, wd)) docids.or_with(wd);
    }
    for (auto p : all_phrases(ctx, term)) {
        int32_t first = -1;
        for (auto w : ctx.phrases[p].words)
            if (w >= 0) {
                first = w;
                break;
            }
        if (first < 0) continue;
        Bitmap wd;
        if (ctx.word_fid_docids(universe, (uint32_t)first, fid, wd)) docids.or_with(bm_and(get_phrase_docids(ctx, p), wd));
    }
    Word pw;
    if (use_prefix_db(ctx, term, pw)) {
        Bitmap pd;
        if (ctx.word_prefix_fid_docids(universe, pw.id, fid, pd)) docids.or_with(pd);
    }
    return docids;
}
// :95-130
inline Bitmap compute_query_term_subset_docids_within_position(Ctx &ctx, const Bitmap *universe, const QueryTermSubset &term,
                                                               uint16_t position) {
    Bitmap docids;
    for (auto w : all_single_words_except_prefix_db(ctx, term)) {
        Bitmap wd;
        if (ctx.word_position_docids(universe, w.id, position, wd)) docids.or_with(wd);
    }
    for (auto p : all_phrases(ctx, term)) {
        int32_t first = -1;
        for (auto w : ctx.phrases[p].words)
            if (w >= 0) {
                first = w;
                break;
            }
        if (first < 0) continue;
        Bitmap wd;
        if (ctx.word_position_docids(universe, (uint32_t)first, position, wd))
            docids.or_with(bm_and(get_phrase_docids(ctx, p), wd));
    }
    Word pw;
    if (use_prefix_db(ctx, term, pw)) {
        Bitmap pd;
        if (ctx.word_prefix_position_docids(universe, pw.id, position, pd)) docids.or_with(pd);
    }
    return docids;
}

// :133-185
inline Bitmap compute_query_graph_docids(Ctx &ctx, const QueryGraph &q, const Bitmap &universe) {
    uint32_t n = (uint32_t)q.nodes.size();
    Bits resolved(n);
    std::vector<Bitmap> path_nodes_docids(n);
    std::deque<uint32_t> next;
    next.push_back(q.root_node);
    while (!next.empty()) {
        uint32_t id = next.front();
        next.pop_front();
        const QueryNode &node = q.nodes[id];
        if (!node.predecessors.is_subset(resolved)) {
            next.push_back(id);
            continue;
        }
        Bitmap preds;
        for (auto p : node.predecessors.items()) preds.or_with(path_nodes_docids[p]);
        Bitmap node_docids;
        if (node.kind == NODE_TERM)
            node_docids = compute_query_term_subset_docids(ctx, &preds, node.term.term_subset);
        else if (node.kind == NODE_START)
            node_docids = universe;
        else if (node.kind == NODE_END)
            return preds;
        else
            throw std::runtime_error("deleted node in compute_query_graph_docids");
        resolved.insert(id);
        path_nodes_docids[id] = std::move(node_docids);
        for (auto s : node.successors.items()) {
            if (std::find(next.begin(), next.end(), s) == next.end() && !resolved.contains(s)) next.push_back(s);
        }
        for (auto p : node.predecessors.items())
            if (q.nodes[p].successors.is_subset(resolved)) path_nodes_docids[p].clear();
    }
    throw std::runtime_error("compute_query_graph_docids: end not reached");
}

// query_graph.rs:303-344
inline std::vector<Bits> removal_order_frequency(Ctx &ctx, const QueryGraph &g) {
    std::map<uint8_t, Bitmap> term_docids;
    for (auto &n : g.nodes) {
        if (n.kind != NODE_TERM) continue;
        Bitmap d = compute_query_term_subset_docids(ctx, nullptr, n.term.term_subset);
        for (int id = n.term.tid_start; id <= n.term.tid_end; id++) {
            auto it = term_docids.find((uint8_t)id);
            if (it == term_docids.end())
                term_docids.emplace((uint8_t)id, d);
            else
                it->second.or_with(d);
        }
    }
    std::vector<std::pair<uint8_t, uint64_t>> twf;
    for (auto &kv : term_docids) twf.push_back({kv.first, kv.second.len() == 0 ? UINT64_MAX : kv.second.len()});
    std::stable_sort(twf.begin(), twf.end(), [](auto &a, auto &b) { return a.second > b.second; });
    std::map<uint8_t, uint16_t> weight_of;
    uint16_t weight = 1;
    for (size_t i = 0; i < twf.size(); i++) {
        weight_of[twf[i].first] = weight;
        if (i + 1 < twf.size() && twf[i].second != twf[i + 1].second) weight++;
    }
    return removal_order(ctx, g, [&](uint8_t t) { return weight_of.at(t); });
}

// query_graph.rs:453-543
inline QueryGraph build_from_paths(
    const std::vector<std::vector<std::pair<std::pair<bool, LocatedQueryTermSubset>, LocatedQueryTermSubset>>> &paths) {
    std::vector<std::vector<LocatedQueryTermSubset>> single;
    for (auto &path : paths) {
        std::vector<LocatedQueryTermSubset> processed;
        bool have_prev = false;
        LocatedQueryTermSubset prev;
        for (auto &step : path) {
            bool has_start = step.first.first;
            if (have_prev) {
                if (has_start) {
                    LocatedQueryTermSubset start = step.first.second;
                    if (start.tid_start == prev.tid_start && start.tid_end == prev.tid_end) {
                        start.term_subset.intersect(prev.term_subset);
                        processed.push_back(start);
                    } else {
                        processed.push_back(prev);
                        processed.push_back(start);
                    }
                } else
                    processed.push_back(prev);
            } else if (has_start)
                processed.push_back(step.first.second);
            prev = step.second;
            have_prev = true;
        }
        if (have_prev) processed.push_back(prev);
        single.push_back(std::move(processed));
    }
    // node identity = (term, suffix of the path from that term on); the reference keys on an FxHash of the
    // suffix, which is the same identity up to hash collisions.
    QueryGraph g;
    g.nodes.resize(2);
    g.nodes[0].kind = NODE_START;
    g.nodes[1].kind = NODE_END;
    std::map<std::string, uint32_t> ids;
    std::vector<std::vector<uint32_t>> paths_with_ids;
    for (auto &path : single) {
        std::vector<std::string> suffix(path.size());
        std::string acc;
        for (size_t i = path.size(); i-- > 0;) {
            acc = path[i].key() + "|" + acc;
            suffix[i] = acc;
        }
        std::vector<uint32_t> pid;
        for (size_t i = 0; i < path.size(); i++) {
            auto it = ids.find(suffix[i]);
            if (it == ids.end()) {
                QueryNode n;
                n.kind = NODE_TERM;
                n.term = path[i];
                g.nodes.push_back(n);
                it = ids.emplace(suffix[i], (uint32_t)g.nodes.size() - 1).first;
            }
            pid.push_back(it->second);
        }
        paths_with_ids.push_back(std::move(pid));
    }
    uint32_t n = (uint32_t)g.nodes.size();
    for (auto &node : g.nodes) {
        node.predecessors = Bits(n);
        node.successors = Bits(n);
    }
    for (auto &path : paths_with_ids) {
        uint32_t prev = g.root_node;
        for (auto id : path) {
            g.nodes[prev].successors.insert(id);
            g.nodes[id].predecessors.insert(prev);
            prev = id;
        }
        g.nodes[prev].successors.insert(g.end_node);
        g.nodes[g.end_node].predecessors.insert(prev);
    }
    return g;
}

}  // namespace orc
