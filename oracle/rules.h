// ORACLE — test infrastructure only. See bitmap.h.
// Ranking-rule graph, the six graph rules, ExactAttribute, VectorSort, bucket_sort.
// Follows search/new/ranking_rules.rs, graph_based_ranking_rule.rs, ranking_rule_graph/** ,
// exact_attribute.rs, vector_sort.rs, bucket_sort.rs, score_details.rs:441-566.
#pragma once
#include <chrono>
#include <cmath>
#include <memory>

#include "graph.h"

namespace orc {

enum ScoreKind { S_WORDS = 0, S_TYPO, S_PROXIMITY, S_FID, S_POSITION, S_EXACT_ATTRIBUTE, S_EXACT_WORDS, S_VECTOR, S_SKIPPED };
struct Score {
    int kind = S_SKIPPED;
    uint32_t rank = 0, max_rank = 1;  // score_details.rs:512-547 `Rank`, as returned by ScoreDetails::rank()
    bool has_similarity = false;
    float similarity = 0;
};
inline void rank_merge(uint64_t &orank, uint64_t &omax, uint64_t irank, uint64_t imax) {  // Rank::merge :537-546
    orank = orank > 0 ? orank - 1 : 0;
    orank *= imax;
    omax *= imax;
    orank += irank;
}
inline double global_score(const std::vector<Score> &details) {  // ScoreDetails::global_score :133-154
    uint64_t r = 1, m = 1;
    bool has_sem = false;
    double sem = 0;
    for (auto &s : details) {
        if (s.kind == S_VECTOR) {
            has_sem = true;
            sem = s.has_similarity ? (double)s.similarity : 0.0;
        } else
            rank_merge(r, m, s.rank, s.max_rank);
    }
    return has_sem ? sem : (double)r / (double)m;
}

enum RuleKind { R_WORDS = 0, R_TYPO, R_PROXIMITY, R_FID, R_POSITION, R_EXACTNESS };

struct Condition {
    int rule = 0;
    LocatedQueryTermSubset term;  // dest / right term
    // typo
    uint8_t nbr_typos = 0;
    // proximity
    bool prox_uninit = false;
    LocatedQueryTermSubset left_term;
    uint8_t cost = 0;
    // fid
    bool has_fid = false;
    uint16_t fid = 0;
    // position
    std::vector<uint16_t> positions;
    // exactness
    bool exact_in_attribute = false;
    std::string key() const {
        std::string s = std::to_string(rule) + ":" + term.key() + ":" + std::to_string(nbr_typos) + (prox_uninit ? "U" : "T");
        if (prox_uninit) s += left_term.key() + "c" + std::to_string(cost);
        s += has_fid ? "f" + std::to_string(fid) : "f-";
        for (auto p : positions) s += "," + std::to_string(p);
        s += exact_in_attribute ? "E" : "A";
        return s;
    }
};

struct ComputedCondition {
    Bitmap docids;
    uint64_t universe_len = 0;
    bool has_start = false;
    LocatedQueryTermSubset start_term_subset, end_term_subset;
};

struct Edge {
    bool alive = true;
    uint32_t source_node, dest_node;
    uint32_t cost;
    int32_t condition;  // -1 = none
    Bits nodes_to_skip;
};

struct RankingRuleGraph {
    QueryGraph query_graph;
    std::vector<Edge> edges_store;
    std::vector<Bits> edges_of_node;
    std::vector<Condition> conditions;
};

// ---------------------------------------------------------------- per-rule build_edges / resolve_condition
inline uint32_t position_cost_from_distance(uint32_t d) {  // position/mod.rs:129-143
    if (d == 0) return 0;
    if (d == 1) return 1;
    if (d <= 4) return 2;
    if (d <= 7) return 3;
    if (d <= 11) return 4;
    if (d <= 16) return 5;
    if (d <= 24) return 6;
    if (d <= 64) return 7;
    if (d <= 256) return 8;
    if (d <= 1024) return 9;
    return 10;
}

struct CondInterner {
    std::vector<Condition> items;
    std::map<std::string, uint32_t> ids;
    uint32_t insert(const Condition &c) {
        std::string k = c.key();
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        items.push_back(c);
        ids.emplace(k, (uint32_t)items.size() - 1);
        return (uint32_t)items.size() - 1;
    }
};

inline std::vector<std::pair<uint32_t, uint32_t>> build_edges(Ctx &ctx, int rule, CondInterner &ci, const LocatedQueryTermSubset *from,
                                                              const LocatedQueryTermSubset &to) {
    std::vector<std::pair<uint32_t, uint32_t>> edges;
    switch (rule) {
        case R_WORDS: {  // words/mod.rs:39-47
            Condition c;
            c.rule = rule;
            c.term = to;
            edges.push_back({0, ci.insert(c)});
            break;
        }
        case R_TYPO: {  // typo/mod.rs:42-77
            uint32_t base = to.term_ids_len() == 1 ? 0 : to.term_ids_len();
            uint8_t mx = max_typo_cost(ctx, to.term_subset);
            for (uint8_t n = 0; n <= mx; n++) {
                Condition c;
                c.rule = rule;
                c.term = to;
                c.nbr_typos = n;
                if (n == 0) {
                    c.term.term_subset.one = NTypoSubset{};
                    c.term.term_subset.two = NTypoSubset{};
                } else if (n == 1) {
                    c.term.term_subset.zero = NTypoSubset{};
                    c.term.term_subset.two = NTypoSubset{};
                } else {
                    c.term.term_subset.zero = NTypoSubset{};
                    c.term.term_subset.one = NTypoSubset{};
                }
                edges.push_back({n + base, ci.insert(c)});
            }
            break;
        }
        case R_PROXIMITY: {  // proximity/build.rs:10-56
            uint32_t right_ngram_max = to.term_ids_len() - 1;
            Condition tc;
            tc.rule = rule;
            tc.term = to;
            if (!from || (uint16_t)(from->pos_end + 1) != to.pos_start) {
                edges.push_back({right_ngram_max, ci.insert(tc)});
                break;
            }
            for (uint32_t cost = right_ngram_max; cost < 3 + right_ngram_max; cost++) {  // MAX_DISTANCE - 1 = 3
                Condition c;
                c.rule = rule;
                c.term = to;
                c.prox_uninit = true;
                c.left_term = *from;
                c.cost = (uint8_t)(cost + 1);
                edges.push_back({cost, ci.insert(c)});
            }
            edges.push_back({3 + right_ngram_max, ci.insert(tc)});
            break;
        }
        case R_FID: {  // fid/mod.rs:49-121 (FxHashSet order unspecified: ascending fid here)
            std::set<uint16_t> all_fields;
            for (auto w : all_single_words_except_prefix_db(ctx, to.term_subset))
                for (auto f : ctx.word_fids(w.id)) all_fields.insert(f);
            for (auto p : all_phrases(ctx, to.term_subset))
                for (auto w : ctx.phrases[p].words)
                    if (w >= 0)
                        for (auto f : ctx.word_fids((uint32_t)w)) all_fields.insert(f);
            Word pw;
            if (use_prefix_db(ctx, to.term_subset, pw))
                for (auto f : ctx.word_prefix_fids(pw.id)) all_fields.insert(f);
            uint16_t current_max_weight = 0;
            for (auto fid : all_fields) {
                if (fid >= ctx.index.settings.weights.size()) continue;
                uint16_t weight = ctx.index.settings.weights[fid];
                if (weight > current_max_weight) current_max_weight = weight;
                Condition c;
                c.rule = rule;
                c.term = to;
                c.has_fid = true;
                c.fid = fid;
                edges.push_back({(uint32_t)weight * to.term_ids_len(), ci.insert(c)});
            }
            uint16_t max_weight = ctx.index.max_searchable_weight();
            if (current_max_weight < max_weight) {
                Condition c;
                c.rule = rule;
                c.term = to;
                c.has_fid = false;
                edges.push_back({(uint32_t)max_weight * to.term_ids_len(), ci.insert(c)});
            }
            break;
        }
        case R_POSITION: {  // position/mod.rs:50-126 (FxHashMap order unspecified: ascending cost here)
            std::set<uint16_t> all_positions;
            for (auto w : all_single_words_except_prefix_db(ctx, to.term_subset))
                for (auto p : ctx.word_positions(w.id)) all_positions.insert(p);
            for (auto ph : all_phrases(ctx, to.term_subset)) {
                for (auto w : ctx.phrases[ph].words)
                    if (w >= 0) {
                        for (auto p : ctx.word_positions((uint32_t)w)) all_positions.insert(p);
                        break;
                    }
            }
            Word pw;
            if (use_prefix_db(ctx, to.term_subset, pw))
                for (auto p : ctx.word_prefix_positions(pw.id)) all_positions.insert(p);
            std::map<uint32_t, std::vector<uint16_t>> positions_for_costs;
            for (auto position : all_positions) {
                uint32_t distance = position > to.pos_start ? position - to.pos_start : to.pos_start - position;
                uint32_t cost = 0;
                for (uint32_t i = 0; i < to.term_ids_len(); i++) cost += position_cost_from_distance(distance + i);
                positions_for_costs[cost].push_back(position);
            }
            uint32_t max_cost = to.term_ids_len() * 10;
            bool max_cost_exists = positions_for_costs.count(max_cost) > 0;
            for (auto &kv : positions_for_costs) {
                Condition c;
                c.rule = rule;
                c.term = to;
                c.positions = kv.second;
                edges.push_back({kv.first, ci.insert(c)});
            }
            if (!max_cost_exists) {
                Condition c;
                c.rule = rule;
                c.term = to;
                edges.push_back({max_cost, ci.insert(c)});
            }
            break;
        }
        case R_EXACTNESS: {  // exactness/mod.rs:77-91
            Condition e;
            e.rule = rule;
            e.term = to;
            e.exact_in_attribute = true;
            Condition a;
            a.rule = rule;
            a.term = to;
            uint32_t ei = ci.insert(e), ai = ci.insert(a);
            edges.push_back({0, ei});
            edges.push_back({to.term_ids_len(), ai});
            break;
        }
    }
    (void)from;
    return edges;
}

// proximity/compute_docids.rs:15-108
inline ComputedCondition proximity_compute_docids(Ctx &ctx, const Condition &c, const Bitmap &universe) {
    ComputedCondition out;
    out.universe_len = universe.len();
    if (!c.prox_uninit) {
        out.docids = compute_query_term_subset_docids(ctx, &universe, c.term.term_subset);
        out.end_term_subset = c.term;
        return out;
    }
    const LocatedQueryTermSubset &left_term = c.left_term, &right_term = c.term;
    uint8_t right_len = (uint8_t)right_term.term_ids_len();
    uint8_t forward_proximity = (uint8_t)(1 + c.cost - right_len);
    uint8_t backward_proximity = (uint8_t)(c.cost - right_len);
    Bitmap docids;
    // last_words_of_term_derivations :213-231
    std::set<std::pair<int32_t, Word>> lefts;
    for (auto w : all_single_words_except_prefix_db(ctx, left_term.term_subset)) lefts.insert({-1, w});
    for (auto p : all_phrases(ctx, left_term.term_subset)) {
        int32_t last = ctx.phrases[p].words.back();
        if (last >= 0) lefts.insert({(int32_t)p, Word{W_ORIGINAL, (uint32_t)last}});
    }
    Word right_prefix;
    if (use_prefix_db(ctx, right_term.term_subset, right_prefix)) {
        for (auto &lw : lefts) {  // compute_prefix_edges :110-170
            Bitmap uni = universe;
            if (lw.first >= 0) {
                uni.and_with(get_phrase_docids(ctx, (uint32_t)lw.first));
                if (uni.is_empty()) continue;
            }
            Bitmap nd;
            ctx.word_prefix_pair_proximity_docids(lw.second.id, right_prefix.id, forward_proximity, nd);
            docids.or_with(nd);
            if (lw.first < 0) {
                Bitmap bd;
                if (ctx.word_pair_proximity_docids(&uni, right_prefix.id, lw.second.id, backward_proximity, bd)) docids.or_with(bd);
            }
        }
    }
    // first_word_of_term_iter :232-251
    std::set<std::pair<uint32_t, int32_t>> rights;
    for (auto w : all_single_words_except_prefix_db(ctx, right_term.term_subset)) rights.insert({w.id, -1});
    for (auto p : all_phrases(ctx, right_term.term_subset)) {
        int32_t first = ctx.phrases[p].words.front();
        if (first >= 0) rights.insert({(uint32_t)first, (int32_t)p});
    }
    for (auto &lw : lefts) {
        if (rights.size() > 1) {
            if (lw.first >= 0) {
                if (universe.is_disjoint(get_phrase_docids(ctx, (uint32_t)lw.first))) continue;
            } else {
                Bitmap lwd;
                if (ctx.word_docids(&universe, lw.second, lwd) && lwd.is_empty()) continue;
            }
        }
        for (auto &rw : rights) {  // compute_non_prefix_edges :172-211
            Bitmap uni = universe;
            bool dead = false;
            if (lw.first >= 0) {
                uni.and_with(get_phrase_docids(ctx, (uint32_t)lw.first));
                if (uni.is_empty()) dead = true;
            }
            if (!dead && rw.second >= 0) {
                uni.and_with(get_phrase_docids(ctx, (uint32_t)rw.second));
                if (uni.is_empty()) dead = true;
            }
            if (dead) continue;
            Bitmap nd;
            if (ctx.word_pair_proximity_docids(&uni, lw.second.id, rw.first, forward_proximity, nd)) docids.or_with(nd);
            if (backward_proximity >= 1 && lw.first < 0 && rw.second < 0) {
                Bitmap bd;
                if (ctx.word_pair_proximity_docids(&uni, rw.first, lw.second.id, backward_proximity, bd)) docids.or_with(bd);
            }
        }
    }
    out.docids = std::move(docids);
    out.has_start = true;
    out.start_term_subset = left_term;
    out.end_term_subset = right_term;
    return out;
}

inline ComputedCondition resolve_condition(Ctx &ctx, const Condition &c, const Bitmap &universe) {
    ComputedCondition out;
    out.universe_len = universe.len();
    out.end_term_subset = c.term;
    switch (c.rule) {
        case R_WORDS:
        case R_TYPO: out.docids = compute_query_term_subset_docids(ctx, &universe, c.term.term_subset); break;
        case R_PROXIMITY: return proximity_compute_docids(ctx, c, universe);
        case R_FID:
            if (c.has_fid) out.docids = compute_query_term_subset_docids_within_field_id(ctx, &universe, c.term.term_subset, c.fid);
            break;
        case R_POSITION:
            for (auto p : c.positions)
                out.docids.or_with(compute_query_term_subset_docids_within_position(ctx, &universe, c.term.term_subset, p));
            break;
        case R_EXACTNESS:
            if (c.exact_in_attribute) {  // exactness/mod.rs:19-43,52-58
                keep_only_exact_term(ctx, out.end_term_subset.term_subset);
                out.end_term_subset.term_subset.mandatory = true;
                ExactTerm e = exact_term(ctx, c.term.term_subset);
                if (e.some) {
                    if (e.is_phrase)
                        out.docids = bm_and(get_phrase_docids(ctx, e.id), universe);
                    else {
                        Bitmap d;
                        if (ctx.word_docids(&universe, Word{W_ORIGINAL, e.id}, d)) out.docids = std::move(d);
                    }
                }
            } else
                out.docids = compute_query_term_subset_docids(ctx, &universe, c.term.term_subset);
            break;
    }
    return out;
}

// ---------------------------------------------------------------- graph build (ranking_rule_graph/build.rs:12-91)
inline RankingRuleGraph build_rr_graph(Ctx &ctx, int rule, const QueryGraph &qg,
                                       const std::vector<std::pair<bool, std::pair<uint32_t, Bits>>> &cost_of_ignoring) {
    RankingRuleGraph g;
    g.query_graph = qg;
    uint32_t n = (uint32_t)qg.nodes.size();
    CondInterner ci;
    std::vector<std::vector<uint32_t>> eon(n);
    auto insert_edge = [&](Edge e) -> uint32_t {
        for (uint32_t i = 0; i < g.edges_store.size(); i++) {
            const Edge &o = g.edges_store[i];
            if (o.source_node == e.source_node && o.dest_node == e.dest_node && o.cost == e.cost && o.condition == e.condition) return i;
        }
        g.edges_store.push_back(std::move(e));
        return (uint32_t)g.edges_store.size() - 1;
    };
    for (uint32_t src = 0; src < n; src++) {
        const QueryNode &sn = qg.nodes[src];
        for (auto dst : sn.successors.items()) {
            const LocatedQueryTermSubset *src_term = nullptr;
            if (sn.kind == NODE_TERM)
                src_term = &sn.term;
            else if (sn.kind != NODE_START)
                throw std::runtime_error("bad source node");
            const QueryNode &dn = qg.nodes[dst];
            if (dn.kind == NODE_END) {
                eon[src].push_back(insert_edge(Edge{true, src, dst, 0, -1, Bits(n)}));
                continue;
            }
            if (dn.kind != NODE_TERM) throw std::runtime_error("bad dest node");
            if (cost_of_ignoring[dst].first) {
                uint32_t c = cost_of_ignoring[dst].second.first * dn.term.term_ids_len();
                eon[src].push_back(insert_edge(Edge{true, src, dst, c, -1, cost_of_ignoring[dst].second.second}));
            }
            auto edges = build_edges(ctx, rule, ci, src_term, dn.term);
            for (auto &e : edges) eon[src].push_back(insert_edge(Edge{true, src, dst, e.first, (int32_t)e.second, Bits(n)}));
        }
    }
    g.conditions = ci.items;
    uint32_t ne = (uint32_t)g.edges_store.size();
    g.edges_of_node.assign(n, Bits(ne));
    for (uint32_t i = 0; i < n; i++)
        for (auto e : eon[i]) g.edges_of_node[i].insert(e);
    return g;
}

// cheapest_paths.rs:285-310 (memoised over successors; same fix-point as the backward BFS)
inline std::vector<std::vector<uint64_t>> find_all_costs_to_end(const RankingRuleGraph &g) {
    uint32_t n = (uint32_t)g.query_graph.nodes.size();
    std::vector<std::vector<uint64_t>> costs(n);
    std::vector<int> state(n, 0);
    std::function<void(uint32_t)> visit = [&](uint32_t node) {
        if (state[node]) return;
        state[node] = 1;
        if (node == g.query_graph.end_node) {
            costs[node] = {0};
            return;
        }
        std::vector<uint64_t> self;
        for (auto ei : g.edges_of_node[node].items()) {
            const Edge &e = g.edges_store[ei];
            visit(e.dest_node);
            for (auto c : costs[e.dest_node]) self.push_back(e.cost + c);
        }
        std::sort(self.begin(), self.end());
        self.erase(std::unique(self.begin(), self.end()), self.end());
        costs[node] = std::move(self);
    };
    for (uint32_t i = 0; i < n; i++) visit(i);
    return costs;
}

// dead_ends_cache.rs
struct DeadEndsCache {
    std::vector<uint32_t> conditions;
    std::vector<DeadEndsCache> next;
    Bits forbidden;
    explicit DeadEndsCache(uint32_t n = 0) : forbidden(n) {}
    DeadEndsCache *advance(uint32_t c) {
        for (size_t i = 0; i < conditions.size(); i++)
            if (conditions[i] == c) return &next[i];
        return nullptr;
    }
    Bits forbidden_for_all_prefixes_up_to(const std::vector<uint32_t> &prefix) {
        Bits f = forbidden;
        DeadEndsCache *cur = this;
        for (auto c : prefix) {
            DeadEndsCache *nx = cur->advance(c);
            if (!nx) break;
            cur = nx;
            f.union_with(cur->forbidden);
        }
        return f;
    }
    bool forbidden_after_prefix(const std::vector<uint32_t> &prefix, Bits &out) {
        DeadEndsCache *cur = this;
        for (auto c : prefix) {
            DeadEndsCache *nx = cur->advance(c);
            if (!nx) return false;
            cur = nx;
        }
        out = cur->forbidden;
        return true;
    }
    void forbid_condition_after_prefix(const uint32_t *prefix, size_t n, uint32_t c) {
        if (n == 0) {
            forbidden.insert(c);
            return;
        }
        DeadEndsCache *nx = advance(prefix[0]);
        if (nx) {
            nx->forbid_condition_after_prefix(prefix + 1, n - 1, c);
            return;
        }
        DeadEndsCache rest(forbidden.n);
        rest.forbid_condition_after_prefix(prefix + 1, n - 1, c);
        conditions.push_back(prefix[0]);
        next.push_back(std::move(rest));
    }
};

// ---------------------------------------------------------------- ranking rule interface (ranking_rules.rs:26-95)
struct RuleOutput {
    QueryGraph query;
    Bitmap candidates;
    Score score;
};
struct RankingRule {
    virtual ~RankingRule() {}
    virtual void start_iteration(Ctx &ctx, const Bitmap &universe, const QueryGraph &query) = 0;
    virtual bool next_bucket(Ctx &ctx, const Bitmap &universe, RuleOutput &out) = 0;
    // ranking_rules.rs:55-74: a bucket if it can be had without blocking; the default says it cannot (Poll::Pending)
    virtual bool non_blocking_next_bucket(Ctx &, const Bitmap &, RuleOutput &) { return false; }
    virtual void end_iteration() = 0;
};

// ---------------------------------------------------------------- GraphBasedRankingRule (graph_based_ranking_rule.rs)
struct GraphRule : RankingRule {
    int rule;
    bool has_tms;
    int tms;
    // state
    bool active = false;
    RankingRuleGraph graph;
    std::map<uint32_t, ComputedCondition> conditions_cache;
    DeadEndsCache dead_ends;
    std::vector<std::vector<uint64_t>> all_costs;
    uint64_t cur_cost = 0, next_max_cost = 0;

    GraphRule(int rule_, bool has_tms_, int tms_) : rule(rule_), has_tms(has_tms_), tms(tms_) {}

    int score_kind() const {
        switch (rule) {
            case R_WORDS: return S_WORDS;
            case R_TYPO: return S_TYPO;
            case R_PROXIMITY: return S_PROXIMITY;
            case R_FID: return S_FID;
            case R_POSITION: return S_POSITION;
            default: return S_EXACT_WORDS;
        }
    }

    void start_iteration(Ctx &ctx, const Bitmap &, const QueryGraph &qg) override {
        uint32_t n = (uint32_t)qg.nodes.size();
        uint64_t nmc = 1;
        std::vector<std::pair<bool, std::pair<uint32_t, Bits>>> removal_cost(n, {false, {0, Bits(n)}});
        if (has_tms) {
            size_t wip = words_in_phrases_count(ctx, qg);
            nmc += wip > 0 ? wip - 1 : 0;
            std::vector<Bits> order;
            if (tms == TMS_LAST)
                order = removal_order_last(ctx, qg);
            else if (tms == TMS_FREQUENCY)
                order = removal_order_frequency(ctx, qg);
            if (tms != TMS_ALL) {
                Bits forbidden(n);
                for (auto &ns : order) {
                    for (auto nd : ns.items()) removal_cost[nd] = {true, {1, forbidden}};
                    forbidden.union_with(ns);
                }
            }
        }
        graph = build_rr_graph(ctx, rule, qg, removal_cost);
        conditions_cache.clear();
        dead_ends = DeadEndsCache((uint32_t)graph.conditions.size());
        all_costs = find_all_costs_to_end(graph);
        uint64_t mx = 0;
        for (auto c : all_costs[graph.query_graph.root_node]) mx = std::max(mx, c);
        nmc += mx;
        cur_cost = 0;
        next_max_cost = nmc;
        active = true;
    }
    void end_iteration() override { active = false; }

    // condition_docids_cache.rs:34-57
    ComputedCondition &get_computed_condition(Ctx &ctx, uint32_t ci, const Bitmap &universe) {
        auto it = conditions_cache.find(ci);
        if (it != conditions_cache.end()) {
            if (it->second.universe_len != universe.len()) {
                it->second.docids.and_with(universe);
                it->second.universe_len = universe.len();
            }
            return it->second;
        }
        ComputedCondition cc = resolve_condition(ctx, graph.conditions[ci], universe);
        return conditions_cache.emplace(ci, std::move(cc)).first->second;
    }

    std::set<uint32_t> remove_edges_with_condition(uint32_t cond) {
        std::set<uint32_t> sources;
        for (uint32_t i = 0; i < graph.edges_store.size(); i++) {
            Edge &e = graph.edges_store[i];
            if (!e.alive || e.condition < 0) continue;
            if ((uint32_t)e.condition == cond) {
                e.alive = false;
                graph.edges_of_node[e.source_node].remove(i);
                sources.insert(e.source_node);
            }
        }
        return sources;
    }

    // visit_path_condition :383-437
    bool visit_path_condition(Ctx &ctx, const Bitmap &universe, std::vector<std::pair<uint32_t, Bitmap>> &subpath,
                              std::set<uint32_t> &removed_sources, uint32_t latest) {
        const Bitmap &cd = get_computed_condition(ctx, latest, universe).docids;
        if (cd.is_empty()) {
            dead_ends.forbidden.insert(latest);
            auto src = remove_edges_with_condition(latest);
            removed_sources.insert(src.begin(), src.end());
            return false;
        }
        Bitmap latest_docids = subpath.empty() ? cd : bm_and(subpath.back().second, cd);
        if (!latest_docids.is_empty()) {
            subpath.push_back({latest, std::move(latest_docids)});
            return true;
        }
        std::vector<uint32_t> pre;
        for (auto &s : subpath) pre.push_back(s.first);
        dead_ends.forbid_condition_after_prefix(pre.data(), pre.size(), latest);
        if (subpath.size() <= 1) return false;
        std::vector<uint32_t> subprefix;
        for (size_t i = 0; i + 1 < subpath.size(); i++) {
            subprefix.push_back(subpath[i].first);
            if (cd.is_disjoint(subpath[i].second))
                dead_ends.forbid_condition_after_prefix(subprefix.data(), subprefix.size(), latest);
        }
        return false;
    }

    // PathVisitor (cheapest_paths.rs:94-282)
    struct Visitor {
        GraphRule &r;
        uint64_t remaining;
        std::vector<uint32_t> path;
        Bits visited_conditions, visited_nodes, forbidden_conditions, nodes_to_skip;
        std::function<int(const std::vector<uint32_t> &)> visit;  // 0 continue, 1 break
        // returns -1 break, 0 none valid, 1 some valid
        int visit_node(uint32_t from) {
            bool any_valid = false;
            std::vector<uint32_t> edges = r.graph.edges_of_node[from].items();
            for (auto ei : edges) {
                Edge e = r.graph.edges_store[ei];
                if (!e.alive) continue;
                if (remaining < e.cost) continue;
                remaining -= e.cost;
                int cf = e.condition >= 0 ? visit_condition((uint32_t)e.condition, e.dest_node, e.nodes_to_skip)
                                          : visit_no_condition(e.dest_node, e.nodes_to_skip);
                remaining += e.cost;
                if (cf < 0) return -1;
                if (cf > 0) {
                    any_valid = true;
                    forbidden_conditions = r.dead_ends.forbidden_for_all_prefixes_up_to(path);
                    if (visited_conditions.intersects(forbidden_conditions)) return 1;
                }
            }
            return any_valid ? 1 : 0;
        }
        bool cost_reachable(uint32_t node) {
            for (auto c : r.all_costs[node])
                if (c == remaining) return true;
            return false;
        }
        int visit_no_condition(uint32_t dest, const Bits &edge_skip) {
            if (!cost_reachable(dest)) return 0;
            if (dest == r.graph.query_graph.end_node) return visit(path) ? -1 : 1;
            Bits old = nodes_to_skip;
            nodes_to_skip.union_with(edge_skip);
            int cf = visit_node(dest);
            nodes_to_skip = old;
            return cf;
        }
        int visit_condition(uint32_t cond, uint32_t dest, const Bits &edge_skip) {
            if (forbidden_conditions.contains(cond) || nodes_to_skip.contains(dest) || edge_skip.intersects(visited_nodes)) return 0;
            if (!cost_reachable(dest)) return 0;
            path.push_back(cond);
            visited_nodes.insert(dest);
            visited_conditions.insert(cond);
            Bits old_forb = forbidden_conditions;
            Bits nf;
            if (r.dead_ends.forbidden_after_prefix(path, nf)) forbidden_conditions.union_with(nf);
            Bits old_skip = nodes_to_skip;
            nodes_to_skip.union_with(edge_skip);
            int cf = visit_node(dest);
            nodes_to_skip = old_skip;
            forbidden_conditions = old_forb;
            visited_conditions.remove(cond);
            visited_nodes.remove(dest);
            path.pop_back();
            return cf;
        }
    };

    bool next_bucket(Ctx &ctx, const Bitmap &universe_in, RuleOutput &out) override {
        const auto &root_costs = all_costs[graph.query_graph.root_node];
        bool found = false;
        uint64_t cost = 0;
        for (auto c : root_costs)
            if (c >= cur_cost) {
                cost = c;
                found = true;
                break;
            }
        if (!found) {
            active = false;
            return false;
        }
        cur_cost = cost + 1;
        Bitmap bucket;
        uint64_t rank = next_max_cost - cost;
        Bitmap universe = universe_in;
        std::vector<std::vector<uint32_t>> good_paths;
        std::vector<std::pair<uint32_t, Bitmap>> subpaths;
        std::set<uint32_t> removed_sources;
        uint32_t nc = (uint32_t)graph.conditions.size(), nn = (uint32_t)graph.query_graph.nodes.size();
        Visitor v{*this, cost, {}, Bits(nc), Bits(nn), Bits(nc), Bits(nn), nullptr};
        v.visit = [&](const std::vector<uint32_t> &path) -> int {
            if (universe.is_empty()) return 1;
            size_t idx = 0;
            while (idx < path.size() && idx < subpaths.size() && path[idx] == subpaths[idx].first) idx++;
            subpaths.resize(idx);
            for (size_t i = idx; i < path.size(); i++)
                if (!visit_path_condition(ctx, universe, subpaths, removed_sources, path[i])) return 0;
            Bitmap path_docids;
            if (subpaths.empty())
                path_docids = universe;
            else {
                path_docids = std::move(subpaths.back().second);
                subpaths.pop_back();
            }
            good_paths.push_back(path);
            bucket.or_with(path_docids);
            universe.sub(path_docids);
            for (auto &sp : subpaths) sp.second.sub(path_docids);
            return universe.is_empty() ? 1 : 0;
        };
        v.visit_node(graph.query_graph.root_node);

        std::vector<std::vector<std::pair<std::pair<bool, LocatedQueryTermSubset>, LocatedQueryTermSubset>>> paths;
        for (auto &gp : good_paths) {
            std::vector<std::pair<std::pair<bool, LocatedQueryTermSubset>, LocatedQueryTermSubset>> p;
            for (auto c : gp) {
                const ComputedCondition &cc = conditions_cache.at(c);
                p.push_back({{cc.has_start, cc.start_term_subset}, cc.end_term_subset});
            }
            paths.push_back(std::move(p));
        }
        out.query = build_from_paths(paths);
        if (!removed_sources.empty()) all_costs = find_all_costs_to_end(graph);
        out.candidates = std::move(bucket);
        out.score = Score{score_kind(), (uint32_t)rank, (uint32_t)next_max_cost, false, 0};
        return true;
    }
};

// ---------------------------------------------------------------- ExactAttribute (exact_attribute.rs)
inline uint16_t bucketed_position(uint16_t rel) {  // lib.rs:248-260
    if (rel < 16) return rel;
    if (rel < 24) return 24;
    uint32_t p = 1;
    while (p < rel) p <<= 1;
    return (uint16_t)p;
}
struct ExactAttributeRule : RankingRule {
    int state = 0;  // 0 uninit, 1 ExactAttribute, 2 AttributeStarts, 3 Empty
    QueryGraph qg;
    std::vector<std::pair<Bitmap, Bitmap>> per_attr;  // (start_with_exact, exact_word_count)

    void start_iteration(Ctx &ctx, const Bitmap &universe, const QueryGraph &query) override {
        qg = query;
        per_attr.clear();
        state = 3;
        struct Info {
            ExactTerm t;
            uint16_t start_position;
            uint8_t start_term_id;
            size_t position_count;
        };
        std::vector<Info> exact_terms;
        for (auto &n : query.nodes) {
            if (n.kind != NODE_TERM) continue;
            ExactTerm e = exact_term(ctx, n.term.term_subset);
            if (!e.some) continue;
            exact_terms.push_back({e, n.term.pos_start, n.term.tid_start, n.term.positions_len()});
        }
        std::stable_sort(exact_terms.begin(), exact_terms.end(), [](const Info &a, const Info &b) { return a.start_term_id < b.start_term_id; });
        {
            std::vector<Info> dd;
            for (auto &e : exact_terms)
                if (dd.empty() || dd.back().start_term_id != e.start_term_id) dd.push_back(e);
            exact_terms.swap(dd);
        }
        size_t count_all_positions = 0;
        for (auto &e : exact_terms) count_all_positions += e.position_count;
        if (exact_terms.empty() || exact_terms[0].start_term_id != 0) return;
        uint8_t previous = 0;
        for (auto &e : exact_terms) {
            if (e.start_term_id < previous || e.start_term_id - previous > 1) return;
            previous = e.start_term_id;
        }
        Bitmap candidates = universe;
        std::vector<std::pair<std::vector<int32_t>, uint16_t>> words_positions;
        for (auto &e : exact_terms) {
            std::vector<int32_t> ws;
            if (e.t.is_phrase)
                ws = ctx.phrases[e.t.id].words;
            else
                ws = {(int32_t)e.t.id};
            words_positions.push_back({ws, e.start_position});
        }
        for (auto &wp : words_positions) {
            if (candidates.is_empty()) return;
            for (size_t off = 0; off < wp.first.size(); off++) {
                if (wp.first[off] < 0) continue;
                uint16_t bp = bucketed_position((uint16_t)(wp.second + off));
                Bitmap d;
                ctx.word_position_docids(&universe, (uint32_t)wp.first[off], bp, d);
                candidates.and_with(d);
                if (candidates.is_empty()) return;
            }
        }
        if (candidates.is_empty()) return;
        for (uint16_t fid = 0; fid < ctx.index.settings.n_fields; fid++) {
            Bitmap inter;
            bool first = true;
            for (auto &wp : words_positions)
                for (auto w : wp.first) {
                    if (w < 0) continue;
                    Bitmap d;
                    ctx.word_fid_docids(&candidates, (uint32_t)w, fid, d);
                    if (first) {
                        inter = std::move(d);
                        first = false;
                    } else
                        inter.and_with(d);
                }
            if (!inter.is_empty()) {
                Bitmap cnt;
                if (count_all_positions < 255) {
                    std::string k;
                    k.push_back((char)(fid >> 8));
                    k.push_back((char)(fid & 0xff));
                    k.push_back((char)count_all_positions);
                    ctx.get_value(DB_FIELD_ID_WORD_COUNT_DOCIDS, k, &universe, cnt);
                }
                per_attr.push_back({std::move(inter), std::move(cnt)});
            }
        }
        state = 1;
    }
    bool next_bucket(Ctx &, const Bitmap &universe, RuleOutput &out) override {
        out.query = qg;
        switch (state) {
            case 0: return false;
            case 1: {
                Bitmap c;
                for (auto &a : per_attr) c.or_with(bm_and(a.first, a.second));
                c.and_with(universe);
                out.candidates = std::move(c);
                out.score = Score{S_EXACT_ATTRIBUTE, 3, 3, false, 0};
                state = 2;
                return true;
            }
            case 2: {
                Bitmap c;
                for (auto &a : per_attr) {
                    Bitmap s = a.first;
                    s.sub(a.second);
                    c.or_with(s);
                }
                c.and_with(universe);
                out.candidates = std::move(c);
                out.score = Score{S_EXACT_ATTRIBUTE, 2, 3, false, 0};
                state = 3;
                return true;
            }
            default:
                out.candidates = universe;
                out.score = Score{S_EXACT_ATTRIBUTE, 1, 3, false, 0};
                return true;
        }
    }
    void end_iteration() override { state = 0; }
};

// ---------------------------------------------------------------- vector store + VectorSort
// arroy/hannoy `Cosine` (arroy 0.6.4 / hannoy 0.1.3, not vendored): distance = (1 - cos)/2 in f32, cos clamped to [-1,1],
// 0 when either norm is ~0.  Exact scan (the reference is approximate).  Ordering contract:
// vector/store.rs:1059,1090 (ascending distance).
// 8-lane partial sums (GCC vector extension): the dot product of a query with a row, as arroy/hannoy compute it with SIMD
// (their summation order is not specified either; results are compared with a 1e-4 relative tolerance)
typedef float orc_v8 __attribute__((vector_size(32)));
inline float dot_f32(const float *a, const float *b, uint32_t d) {
    orc_v8 acc = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t i = 0;
    for (; i + 8 <= d; i += 8) {
        orc_v8 x, y;
        memcpy(&x, a + i, 32);
        memcpy(&y, b + i, 32);
        acc += x * y;
    }
    float s = ((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]));
    for (; i < d; i++) s += a[i] * b[i];
    return s;
}
inline float cosine_distance(float dot, float qn, float vn) {
    float pnqn = qn * vn;
    if (pnqn > 1.1920929e-7f) {
        float cs = dot / pnqn;
        cs = std::max(-1.0f, std::min(1.0f, cs));
        return (1.0f - cs) / 2.0f;
    }
    return 0.0f;
}
inline std::vector<std::pair<uint32_t, float>> nns_by_vector(const Index &ix, const float *q, size_t limit, const Bitmap *filter,
                                                             const float *cache_key = nullptr) {
    // served from the batch pass when the call covers every stored embedding (execute_hybrid / semantic search over documents_ids)
    if (!ix.nns_cache.empty()) {
        auto it = ix.nns_cache.find(cache_key ? cache_key : q);
        if (it != ix.nns_cache.end() && (limit <= it->second.size() || it->second.size() == ix.emb_docids.size()) &&
            (!filter || filter->len() >= ix.documents_ids.len())) {
            std::vector<std::pair<uint32_t, float>> r(it->second.begin(), it->second.begin() + std::min(limit, it->second.size()));
            return r;
        }
    }
    std::vector<std::pair<uint32_t, float>> res;
    uint32_t d = ix.dim;
    float qn = 0;
    for (uint32_t i = 0; i < d; i++) qn += q[i] * q[i];
    qn = std::sqrt(qn);
    size_t n = ix.emb_docids.size();
    std::vector<float> row(d);
    for (size_t r = 0; r < n; r++) {
        uint32_t doc = ix.emb_docids[r];
        if (filter && !filter->contains(doc)) continue;
        const float *v;
        if (ix.embeddings_f16.empty())
            v = ix.embeddings.data() + r * d;
        else {
            for (uint32_t i = 0; i < d; i++) row[i] = ix.emb_at(r, i);
            v = row.data();
        }
        float dot = 0;
        for (uint32_t i = 0; i < d; i++) dot += q[i] * v[i];
        res.push_back({doc, cosine_distance(dot, qn, ix.emb_norms[r])});
    }
    std::sort(res.begin(), res.end(), [](auto &a, auto &b) { return a.second != b.second ? a.second < b.second : a.first < b.first; });
    if (res.size() > limit) res.resize(limit);
    return res;
}
// One blocked pass over the store for a whole batch of query vectors (the shape of the scan is ours; every distance is the same
// arroy/hannoy `Cosine` as above): row blocks are spread over threads, every thread keeps the k best (distance, docid) of every
// query, the per-thread lists are merged.  Fills ix.nns_cache.
inline void batch_nns(const Index &ix, const float *queries, uint32_t nq, size_t k, unsigned n_threads) {
    ix.nns_cache.clear();
    const uint32_t d = ix.dim;
    const size_t n = ix.emb_docids.size();
    if (!nq || !n || !d) return;
    std::vector<float> qn(nq);
    for (uint32_t q = 0; q < nq; q++) {
        float s = 0;
        for (uint32_t i = 0; i < d; i++) s += queries[(size_t)q * d + i] * queries[(size_t)q * d + i];
        qn[q] = std::sqrt(s);
    }
    n_threads = std::max(1u, n_threads);
    typedef std::pair<float, uint32_t> DK;  // (distance, docid): the nns ordering
    std::vector<std::vector<std::vector<DK>>> best(n_threads, std::vector<std::vector<DK>>(nq));
    const size_t BLOCK = 256;
    std::atomic<size_t> next{0};
    auto worker = [&](unsigned t) {
        std::vector<float> blk(BLOCK * d);
        auto &mine = best[t];
        for (auto &h : mine) h.reserve(k + 1);
        for (;;) {
            const size_t r0 = next.fetch_add(BLOCK);
            if (r0 >= n) break;
            const size_t nr = std::min(BLOCK, n - r0);
            const float *rows;
            if (ix.embeddings_f16.empty())
                rows = ix.embeddings.data() + r0 * d;
            else {
                for (size_t r = 0; r < nr; r++)
                    for (uint32_t i = 0; i < d; i++) blk[r * d + i] = ix.emb_at(r0 + r, i);
                rows = blk.data();
            }
            for (uint32_t q = 0; q < nq; q++) {
                const float *qv = queries + (size_t)q * d;
                auto &h = mine[q];  // max-heap on (distance, docid) of the k best so far
                for (size_t r = 0; r < nr; r++) {
                    DK cand{cosine_distance(dot_f32(qv, rows + r * d, d), qn[q], ix.emb_norms[r0 + r]), ix.emb_docids[r0 + r]};
                    if (h.size() < k) {
                        h.push_back(cand);
                        std::push_heap(h.begin(), h.end());
                    } else if (cand < h.front()) {
                        std::pop_heap(h.begin(), h.end());
                        h.back() = cand;
                        std::push_heap(h.begin(), h.end());
                    }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n_threads; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    for (uint32_t q = 0; q < nq; q++) {
        std::vector<DK> all;
        for (unsigned t = 0; t < n_threads; t++) all.insert(all.end(), best[t][q].begin(), best[t][q].end());
        std::sort(all.begin(), all.end());
        if (all.size() > k) all.resize(k);
        auto &out = ix.nns_cache[queries + (size_t)q * d];
        for (auto &x : all) out.push_back({x.second, x.first});
    }
}
// vector/distribution.rs:103-130.  Rust never contracts a*b+c into a fused multiply-add; this file is built with -march=x86-64-v3, so
// contraction is switched off here (pinned by hybrid.rs:549-568: 0.19161224365234375, one f32 rounding away from the fused result).
__attribute__((optimize("fp-contract=off"))) inline float distribution_shift(float mean, float sigma, float score) {
    float factor = 0.4f / sigma;
    float offset = 0.5f - (factor * mean);
    float s = factor * score + offset;
    if (s <= 0.0f) s = 1.1920929e-7f;
    if (s > 1.0f) s = 1.0f;
    return s;
}
struct VectorSortRule : RankingRule {
    std::vector<float> target;
    const float *source = nullptr;  // the caller's vector (key of Index::nns_cache)
    Bitmap vector_candidates;
    size_t limit;
    std::vector<std::pair<uint32_t, float>> cached;
    size_t cursor = 0;
    VectorSortRule(std::vector<float> t, Bitmap c, size_t l) : target(std::move(t)), vector_candidates(std::move(c)), limit(l) {}
    size_t fill_buffer(Ctx &ctx, const Bitmap &cands) {
        cached = nns_by_vector(ctx.index, target.data(), limit, &cands, source);
        cursor = 0;
        return cached.size();
    }
    bool next_results(Ctx &ctx, const Bitmap &cands, Bitmap &out, float &score) {
        while (cursor < cached.size()) {
            float dist = cached[cursor].second;
            std::vector<uint32_t> grp;
            while (cursor < cached.size() && cached[cursor].second == dist) grp.push_back(cached[cursor++].first);
            std::sort(grp.begin(), grp.end());
            Bitmap c = Bitmap::from_sorted(grp);
            c.and_with(cands);
            if (!c.is_empty()) {
                score = 1.0f - dist;
                if (ctx.index.has_distribution) score = distribution_shift(ctx.index.dist_mean, ctx.index.dist_sigma, score);
                out = std::move(c);
                return true;
            }
        }
        return false;
    }
    void start_iteration(Ctx &ctx, const Bitmap &universe, const QueryGraph &) override {
        fill_buffer(ctx, bm_and(vector_candidates, universe));
    }
    bool next_bucket(Ctx &ctx, const Bitmap &universe, RuleOutput &out) override {
        for (;;) {
            Bitmap cands = bm_and(vector_candidates, universe);
            if (cands.is_empty()) {
                out.candidates = universe;
                out.score = Score{S_VECTOR, 0, 1, false, 0};
                return true;
            }
            Bitmap c;
            float score;
            if (next_results(ctx, cands, c, score)) {
                out.candidates = std::move(c);
                out.score = Score{S_VECTOR, 0, 1, true, score};
                return true;
            }
            if (fill_buffer(ctx, cands) == 0) {
                out.candidates = universe;
                out.score = Score{S_VECTOR, 0, 1, false, 0};
                return true;
            }
        }
    }
    bool non_blocking_next_bucket(Ctx &ctx, const Bitmap &universe, RuleOutput &out) override {  // vector_sort.rs:175-201
        Bitmap cands = bm_and(vector_candidates, universe);
        if (cands.is_empty()) {
            out.candidates = universe;
            out.score = Score{S_VECTOR, 0, 1, false, 0};
            return true;
        }
        Bitmap c;
        float score;
        if (next_results(ctx, cands, c, score)) {
            out.candidates = std::move(c);
            out.score = Score{S_VECTOR, 0, 1, true, score};
            return true;
        }
        return false;
    }
    void end_iteration() override {}
};

// ---------------------------------------------------------------- bucket_sort (bucket_sort.rs:23-343, 382-460)
struct BucketSortOutput {
    std::vector<uint32_t> docids;
    std::vector<std::vector<Score>> scores;
    Bitmap all_candidates;
    bool degraded = false;
};

// crates/milli/src/lib.rs:154-226: a time budget, or (the reference's test hook) "exceeded from the n-th poll on"
struct Deadline {
    bool has_time = false;
    std::chrono::steady_clock::time_point at;
    long stop_after = -1;
    mutable long polls = 0;
    bool exceeded() const {
        if (stop_after >= 0) return polls++ >= stop_after;  // a poll count ignores the clock entirely (lib.rs:207-216)
        return has_time && std::chrono::steady_clock::now() > at;
    }
};

inline BucketSortOutput bucket_sort(Ctx &ctx, std::vector<std::unique_ptr<RankingRule>> &rules, const QueryGraph &query,
                                    const Bitmap &universe, size_t from, size_t length, int scoring_strategy,
                                    bool has_threshold, double threshold, const Deadline &deadline = Deadline()) {
    BucketSortOutput out;
    if (universe.len() < from) {
        out.all_candidates = universe;
        return out;
    }
    if (rules.empty()) {
        size_t i = 0;
        universe.for_each([&](uint32_t d) {
            if (i >= from && out.docids.size() < length) out.docids.push_back(d);
            i++;
        });
        out.scores.assign(out.docids.size(), {});
        out.all_candidates = universe;
        return out;
    }
    size_t n = rules.size();
    rules[0]->start_iteration(ctx, universe, query);
    std::vector<Score> rr_scores;
    std::vector<Bitmap> universes(n);
    universes[0] = universe;
    size_t cur = 0;
    Bitmap all_candidates = universe;
    size_t cur_offset = 0;

    auto maybe_add = [&](const Bitmap &candidates) {
        all_candidates.or_with(candidates);
        if (candidates.is_empty()) return;
        size_t clen = candidates.len();
        if (cur_offset < from) {
            if (cur_offset + clen < from) {
                // skip
            } else {
                std::vector<uint32_t> v = candidates.to_vec();
                size_t skip = from - cur_offset;
                for (size_t i = skip; i < v.size() && out.docids.size() < length; i++) {
                    out.docids.push_back(v[i]);
                    out.scores.push_back(rr_scores);
                }
            }
        } else {
            std::vector<uint32_t> v = candidates.to_vec();
            for (size_t i = 0; i < v.size() && out.docids.size() < length; i++) {
                out.docids.push_back(v[i]);
                out.scores.push_back(rr_scores);
            }
        }
        cur_offset += clen;
    };

    while (out.docids.size() < length) {
        if (universes[cur].is_empty() || (scoring_strategy == SCORING_SKIP && universes[cur].len() == 1)) {
            Bitmap bucket = std::move(universes[cur]);
            universes[cur].clear();
            maybe_add(bucket);
            // back!()
            universes[cur].clear();
            rules[cur]->end_iteration();
            if (cur == 0) break;
            cur--;
            if (rr_scores.size() > cur) rr_scores.pop_back();
            continue;
        }
        RuleOutput nb;
        if (deadline.exceeded()) {
            // bucket_sort.rs:206-264: the graph rules cannot answer without blocking (ranking_rules.rs:67-74), so the rule's whole
            // remaining universe is returned as it is with a Skipped score, rule after rule up to the first one
            bool ready = false;
            for (;;) {
                if (rules[cur]->non_blocking_next_bucket(ctx, universes[cur], nb)) {
                    ready = true;
                    break;
                }
                Bitmap bucket = std::move(universes[cur]);
                universes[cur].clear();
                rr_scores.push_back(Score{S_SKIPPED, 0, 1, false, 0});
                bool below = has_threshold && global_score(rr_scores) < threshold;
                if (below) {
                    all_candidates.sub(bucket);
                } else
                    maybe_add(bucket);
                rr_scores.pop_back();
                if (cur == 0) {
                    out.all_candidates = std::move(all_candidates);
                    out.degraded = true;
                    return out;
                }
                universes[cur].clear();
                rules[cur]->end_iteration();
                cur--;
                if (rr_scores.size() > cur) rr_scores.pop_back();
            }
            (void)ready;
        } else if (!rules[cur]->next_bucket(ctx, universes[cur], nb)) {
            universes[cur].clear();
            rules[cur]->end_iteration();
            if (cur == 0) break;
            cur--;
            if (rr_scores.size() > cur) rr_scores.pop_back();
            continue;
        }
        rr_scores.push_back(nb.score);
        bool below = has_threshold && global_score(rr_scores) < threshold;
        universes[cur].sub(nb.candidates);
        if (cur == n - 1 || (scoring_strategy == SCORING_SKIP && nb.candidates.len() <= 1) ||
            cur_offset + nb.candidates.len() < from || below) {
            if (below) {
                all_candidates.sub(nb.candidates);
                all_candidates.sub(universes[cur]);
            } else
                maybe_add(nb.candidates);
            rr_scores.pop_back();
            continue;
        }
        cur++;
        universes[cur] = nb.candidates;
        rules[cur]->start_iteration(ctx, nb.candidates, nb.query);
    }
    out.all_candidates = std::move(all_candidates);
    return out;
}

}  // namespace orc
