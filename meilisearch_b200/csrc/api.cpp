// C ABI (include/b200milli.h) over the engine; hybrid merge (search/hybrid.rs) lives here because it is pure host logic
// over two result lists.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <new>

#include "engine.h"

using namespace b200;

struct b200_index {
    Engine e;
};
struct b200_rule {
    Engine::RuleRun *run;
    uint64_t n_words;
};
namespace b200 {
int rule_next_impl(Engine::RuleRun *run, uint64_t n_words64, const uint64_t *universe, uint64_t *out_bitmap, uint64_t n_words, uint32_t *rank,
                   uint32_t *max_rank, GraphObj **out_query);
void rule_end_impl(Engine::RuleRun *run);
}  // namespace b200

static thread_local std::string g_open_error;

// no exception crosses the C boundary: every entry point runs its body through this
template <class F>
static int guarded(b200_index *h, F body) {
    if (!h) return B200_ERR_INVALID;
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return h->e.fail(B200_ERR_CAPACITY, "out of host memory");
    } catch (const std::exception &ex) {
        return h->e.fail(B200_ERR_INVALID, ex.what());
    } catch (...) {
        return h->e.fail(B200_ERR_INVALID, "unexpected exception");
    }
}

extern "C" {

const char *b200_open_error(void) { return g_open_error.c_str(); }

int b200_open(int device_ordinal, b200_index **out) {
    *out = nullptr;
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        g_open_error = std::string("no CUDA device: ") + (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                       " (b200milli has no CPU fallback)";
        return B200_ERR_NO_DEVICE;
    }
    if (device_ordinal < 0 || device_ordinal >= n) {
        g_open_error = "device ordinal out of range";
        return B200_ERR_INVALID;
    }
    b200_index *h = new (std::nothrow) b200_index();
    if (!h) return B200_ERR_INVALID;
    h->e.device = device_ordinal;
    if ((err = cudaSetDevice(device_ordinal)) != cudaSuccess || (err = cudaStreamCreateWithFlags(&h->e.stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (err = cudaStreamCreateWithFlags(&h->e.vt.stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (err = cudaEventCreate(&h->e.ev0)) != cudaSuccess || (err = cudaEventCreate(&h->e.ev1)) != cudaSuccess) {
        g_open_error = std::string("CUDA init failed: ") + cudaGetErrorString(err);
        delete h;
        return B200_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) == cudaSuccess) h->e.sm_count = prop.multiProcessorCount;
    h->e.affinity.detect(device_ordinal);
    *out = h;
    return B200_OK;
}
void b200_close(b200_index *h) { delete h; }
const char *b200_last_error(const b200_index *h) { return h ? h->e.last_error.c_str() : "null handle"; }

int b200_stage_dictionary(b200_index *h, const uint8_t *bytes, const uint64_t *offsets, uint64_t n) {
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> g(h->e.mu);
        if (h->e.staged) return h->e.fail(B200_ERR_STATE, "staging after b200_stage_finish: open a new handle");
        if (!offsets || (!bytes && n && offsets[n])) return h->e.fail(B200_ERR_INVALID, "stage_dictionary: null argument");
        for (uint64_t i = 0; i < n; i++)
            if (offsets[i + 1] < offsets[i]) return h->e.fail(B200_ERR_INVALID, "stage_dictionary: offsets must not decrease");
        h->e.raw_dict_off.assign(offsets, offsets + n + 1);
        h->e.raw_dict_bytes.assign(bytes, bytes + offsets[n]);
        return B200_OK;
    });
}
int b200_stage_db(b200_index *h, int db, uint64_t n, const uint8_t *kb, const uint64_t *ko, const uint8_t *vb, const uint64_t *vo) {
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> g(h->e.mu);
        if (h->e.staged) return h->e.fail(B200_ERR_STATE, "staging after b200_stage_finish: open a new handle");
        if (db < 0 || db >= B200_DB_COUNT) return h->e.fail(B200_ERR_INVALID, "unknown database id");
        if (!ko || !vo || (n && (!kb || !vb))) return h->e.fail(B200_ERR_INVALID, "stage_db: null argument");
        for (uint64_t i = 0; i < n; i++)
            if (ko[i + 1] < ko[i] || vo[i + 1] < vo[i]) return h->e.fail(B200_ERR_INVALID, "stage_db: offsets must not decrease");
        RawDb &d = h->e.raw_dbs[db];
        d.n = n;
        d.koff.assign(ko, ko + n + 1);
        d.voff.assign(vo, vo + n + 1);
        d.keys.assign(kb, kb + ko[n]);
        d.vals.assign(vb, vb + vo[n]);
        return B200_OK;
    });
}
int b200_stage_documents_ids(b200_index *h, const uint8_t *cbo, uint64_t len) {
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> g(h->e.mu);
        if (h->e.staged) return h->e.fail(B200_ERR_STATE, "staging after b200_stage_finish: open a new handle");
        if (!cbo && len) return h->e.fail(B200_ERR_INVALID, "stage_documents_ids: null argument");
        h->e.raw_docids.assign(cbo, cbo + len);
        return B200_OK;
    });
}
int b200_stage_settings(b200_index *h, const b200_settings *s) {
    if (!h || !s) return B200_ERR_INVALID;
    std::lock_guard<std::mutex> g(h->e.mu);
    Settings &t = h->e.hix.settings;
    if (s->n_fields == 0 || s->n_fields > 1024) return h->e.fail(B200_ERR_INVALID, "n_fields out of range");
    if (!s->weights || (s->n_criteria && !s->criteria)) return h->e.fail(B200_ERR_INVALID, "stage_settings: null weights / criteria");
    t.n_fields = s->n_fields;
    t.weights.assign(s->weights, s->weights + s->n_fields);
    t.criteria.assign(s->criteria, s->criteria + s->n_criteria);
    t.authorize_typos = s->authorize_typos != 0;
    t.one_typo = s->min_word_len_one_typo;
    t.two_typos = s->min_word_len_two_typos;
    t.prefix_search = s->prefix_search != 0;
    t.exact_words.clear();
    if (s->exact_words) {
        std::string cur;
        for (const char *p = s->exact_words;; p++) {
            if (*p == '\n' || *p == 0) {
                if (!cur.empty()) t.exact_words[cur] = 1;
                cur.clear();
                if (!*p) break;
            } else
                cur.push_back(*p);
        }
    }
    return B200_OK;
}
static std::vector<std::string> split_ws(const char *s) {
    std::vector<std::string> out;
    std::string cur;
    for (const char *p = s;; p++) {
        if (*p == ' ' || *p == 0) {
            if (!cur.empty()) out.push_back(cur);
            cur.clear();
            if (!*p) break;
        } else
            cur.push_back(*p);
    }
    return out;
}
int b200_stage_synonyms(b200_index *h, uint32_t n, const char *const *from_words, const char *const *to_words) {
    std::lock_guard<std::mutex> g(h->e.mu);
    auto &syn = h->e.hix.settings.synonyms;
    syn.clear();
    for (uint32_t i = 0; i < n; i++) syn[split_ws(from_words[i])].push_back(split_ws(to_words[i]));
    return B200_OK;
}
int b200_stage_finish(b200_index *h) {
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> g(h->e.mu);
        if (h->e.staged) return h->e.fail(B200_ERR_STATE, "b200_stage_finish was already called on this handle");
        Settings keep = h->e.hix.settings;
        int rc = h->e.stage_finish();
        h->e.hix.settings = keep;
        return rc;
    });
}
int b200_stage_embeddings(b200_index *h, const float *v, uint64_t n, uint32_t d, const uint32_t *docids) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.stage_embeddings(v, nullptr, n, d, docids);
}
int b200_stage_embeddings_f16(b200_index *h, const uint16_t *rows, uint64_t n, uint32_t d, const uint32_t *docids) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.stage_embeddings(nullptr, rows, n, d, docids);
}
int b200_stage_distribution(b200_index *h, int enabled, float mean, float sigma) {
    std::lock_guard<std::mutex> g(h->e.mu);
    h->e.has_distribution = enabled != 0 && sigma > 0.f;
    h->e.dist_mean = mean;
    h->e.dist_sigma = sigma;
    return B200_OK;
}
int b200_derive_batch(b200_index *h, uint32_t n, const char *words, const uint32_t *off, const uint8_t *mt, const uint8_t *ip, uint32_t *one,
                      uint32_t *n_one, uint32_t *two, uint32_t *n_two) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.derive_batch(n, words, off, mt, ip, one, n_one, two, n_two);
}
int b200_nns_batch(b200_index *h, const float *q, uint32_t n_q, uint32_t d, uint32_t limit, const uint64_t *cand, uint64_t ncw, uint32_t *ids,
                   float *dist, uint32_t *n_out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    int rc = h->e.nns_batch(q, n_q, d, limit, cand, ncw, ids, dist, n_out);
    h->e.fold_vector_stats();
    return rc;
}
int b200_comm_unique_id(b200_index *h, uint8_t *out128) {
    std::lock_guard<std::mutex> g(h->e.mu);
    int rc = h->e.comm_load();
    if (rc != B200_OK) return rc;
    int e = h->e.sc.get_unique_id(out128);
    return e == 0 ? B200_OK : h->e.fail(B200_ERR_CUDA, "ncclGetUniqueId failed");
}
int b200_comm_init(b200_index *h, int rank, int world, const uint8_t *unique_id128) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.comm_init(rank, world, unique_id128);
}
int b200_nns_batch_sharded(b200_index *h, const float *q, uint32_t n_q, uint32_t d, uint32_t limit, const uint64_t *cand, uint64_t ncw, uint32_t *ids,
                           float *dist, uint32_t *n_out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    int rc = h->e.nns_batch(q, n_q, d, limit, cand, ncw, ids, dist, n_out, true);
    h->e.fold_vector_stats();
    return rc;
}
int b200_union_postings(b200_index *h, int db, const uint32_t *key_index, uint32_t n_keys, const uint64_t *universe, uint64_t n_universe_words,
                        uint64_t *out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.union_postings(db, key_index, n_keys, universe, n_universe_words, out);
}
int b200_proximity_pairs(b200_index *h, const uint32_t *left, uint32_t n_left, const uint32_t *right, uint32_t n_right, uint32_t fwd_prox,
                         uint32_t bwd_prox, const uint64_t *universe, uint64_t n_universe_words, uint64_t *out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    return h->e.proximity_pairs(left, n_left, right, n_right, fwd_prox, bwd_prox, universe, n_universe_words, out);
}
int b200_graph_from_tokens(b200_index *h, const b200_query_batch *one_query, b200_graph **out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    if (!h->e.staged) return h->e.fail(B200_ERR_STATE, "graph_from_tokens before b200_stage_finish");
    GraphObj *go = nullptr;
    int rc = h->e.graph_from_tokens(one_query, &go);
    *out = reinterpret_cast<b200_graph *>(go);
    return rc;
}
void b200_graph_free(b200_graph *g) { free_graph(reinterpret_cast<GraphObj *>(g)); }
int b200_rule_start(b200_index *h, int rule_kind, int terms_matching_strategy, const b200_graph *query, const uint64_t *universe,
                    uint64_t n_universe_words, b200_rule **out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    *out = nullptr;
    if (!h->e.staged) return h->e.fail(B200_ERR_STATE, "rule_start before b200_stage_finish");
    Engine::RuleRun *run = nullptr;
    int rc = h->e.rule_start(rule_kind, terms_matching_strategy, reinterpret_cast<const GraphObj *>(query), universe, n_universe_words, &run);
    if (rc != B200_OK) return rc;
    b200_rule *r = new b200_rule{run, h->e.hix.n_words64};
    *out = r;
    return B200_OK;
}
int b200_rule_next(b200_rule *r, const uint64_t *universe, uint64_t *out_bitmap, uint64_t n_words, uint32_t *rank, uint32_t *max_rank,
                   b200_graph **out_query) {
    if (!r) return B200_ERR_INVALID;
    GraphObj *child = nullptr;
    int rc = rule_next_impl(r->run, r->n_words, universe, out_bitmap, n_words, rank, max_rank, out_query ? &child : nullptr);
    if (out_query) *out_query = reinterpret_cast<b200_graph *>(child);
    return rc;
}
void b200_rule_end(b200_rule *r) {
    if (!r) return;
    rule_end_impl(r->run);
    delete r;
}
int b200_search_batch(b200_index *h, const b200_query_batch *b, b200_results *r) {
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> g(h->e.mu);
        if (!h->e.staged) return h->e.fail(B200_ERR_STATE, "search before b200_stage_finish");
        if (!b || !r || !r->docids || !r->n_hits) return h->e.fail(B200_ERR_INVALID, "search: null batch / results / docids / n_hits");
        if (b->n_queries && (!b->token_begin || !b->lemma_off)) return h->e.fail(B200_ERR_INVALID, "search: null token arrays");
        // every query starts with a definite status and no hits, whatever happens later
        for (uint32_t i = 0; i < b->n_queries; i++) {
            r->n_hits[i] = 0;
            if (r->status) r->status[i] = 0;
        }
        return h->e.search_batch(b, r);
    });
}
int b200_get_stats(b200_index *h, b200_stats *out) {
    std::lock_guard<std::mutex> g(h->e.mu);
    *out = h->e.stats;
    return B200_OK;
}
int b200_reset_stats(b200_index *h) {
    std::lock_guard<std::mutex> g(h->e.mu);
    uint64_t staged = h->e.stats.hbm_bytes_staged;
    h->e.stats = b200_stats{};
    h->e.stats.hbm_bytes_staged = staged;
    return B200_OK;
}
}

// ================================================================================= semantic + hybrid
namespace b200 {

static float distribution_shift(float mean, float sigma, float score) {  // vector/distribution.rs:103-130
    float factor = 0.4f / sigma, offset = 0.5f - factor * mean;
    float s = factor * score + offset;
    if (s <= 0.f) s = 1.1920929e-7f;
    if (s > 1.f) s = 1.f;
    return s;
}

// execute_vector_search (search/new/mod.rs:744-808): one VectorSort rule over documents_ids; buckets = runs of equal
// distance in ascending docid order (vector_sort.rs:80-95), which the (distance, docid) ordering of nns_batch reproduces.
int Engine::semantic_batch(const b200_query_batch *b, b200_results *r, uint32_t offset, uint32_t limit) {
    if (!b->vectors) return fail(B200_ERR_INVALID, "semantic search without vectors");
    uint32_t k = offset + limit;
    if (k == 0) k = 1;
    std::vector<uint32_t> ids((size_t)b->n_queries * k), n(b->n_queries);
    std::vector<float> dist((size_t)b->n_queries * k);
    std::vector<uint64_t> n_cand(b->n_queries, hix.n_documents);
    if (!b->universes) {
        int rc = nns_batch(b->vectors, b->n_queries, emb_d_user, k, nullptr, 0, ids.data(), dist.data(), n.data());
        if (rc != B200_OK) return rc;
    } else {
        // filtered_universe restricts the vector candidates (vector_sort.rs:58-78: `vector_candidates & universe`): the queries
        // are grouped by bitmap and every group is one scan with that candidate filter
        if (b->n_universe_words < hix.n_words64) return fail(B200_ERR_INVALID, "universe bitmaps shorter than the document range");
        std::map<const uint64_t *, std::vector<uint32_t>> groups;
        for (uint32_t q = 0; q < b->n_queries; q++) groups[b->universes[q]].push_back(q);
        const uint32_t d = emb_d_user;
        for (auto &g : groups) {
            const uint32_t m = (uint32_t)g.second.size();
            std::vector<float> vq((size_t)m * d);
            for (uint32_t i = 0; i < m; i++) memcpy(vq.data() + (size_t)i * d, b->vectors + (size_t)g.second[i] * d, (size_t)d * 4);
            std::vector<uint32_t> gi((size_t)m * k), gn(m);
            std::vector<float> gd((size_t)m * k);
            uint64_t cnt = hix.n_documents;
            if (g.first) {
                cnt = 0;
                for (uint64_t w = 0; w < hix.n_words64; w++) cnt += (uint64_t)__builtin_popcountll(g.first[w] & hix.base_ub[w]);
            }
            int rc = nns_batch(vq.data(), m, d, k, g.first, g.first ? hix.n_words64 : 0, gi.data(), gd.data(), gn.data());
            if (rc != B200_OK) return rc;
            for (uint32_t i = 0; i < m; i++) {
                const uint32_t q = g.second[i];
                n[q] = gn[i];
                n_cand[q] = cnt;
                memcpy(ids.data() + (size_t)q * k, gi.data() + (size_t)i * k, (size_t)gn[i] * 4);
                memcpy(dist.data() + (size_t)q * k, gd.data() + (size_t)i * k, (size_t)gn[i] * 4);
            }
        }
    }
    // VectorSort's last bucket (vector_sort.rs:128-160): once the embedded candidates are exhausted, the rest of the universe
    // follows in docid order with `similarity: None`
    for (uint32_t q = 0; q < b->n_queries; q++) {
        if (n[q] >= k) continue;
        const uint64_t *u = b->universes ? b->universes[q] : nullptr;
        for (uint64_t w = 0; w < hix.n_words64 && n[q] < k; w++) {
            uint64_t bits = hix.base_ub[w] & (u ? u[w] : ~0ull) & ~(w < emb_bitmap.size() ? emb_bitmap[w] : 0ull);
            while (bits && n[q] < k) {
                ids[(size_t)q * k + n[q]] = (uint32_t)(w * 64 + (uint64_t)__builtin_ctzll(bits));
                dist[(size_t)q * k + n[q]] = 2.0f;  // marks "no similarity"
                n[q]++;
                bits &= bits - 1;
            }
        }
    }
    for (uint32_t q = 0; q < b->n_queries; q++) {
        uint32_t hits = n[q] > offset ? std::min(limit, n[q] - offset) : 0;
        r->n_hits[q] = hits;
        if (r->status) r->status[q] = 0;
        if (r->degraded) r->degraded[q] = 0;
        if (r->used_negative_operator) r->used_negative_operator[q] = 0;
        if (r->n_candidates) r->n_candidates[q] = n_cand[q];
        for (uint32_t i = 0; i < hits; i++) {
            size_t src = (size_t)q * k + offset + i, at = (size_t)q * limit + i;
            r->docids[at] = ids[src];
            float sim = 1.0f - dist[src];
            if (dist[src] > 1.5f)
                sim = -1.f;  // Vector { similarity: None }
            else if (has_distribution)
                sim = distribution_shift(dist_mean, dist_sigma, sim);
            if (r->n_scores) {
                r->n_scores[at] = 1;
                r->score_kind[at * B200_MAX_SCORES] = B200_S_VECTOR;
                r->score_rank[at * B200_MAX_SCORES] = 0;
                r->score_max[at * B200_MAX_SCORES] = 1;
                r->score_sim[at * B200_MAX_SCORES] = sim;
            }
        }
    }
    return B200_OK;
}

namespace {
struct Hit {
    uint32_t doc;
    std::vector<double> values;  // ScoreDetails::score_values (score_details.rs:156-175)
    uint8_t n_scores;
    uint8_t kind[B200_MAX_SCORES];
    uint32_t rank[B200_MAX_SCORES], maxr[B200_MAX_SCORES];
    float sim[B200_MAX_SCORES];
};
void rank_merge(uint64_t &orank, uint64_t &omax, uint64_t irank, uint64_t imax) {
    orank = orank > 0 ? orank - 1 : 0;
    orank *= imax;
    omax *= imax;
    orank += irank;
}
void fill_values(Hit &h) {
    bool in_rank = false;
    uint64_t rk = 0, mx = 1;
    for (int s = 0; s < h.n_scores; s++) {
        if (h.kind[s] == B200_S_VECTOR) {
            if (in_rank) {
                h.values.push_back((double)rk / (double)mx);
                in_rank = false;
            }
            h.values.push_back(h.sim[s] < 0 ? 0.0 : (double)h.sim[s]);
        } else if (!in_rank) {
            rk = h.rank[s];
            mx = h.maxr[s];
            in_rank = true;
        } else
            rank_merge(rk, mx, h.rank[s], h.maxr[s]);
    }
    if (in_rank) h.values.push_back((double)rk / (double)mx);
}
double global_score(const Hit &h) {
    uint64_t rk = 1, mx = 1;
    bool sem = false;
    double sv = 0;
    for (int s = 0; s < h.n_scores; s++) {
        if (h.kind[s] == B200_S_VECTOR) {
            sem = true;
            sv = h.sim[s] < 0 ? 0.0 : (double)h.sim[s];
        } else
            rank_merge(rk, mx, h.rank[s], h.maxr[s]);
    }
    return sem ? sv : (double)rk / (double)mx;
}
int compare_scores(const Hit &l, float lr, const Hit &r, float rr) {  // hybrid.rs:32-80
    for (size_t i = 0;; i++) {
        bool hl = i < l.values.size(), hr = i < r.values.size();
        if (!hl && !hr) return 0;
        if (!hl) return -1;
        if (!hr) return 1;
        double a = l.values[i] * (double)lr, b = r.values[i] * (double)rr;
        if (std::fabs(a - b) <= 2.220446049250313e-16) continue;
        return a < b ? -1 : 1;
    }
}
}  // namespace

// Search::execute_hybrid (search/hybrid.rs:264-366)
int Engine::search_batch(const b200_query_batch *b, b200_results *r) {
    if (b->mode == 0) return keyword_batch(b, r, b->offset, b->limit, b->scoring_strategy);
    if (b->has_ranking_score_threshold || b->time_budget_ns || b->stop_after >= 0 || r->candidates)
        return fail(B200_ERR_UNSUPPORTED, "ranking-score threshold, deadlines and the candidates bitmap are implemented for keyword searches (mode 0) only");
    if (b->mode == 1) {
        int rc1 = semantic_batch(b, r, b->offset, b->limit);
        fold_vector_stats();
        return rc1;
    }
    if (b->mode != 2) return fail(B200_ERR_INVALID, "unknown search mode");
    const uint32_t NQ = b->n_queries, L = b->limit + b->offset, lim = std::max(1u, L);
    struct Side {
        std::vector<uint32_t> docids, n_hits, rank, maxr;
        std::vector<uint8_t> n_scores, kind;
        std::vector<float> sim;
        std::vector<uint64_t> n_cand;
        std::vector<int32_t> status;
        std::vector<uint8_t> neg;
        b200_results view;
        Side(uint32_t nq, uint32_t l)
            : docids((size_t)nq * l), n_hits(nq), rank((size_t)nq * l * B200_MAX_SCORES), maxr((size_t)nq * l * B200_MAX_SCORES),
              n_scores((size_t)nq * l), kind((size_t)nq * l * B200_MAX_SCORES), sim((size_t)nq * l * B200_MAX_SCORES), n_cand(nq), status(nq), neg(nq) {
            view = b200_results{docids.data(), n_hits.data(), n_scores.data(), kind.data(), rank.data(), maxr.data(), sim.data(), n_cand.data(), nullptr, status.data(),
                                nullptr, neg.data(), nullptr, 0};
        }
    };
    Side kw(NQ, lim), vec(NQ, lim);
    // the vector stage runs beside the keyword stage: its own host thread, stream and timers (the two share nothing mutable)
    bool have_vec = b->vectors != nullptr;
    int rc_vec = B200_OK;
    // It starts once the keyword stage has derived its terms (kw_derived counts the derivation waves done; B200_VEC_START picks the
    // wave, 0 = at once): the term sweep is a short, SM-filling kernel that the persistent GEMM would otherwise hold up for its whole
    // run time, while the step loop that follows leaves part of the GPU idle.
    // B200_HYBRID_SERIAL=1 runs the two stages one after the other (bench.py's device-time pass: kernel intervals must not overlap).
    std::thread sem;
    const bool serial = getenv("B200_HYBRID_SERIAL") != nullptr;
    const int vec_start = getenv("B200_VEC_START") ? atoi(getenv("B200_VEC_START")) : VEC_START_DEFAULT;
    kw_derived.store(0);
    if (have_vec && !serial)
        sem = std::thread([&]() {
            while (kw_derived.load(std::memory_order_acquire) < vec_start) std::this_thread::yield();
            rc_vec = semantic_batch(b, &vec.view, 0, lim);
        });
    int rc = keyword_batch(b, &kw.view, 0, lim, 1);
    kw_derived.store(KW_DERIVED_ALL, std::memory_order_release);  // (also when the keyword stage returned early)
    if (sem.joinable()) sem.join();
    if (have_vec && serial) rc_vec = semantic_batch(b, &vec.view, 0, lim);
    fold_vector_stats();
    if (rc != B200_OK) return rc;
    if (rc_vec != B200_OK) return rc_vec;
    float kr = 1.0f - b->semantic_ratio, vr = b->semantic_ratio;
    auto load = [&](const Side &s, uint32_t q, uint32_t i) {
        Hit h;
        size_t at = (size_t)q * lim + i;
        h.doc = s.docids[at];
        h.n_scores = s.n_scores[at];
        for (int k = 0; k < h.n_scores; k++) {
            h.kind[k] = s.kind[at * B200_MAX_SCORES + k];
            h.rank[k] = s.rank[at * B200_MAX_SCORES + k];
            h.maxr[k] = s.maxr[at * B200_MAX_SCORES + k];
            h.sim[k] = s.sim[at * B200_MAX_SCORES + k];
        }
        fill_values(h);
        return h;
    };
    auto merge_one = [&](size_t qi) {
        const uint32_t q = (uint32_t)qi;
        if (r->status) r->status[q] = kw.status[q];
        if (r->degraded) r->degraded[q] = 0;
        if (r->used_negative_operator) r->used_negative_operator[q] = kw.neg[q];
        if (kw.status[q] != 0) {
            r->n_hits[q] = 0;
            return;
        }
        std::vector<Hit> K, V;
        for (uint32_t i = 0; i < kw.n_hits[q]; i++) K.push_back(load(kw, q, i));
        // results_good_enough (:368-386)
        bool good = K.size() >= L;
        if (good)
            for (auto &h : K)
                if (global_score(h) * (double)kr < 0.45) {
                    good = false;
                    break;
                }
        std::vector<const Hit *> merged;
        uint32_t sem = 0;
        bool semantic_used = false;
        if (good || !have_vec) {
            for (auto &h : K) merged.push_back(&h);
            if (merged.size() > b->offset)
                merged.erase(merged.begin(), merged.begin() + b->offset);
            else
                merged.clear();
            if (merged.size() > b->limit) merged.resize(b->limit);
        } else {
            semantic_used = true;
            for (uint32_t i = 0; i < vec.n_hits[q]; i++) V.push_back(load(vec, q, i));
            size_t vi = 0, ki = 0, skipped = 0;
            std::vector<uint32_t> seen;
            while ((vi < V.size() || ki < K.size()) && merged.size() < b->limit) {
                bool take_vec = vi >= V.size() ? false : (ki >= K.size() ? true : compare_scores(V[vi], vr, K[ki], kr) >= 0);
                const Hit *h = take_vec ? &V[vi++] : &K[ki++];
                if (std::find(seen.begin(), seen.end(), h->doc) != seen.end()) continue;
                seen.push_back(h->doc);
                if (skipped < b->offset) {
                    skipped++;
                    continue;
                }
                if (take_vec) sem++;
                merged.push_back(h);
            }
        }
        r->n_hits[q] = (uint32_t)merged.size();
        if (r->semantic_hits) r->semantic_hits[q] = semantic_used ? sem : 0;
        if (r->n_candidates) r->n_candidates[q] = semantic_used ? std::max<uint64_t>(kw.n_cand[q], vec.n_cand[q]) : kw.n_cand[q];
        for (size_t i = 0; i < merged.size(); i++) {
            size_t at = (size_t)q * b->limit + i;
            const Hit &h = *merged[i];
            r->docids[at] = h.doc;
            if (r->n_scores) {
                r->n_scores[at] = h.n_scores;
                for (int k = 0; k < h.n_scores; k++) {
                    r->score_kind[at * B200_MAX_SCORES + k] = h.kind[k];
                    r->score_rank[at * B200_MAX_SCORES + k] = h.rank[k];
                    r->score_max[at * B200_MAX_SCORES + k] = h.maxr[k];
                    r->score_sim[at * B200_MAX_SCORES + k] = h.sim[k];
                }
            }
        }
    };
    // the merges of different queries are independent: spread them over the worker pool
    if (pool)
        pool->run(NQ, merge_one);
    else
        for (uint32_t q = 0; q < NQ; q++) merge_one(q);
    return B200_OK;
}

}  // namespace b200
