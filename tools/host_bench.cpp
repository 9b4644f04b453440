// Host-side micro-benchmark (dev tool, no GPU needed): times the per-activation host work of the engine —
// prepare_graph_rule + emit_activation_work (+ exact attribute) — on the query graphs of a synthetic batch.
// Build: see tools/host_bench.sh.  Not part of the product or of the tests.
#include "../meilisearch_b200/csrc/engine_search.cpp"

#include <chrono>
#include <cstdio>
#include <sstream>

#include "../corpus/indexgen.h"

using namespace b200;

int main(int argc, char **argv) {
    uint32_t n_docs = argc > 1 ? atoi(argv[1]) : 200000, vocab = argc > 2 ? atoi(argv[2]) : 100000;
    int reps = argc > 3 ? atoi(argv[3]) : 20;
    ig_builder *b = ig_new(1, 0);
    ig_add_synthetic(b, n_docs, vocab, 1.07, 3, 15, 0xB200);
    ig_build(b);
    const uint8_t *db; const uint64_t *doff;
    ig_dictionary(b, &db, &doff);
    uint64_t nw = ig_n_words(b);
    std::vector<uint8_t> dict_bytes(db, db + doff[nw]);
    std::vector<uint64_t> dict_off(doff, doff + nw + 1);
    RawDb dbs[10];
    for (int i = 0; i < 10; i++) {
        ig_db_view v;
        ig_db(b, i, &v);
        dbs[i].n = v.n_keys;
        dbs[i].keys.assign(v.key_bytes, v.key_bytes + v.key_offsets[v.n_keys]);
        dbs[i].vals.assign(v.val_bytes, v.val_bytes + v.val_offsets[v.n_keys]);
        dbs[i].koff.assign(v.key_offsets, v.key_offsets + v.n_keys + 1);
        dbs[i].voff.assign(v.val_offsets, v.val_offsets + v.n_keys + 1);
    }
    const uint8_t *dc; uint64_t dl;
    ig_documents_ids(b, &dc, &dl);
    std::vector<uint8_t> docids(dc, dc + dl);
    HostIndex hix;
    build_host_index(dict_bytes, dict_off, dbs, docids, hix);
    hix.pool.clear();
    // queries
    char *qs = ig_synthetic_queries(b, 1024, 0, 1);
    std::vector<std::string> queries;
    {
        std::stringstream ss(qs);
        std::string line;
        while (std::getline(ss, line)) queries.push_back(line);
    }
    std::vector<uint32_t> token_begin{0}, lemma_off{0};
    std::vector<uint8_t> kinds;
    std::string lemmas;
    for (auto &q : queries) {
        std::stringstream ss(q);
        std::string w;
        bool first = true;
        while (ss >> w) {
            if (!first) {
                kinds.push_back(2);
                lemmas += " ";
                lemma_off.push_back((uint32_t)lemmas.size());
            }
            first = false;
            kinds.push_back(0);
            lemmas += w;
            lemma_off.push_back((uint32_t)lemmas.size());
        }
        token_begin.push_back((uint32_t)kinds.size());
    }
    b200_query_batch qb{};
    qb.n_queries = (uint32_t)queries.size();
    qb.token_begin = token_begin.data();
    qb.token_kind = kinds.data();
    qb.lemma_off = lemma_off.data();
    qb.lemma_bytes = lemmas.data();
    qb.limit = 20;
    qb.words_limit = 10;
    std::vector<std::unique_ptr<QState>> st;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < qb.n_queries; i++) {
        st.emplace_back(new QState(hix));
        parse_query(*st.back(), &qb, i);
    }
    auto ms = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    printf("parse: %.2f ms for %u queries\n", ms(t0), qb.n_queries);
    const int kinds_to_run[] = {RK_WORDS, RK_TYPO, RK_PROXIMITY, RK_FID, RK_POSITION, RK_EXACTNESS};
    const char *names[] = {"words", "typo", "proximity", "fid", "position", "exactness"};
    for (int ki = 0; ki < 6; ki++) {
        double t_prep = 0, t_emit = 0;
        size_t n = 0, jobs = 0;
        for (int r = 0; r < reps; r++)
            for (auto &q : st) {
                Level L;
                L.kind = kinds_to_run[ki];
                L.graph = q->graph;
                auto a = std::chrono::steady_clock::now();
                prepare_graph_rule(q->ctx, L.kind, L.kind == RK_WORDS, B200_TMS_LAST, L);
                t_prep += ms(a);
                StepOut o;
                a = std::chrono::steady_clock::now();
                emit_activation_work(q->ctx, L, o);
                t_emit += ms(a);
                n++;
                jobs += o.jobs.size();
            }
        printf("%-10s prepare %.2f us/act   emit %.2f us/act   (%.1f jobs/act)\n", names[ki], 1e3 * t_prep / n, 1e3 * t_emit / n, (double)jobs / n);
    }
    return 0;
}
