"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI vs the CPU oracle and the
reference goldens.  Integer/bit work must match exactly; vector scores within 1e-4 relative."""
import numpy as np
import pytest

from tests.helpers import image_from_corpus, load_goldens, synthetic_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import meilisearch_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def synth():
    return synthetic_image(60000, 25000, seed=11)


def _edit(rng, w):
    k = rng.integers(4)
    if len(w) < 2:
        k = 1
    p = int(rng.integers(len(w)))
    c = chr(ord("a") + int(rng.integers(26)))
    if k == 0:
        return w[:p] + c + w[p + 1:]
    if k == 1:
        return w[:p] + c + w[p:]
    if k == 2:
        return w[:p] + w[p + 1:]
    p = min(p, len(w) - 2)
    return w[:p] + w[p + 1] + w[p] + w[p + 2:]


def test_derive_matches_oracle(mb, synth):
    from oracle.pyoracle import OracleIndex

    rng = np.random.default_rng(5)
    words = []
    for _ in range(300):
        w = synth.word(int(rng.integers(synth.n_words)))
        for _ in range(int(rng.integers(3))):
            w = _edit(rng, w) or w
        if len(w) >= 3:
            words.append(w)
    words += ["a" * 5, "zzzzzzzzz", synth.word(0), synth.word(synth.n_words - 1)]
    ix = mb.Index(synth)
    o = OracleIndex(synth)
    for max_typo in (1, 2):
        for is_prefix in (0, 1):
            got = ix.derive(words, [max_typo] * len(words), [is_prefix] * len(words))
            for w, (g1, g2) in zip(words, got):
                o1, o2 = o.derive(w, max_typo, is_prefix)
                assert list(g1) == list(o1), (w, max_typo, is_prefix, "one")
                assert list(g2) == list(o2), (w, max_typo, is_prefix, "two")


def test_derive_caps(mb):
    # MAX_ONE_TYPO_COUNT / MAX_TWO_TYPOS_COUNT and the first-letter quirk (limits.rs:7-9, compute_derivations.rs:129-166)
    from corpus.pyindexgen import IndexImage
    from oracle.pyoracle import OracleIndex

    img = IndexImage(1)
    base = "abcdefghij"
    docs = []
    for a in "abcdefghijklmnopqrstuvwxyz":
        for b in "abcdefghijklmnopqrstuvwxyz":
            docs.append(base[:4] + a + base[5:8] + b + base[9:])   # <= 2 substitutions, same first letter
        docs.append(a + base[1:])                                   # first-letter substitutions
    img.add_text(0, 0, " ".join(docs))
    img.build()
    ix, o = mb.Index(img), OracleIndex(img)
    for w, p in ((base, 0), (base, 1), ("xbcdefghij", 0)):
        (g1, g2), = ix.derive([w], [2], [p])
        o1, o2 = o.derive(w, 2, p)
        assert list(g1) == list(o1) and list(g2) == list(o2)
        assert len(o1) <= 150 and len(o2) <= 50


def test_nns_matches_oracle(mb, synth):
    from oracle.pyoracle import OracleIndex

    rng = np.random.default_rng(1)
    n, d = 30000, 768
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb[100] = emb[50]          # an exact tie: equal distances must come out in docid order
    docids = rng.permutation(n).astype(np.uint32)
    ix = mb.Index(synth)
    ix.set_embeddings(emb, docids)
    o = OracleIndex(synth)
    o.set_embeddings(emb.astype(np.float16).astype(np.float32), docids)   # the oracle scans the same fp16-rounded rows
    q = rng.standard_normal((11, d)).astype(np.float32)
    q[3] = emb[50]
    cand = np.zeros((n + 63) // 64, np.uint64)
    keep = rng.random(n) < 0.1
    for doc in np.nonzero(keep)[0]:
        cand[doc >> 6] |= np.uint64(1) << np.uint64(doc & 63)
    for cw in (None, cand):
        ids, dist, cnt = ix.nns_by_vector(q, 100, cw)
        for i in range(len(q)):
            oid, od = o.nns(q[i], 100, cw)
            assert cnt[i] == len(oid)
            # 1e-4 relative on the similarity score (north_star); ids equal except where oracle scores tie within tolerance
            assert np.allclose(1 - dist[i, : cnt[i]], 1 - od, rtol=1e-4, atol=1e-6)
            diff = [k for k in range(len(oid)) if ids[i, k] != oid[k]]
            for k in diff:
                assert abs(od[k] - dist[i, k]) <= 1e-4 * max(1 - od[k], 1e-3)


def test_nns_tensor_core_batch_matches_oracle(mb, synth, monkeypatch):
    """The batched vector stage (tcgen05 GEMM + fused top-k) against the oracle and against the GEMV scan."""
    from oracle.pyoracle import OracleIndex

    rng = np.random.default_rng(7)
    for n, d, nq, k in ((20011, 768, 150, 100), (5000, 128, 40, 7), (300, 64, 17, 128)):
        emb = rng.standard_normal((n, d)).astype(np.float32)
        emb[100] = emb[50]
        docids = rng.permutation(n).astype(np.uint32)
        ix = mb.Index(synth)
        ix.set_embeddings(emb, docids)
        o = OracleIndex(synth)
        o.set_embeddings(emb.astype(np.float16).astype(np.float32), docids)
        q = rng.standard_normal((nq, d)).astype(np.float32)
        q[3] = emb[50]
        cand = np.zeros((n + 63) // 64, np.uint64)
        for doc in np.nonzero(rng.random(n) < 0.2)[0]:
            cand[doc >> 6] |= np.uint64(1) << np.uint64(doc & 63)
        for cw in (None, cand):
            monkeypatch.setenv("B200_VEC_GEMM", "1")
            ix.reset_stats()
            ids, dist, cnt = ix.nns_by_vector(q, k, cw)
            assert ix.stats()["kernels"]["vec_gemm_topk"]["count"] >= 1
            monkeypatch.setenv("B200_VEC_GEMM", "0")
            ids2, dist2, cnt2 = ix.nns_by_vector(q, k, cw)
            assert (cnt == cnt2).all()
            # queries are rounded to fp16 for the tensor cores: 1e-4 relative on the similarity (north_star)
            for i in range(nq):
                oid, od = o.nns(q[i], k, cw)
                assert cnt[i] == len(oid)
                assert np.allclose(1 - dist[i, : cnt[i]], 1 - od, rtol=1e-4, atol=2e-5), (n, d, i)
                assert (np.diff(dist[i, : cnt[i]]) >= 0).all()
                assert len(set(ids[i, : cnt[i]].tolist())) == cnt[i]
                for j in np.nonzero(ids[i, : cnt[i]] != oid)[0]:
                    assert abs(od[j] - dist[i, j]) <= 1e-4 * max(1 - od[j], 1e-3) + 2e-5
            assert np.allclose(dist[:, : cnt.min()], dist2[:, : cnt.min()], rtol=1e-4, atol=2e-5)


G = load_goldens()


def test_reference_goldens_on_gpu(mb):
    """Every golden extracted from the reference's ranking-rule tests (docids and, where snapshotted, ScoreDetails) through the CUDA path."""
    ran = 0
    images = {}
    for case in G["cases"]:
        s = case["settings"]
        ci = case["index"]
        if ci not in images:
            images[ci] = image_from_corpus(G["corpora"][ci])
        img = images[ci]
        ix = mb.Index(img, criteria=s.get("criteria"), authorize_typos=s.get("authorize_typos", True), one_typo=s.get("one_typo", 5),
                      two_typos=s.get("two_typos", 9), exact_words=s.get("exact_words", []), synonyms=s.get("synonyms"))
        res = (ix.search().query(mb.TokenBatch([case["query"]], img.stop_words)).terms_matching_strategy(case["tms"])
               .scoring_strategy(case["scoring"]).limit(max(case["limit"], 1)).offset(case["offset"]).execute())
        assert res.status[0] == 0
        assert res.ids(0) == case["expected_ids"], case["source"]
        if "expected_scores" in case and case["scoring"] == "detailed":
            assert [[list(x) for x in row] for row in res.scores(0)] == case["expected_scores"], case["scores_source"]
        ix.close()
        ran += 1
    assert ran == len(G["cases"])


@pytest.mark.parametrize("tms,scoring", [("last", "detailed"), ("last", "skip"), ("all", "detailed"), ("frequency", "detailed")])
def test_keyword_batch_matches_oracle(mb, synth, tms, scoring):
    from oracle.pyoracle import OracleIndex

    queries = synth.synthetic_queries(300, seed=21) + ["", "   ", synth.word(5), "zzzzqqqq xxxxyyyy"]
    tokens = mb.TokenBatch(queries)
    ix = mb.Index(synth)
    got = ix.search().query(tokens).terms_matching_strategy(tms).scoring_strategy(scoring).execute()
    want = OracleIndex(synth).search_batch(tokens, tms=tms, scoring=scoring, n_threads=8)
    for q in range(len(queries)):
        assert got.status[q] == 0
        assert got.ids(q) == want.ids(q), (queries[q], tms, scoring)
        assert got.scores(q) == want.scores(q), (queries[q], tms, scoring)
        assert int(got.n_candidates[q]) == int(want.n_candidates[q]), queries[q]


def test_offset_limit(mb, synth):
    from oracle.pyoracle import OracleIndex

    queries = synth.synthetic_queries(40, seed=4, with_typos=False)
    tokens = mb.TokenBatch(queries)
    ix, o = mb.Index(synth), OracleIndex(synth)
    for off, lim in ((0, 5), (3, 7), (15, 20), (0, 100)):
        got = ix.search().query(tokens).offset(off).limit(lim).execute()
        want = o.search_batch(tokens, offset=off, limit=lim)
        for q in range(len(queries)):
            assert got.ids(q) == want.ids(q), (queries[q], off, lim)


def test_two_fields_full_stack(mb):
    from oracle.pyoracle import OracleIndex

    img = synthetic_image(8000, 3000, seed=3, n_fields=2)
    queries = img.synthetic_queries(120, seed=8)
    tokens = mb.TokenBatch(queries)
    crit = ["words", "typo", "proximity", "attribute", "exactness"]
    ix, o = mb.Index(img, criteria=crit), OracleIndex(img, criteria=crit)
    got = ix.search().query(tokens).scoring_strategy("detailed").execute()
    want = o.search_batch(tokens, scoring="detailed")
    for q in range(len(queries)):
        assert got.ids(q) == want.ids(q), queries[q]
        assert got.scores(q) == want.scores(q), queries[q]


def test_hybrid_matches_oracle(mb, synth):
    from oracle.pyoracle import OracleIndex

    rng = np.random.default_rng(2)
    n, d = synth.n_docs, 64
    emb = rng.standard_normal((n, d)).astype(np.float16).astype(np.float32)
    queries = synth.synthetic_queries(50, seed=9)
    tokens = mb.TokenBatch(queries)
    vec = rng.standard_normal((len(queries), d)).astype(np.float16).astype(np.float32)
    ix, o = mb.Index(synth), OracleIndex(synth)
    ix.set_embeddings(emb)
    o.set_embeddings(emb)
    for ratio in (0.1, 0.5, 0.9):
        got = ix.search().query(tokens).semantic(vec).execute_hybrid(ratio)
        want = o.search_batch(tokens, vectors=vec, hybrid=True, semantic_ratio=ratio)
        for q in range(len(queries)):
            assert got.ids(q) == want.ids(q), (queries[q], ratio)
            assert int(got.semantic_hit_count[q]) == int(want.semantic_hits[q])


def test_phrases_negatives_match_oracle(mb, synth):
    from oracle.pyoracle import OracleIndex

    base = synth.synthetic_queries(60, seed=33, with_typos=False)
    queries = []
    for i, q in enumerate(base):
        w = q.split()
        if i % 3 == 0 and len(w) >= 2:
            queries.append('"' + " ".join(w[:2]) + '" ' + " ".join(w[2:]))
        elif i % 3 == 1 and len(w) >= 2:
            queries.append(" ".join(w[:-1]) + " -" + w[-1])
        else:
            queries.append(w[0] + ' "' + " ".join(w[1:]) + '"')
    tokens = mb.TokenBatch(queries)
    ix, o = mb.Index(synth), OracleIndex(synth)
    got = ix.search().query(tokens).scoring_strategy("detailed").execute()
    want = o.search_batch(tokens, scoring="detailed", n_threads=8)
    for q in range(len(queries)):
        assert got.status[q] == 0
        assert got.ids(q) == want.ids(q), queries[q]
        assert got.scores(q) == want.scores(q), queries[q]


def test_unsupported_is_reported_not_faked(mb, synth):
    ix = mb.Index(synth)
    w = synth.synthetic_queries(1, seed=5, with_typos=False)[0].split()[0]
    long_query = " ".join(synth.synthetic_queries(8, seed=6, with_typos=False))
    assert len(long_query.split()) > 12
    res = ix.search().query(["-" + w, "plain", long_query]).words_limit(20).execute()
    assert res.status[0] == -4 and res.n_hits[0] == 0
    assert res.status[1] == 0
    assert res.status[2] == -4 and res.n_hits[2] == 0
    # with the default words_limit (10) the same long query is in scope
    assert ix.search().query([long_query]).execute().status[0] == 0


def test_large_universe_matches_oracle(mb):
    """More than 8192 x 64 documents: universes span several compaction segments and the emit scan runs many rounds."""
    from oracle.pyoracle import OracleIndex
    from tests.helpers import synthetic_image

    img = synthetic_image(700_000, 60_000, seed=0xB201)
    queries = img.synthetic_queries(48, seed=9)
    tokens = mb.TokenBatch(queries)
    got = mb.Index(img).search().query(tokens).scoring_strategy("detailed").execute()
    want = OracleIndex(img).search_batch(tokens, scoring="detailed", n_threads=8)
    for q in range(len(queries)):
        assert got.status[q] == 0
        assert got.ids(q) == want.ids(q), queries[q]
        assert got.scores(q) == want.scores(q), queries[q]
        assert int(got.n_candidates[q]) == int(want.n_candidates[q])


def test_union_postings_matches_decoded_lists(mb, synth):
    """S2 (b200_union_postings): OR of posting lists AND universe, against the CBO values decoded by the oracle's codec."""
    from oracle.pyoracle import cbo_decode

    ix = mb.Index(synth)
    rng = np.random.default_rng(3)
    n_words = (synth.n_docs + 63) // 64
    for db in (0, 4, 5):  # word_docids, word_pair_proximity_docids, word_position_docids
        view = synth.dbs[db]
        n_keys = int(view.n_keys)
        # a mix of the longest lists (dense on the device) and random ones
        lens = np.diff(np.asarray(view.val_offsets))
        keys = np.unique(np.concatenate([np.argsort(lens)[-3:], rng.integers(0, n_keys, 40)])).astype(np.uint32)
        want = np.zeros(n_words, np.uint64)
        for k in keys:
            ids = cbo_decode(view.val(int(k)))
            np.bitwise_or.at(want, ids >> 6, np.uint64(1) << (ids & 63).astype(np.uint64))
        got = ix.union_postings(db, keys)
        assert np.array_equal(got, want), db
        universe = rng.integers(0, 2**63, n_words, dtype=np.uint64)
        got_u = ix.union_postings(db, keys, universe)
        assert np.array_equal(got_u, want & universe), db
    assert np.array_equal(ix.union_postings(0, np.zeros(0, np.uint32)), np.zeros(n_words, np.uint64))
