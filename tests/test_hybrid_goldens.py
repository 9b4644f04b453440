"""Hybrid merge known answers of the reference: crates/meilisearch/tests/search/hybrid.rs:195-430 (`simple_search`): three documents
with user-provided 2-d vectors, q = "Captain", vector = [1, 1], semanticRatio 0.2 / 0.5 / 0.8 — hit order, `_rankingScore` and
`semanticHitCount`.  Default settings: every field searchable with the same weight (searchableAttributes = ["*"]).
The oracle is pinned here; tests/test_gpu_parity.py::test_hybrid_goldens_on_gpu runs the same table through the C ABI."""
import numpy as np
import pytest

DOCS = [("Shazam!", "a Captain Marvel ersatz", "1", [1.0, 3.0]), ("Captain Planet", "He's not part of the Marvel Cinematic Universe", "2", [1.0, 2.0]),
        ("Captain Marvel", "a Shazam ersatz", "3", [2.0, 3.0])]
# semanticRatio -> (external ids in hit order, _rankingScore per hit or None where the snapshot has none, semanticHitCount)
HYBRID_CASES = {
    0.2: (["2", "3", "1"], None, 0),                                                              # hybrid.rs:201-270
    0.5: (["3", "2", "1"], [0.990290343761444, 0.9848484848484848, 0.9472135901451112], 2),      # hybrid.rs:274-346
    0.8: (["3", "2", "1"], [0.990290343761444, 0.974341630935669, 0.9472135901451112], 3),       # hybrid.rs:350-428
}


def hybrid_image():
    from corpus.pyindexgen import IndexImage

    img = IndexImage(3)  # document key order: title, desc, id
    for d, (title, desc, ext, _) in enumerate(DOCS):
        img.add_text(d, 0, title)
        img.add_text(d, 1, desc)
        img.add_text(d, 2, ext)
    return img.build()


def embeddings():
    return np.array([d[3] for d in DOCS], np.float32)


def check(ids, scores, sem, ratio):
    from tests.test_cutoff_goldens import global_score

    want_ids, want_scores, want_sem = HYBRID_CASES[ratio]
    assert [DOCS[i][2] for i in ids] == want_ids
    assert sem == want_sem
    if want_scores is not None:
        assert np.allclose([global_score(s) for s in scores], want_scores, rtol=0, atol=1e-12)


@pytest.mark.parametrize("ratio", list(HYBRID_CASES))
def test_oracle_hybrid_simple_search(ratio):
    from meilisearch_b200.tokenizer import TokenBatch
    from oracle.pyoracle import OracleIndex

    o = OracleIndex(hybrid_image(), weights=[0, 0, 0])
    o.set_embeddings(embeddings())
    r = o.search_batch(TokenBatch(["Captain"]), vectors=np.array([[1.0, 1.0]], np.float32), hybrid=True, semantic_ratio=ratio, scoring="detailed")
    check(r.ids(0), r.scores(0), int(r.semantic_hits[0]), ratio)


# The other tests of hybrid.rs on the same three documents.  The HTTP route picks the milli call from (q, vector, semanticRatio)
# (crates/meilisearch/src/routes/indexes/search.rs:981-1025 `search_kind`; crates/meilisearch/src/search/mod.rs:2133-2146):
# ratio 0 or neither q nor vector -> Search::execute (keyword / placeholder), semanticHitCount null; ratio 1 or no q -> Search::execute
# with only the vector (semanticHitCount = number of hits); otherwise Search::execute_hybrid(ratio).  `route()` restates that.
# case: (hybrid.rs lines, q, vector, ratio, offset, limit, distribution, n_docs, ids, _rankingScore per hit or None, semanticHitCount or None)
S3 = [0.990290343761444, 0.974341630935669, 0.9472135901451112]
MORE_CASES = [
    ("426-438 limit_offset", "Captain", [1.0, 1.0], 0.2, 1, 1, None, 3, ["3"], None, 0),
    ("441-452 limit_offset", "Captain", [1.0, 1.0], 0.9, 1, 1, None, 3, ["2"], None, 1),
    ("544-547 distribution_shift (before)", "Captain", [1.0, 1.0], 1.0, 0, 20, None, 3, ["3", "2", "1"], S3, 3),
    ("549-568 distribution_shift mean 0.998 sigma 0.01", "Captain", [1.0, 1.0], 1.0, 0, 20, (0.998, 0.01), 3, ["3", "2", "1"],
     [0.19161224365234375, 1.1920928955078125e-7, 1.1920928955078125e-7], 3),
    ("577-590 highlighter", "Captain Marvel", [1.0, 1.0], 0.2, 0, 20, None, 3, ["3", "1", "2"], None, 0),
    ("593-606 highlighter", "Captain Marvel", [1.0, 1.0], 0.8, 0, 20, None, 3, ["3", "2", "1"], S3, 3),
    ("610-623 highlighter", "Captain Marvel", [1.0, 1.0], 1.0, 0, 20, None, 3, ["3", "2", "1"], S3, 3),
    ("699-711 single_document", None, [1.0, 3.0], 1.0, 0, 20, None, 1, ["1"], [1.0], 1),
    ("721-726 query_combination: placeholder", None, None, 1.0, 0, 20, None, 3, ["1", "2", "3"], [1.0, 1.0, 1.0], None),
    ("730-735 query_combination: placeholder", None, None, 0.76, 0, 20, None, 3, ["1", "2", "3"], [1.0, 1.0, 1.0], None),
    ("754-759 query_combination: full vector", None, [1.0, 0.0], 1.0, 0, 20, None, 3, ["3", "2", "1"],
     [0.7773500680923462, 0.7236068248748779, 0.6581138968467712], 3),
    ("763-768 query_combination: vector, ratio 0", None, [1.0, 0.0], 0.0, 0, 20, None, 3, ["1", "2", "3"], [1.0, 1.0, 1.0], None),
    ("772-777 query_combination: q + vector, ratio 0", "Captain", [1.0, 0.0], 0.0, 0, 20, None, 3, ["2", "3", "1"],
     [0.9848484848484848, 0.9848484848484848, 0.9242424242424242], None),
]


def route(q, vector, ratio):
    """search_kind: 'keyword' | 'semantic' | 'hybrid'"""
    placeholder = q is None or not q.strip()
    if ratio == 0.0 or (placeholder and vector is None):
        return "keyword"
    if ratio == 1.0 or placeholder:
        return "semantic"
    return "hybrid"


def more_image(n_docs):
    from corpus.pyindexgen import IndexImage

    img = IndexImage(3)
    for d, (title, desc, ext, _) in enumerate(DOCS[:n_docs]):
        img.add_text(d, 0, title)
        img.add_text(d, 1, desc)
        img.add_text(d, 2, ext)
    return img.build()


def check_more(case, ids, scores, sem):
    from tests.test_cutoff_goldens import global_score

    _, q, vector, ratio, _, _, _, _, want_ids, want_scores, want_sem = case
    assert [DOCS[i][2] for i in ids] == want_ids, case[0]
    if want_sem is not None:
        assert sem == want_sem, case[0]
    if want_scores is not None:
        # the reference prints f32 similarities widened to f64: compare at f32 resolution
        assert np.allclose([global_score(s) for s in scores], want_scores, rtol=0, atol=2e-7), case[0]


@pytest.mark.parametrize("case", MORE_CASES, ids=[c[0].split()[0] + "_" + str(i) for i, c in enumerate(MORE_CASES)])
def test_oracle_hybrid_more(case):
    from meilisearch_b200.tokenizer import TokenBatch
    from oracle.pyoracle import OracleIndex

    _, q, vector, ratio, offset, limit, distribution, n_docs, _, _, _ = case
    o = OracleIndex(more_image(n_docs), weights=[0, 0, 0])
    o.set_embeddings(embeddings()[:n_docs], distribution=distribution)
    kind = route(q, vector, ratio)
    vec = None if vector is None else np.array([vector], np.float32)
    tb = TokenBatch([q or ""])
    if kind == "keyword":
        r = o.search_batch(tb, scoring="detailed", offset=offset, limit=limit)
        sem = None
    elif kind == "semantic":
        r = o.search_batch(TokenBatch([""]), vectors=vec, vector_only=True, scoring="detailed", offset=offset, limit=limit)
        sem = len(r.ids(0))  # search/mod.rs:2138-2141
    else:
        r = o.search_batch(tb, vectors=vec, hybrid=True, semantic_ratio=ratio, scoring="detailed", offset=offset, limit=limit)
        sem = int(r.semantic_hits[0])
    check_more(case, r.ids(0), r.scores(0), sem)
