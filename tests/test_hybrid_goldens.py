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
