"""Pin the CPU oracle against golden vectors extracted from the reference's own tests
(crates/milli/src/search/new/tests/*.rs — see tests/golden/extract_reference_goldens.py) and the
vector-score goldens of SURVEY.md Appendix C.3."""
import numpy as np
import pytest

from meilisearch_b200.tokenizer import TokenBatch
from oracle.pyoracle import OracleIndex, cbo_decode
from tests.helpers import image_from_corpus, load_goldens

G = load_goldens()
_images = {}


def _image(ci):
    if ci not in _images:
        _images[ci] = image_from_corpus(G["corpora"][ci])
    return _images[ci]


def check_golden(case, ids, scores, n_candidates):
    """what a golden case pins: hit order; ScoreDetails tuples; `_rankingScore`; per-rule scores; `estimatedTotalHits`"""
    from tests.test_cutoff_goldens import global_score

    if case["expected_ids"] is not None:
        assert ids == case["expected_ids"], case["source"]
    for doc, want in case.get("expected_doc_scores", []):
        assert doc in ids, case["source"]
        assert abs(global_score(scores[ids.index(doc)]) - want) < 1e-14, case["source"]
    if "expected_scores" in case and case["scoring"] == "detailed":
        assert [[list(x) for x in row] for row in scores] == case["expected_scores"], case["scores_source"]
    if "expected_ranking_scores" in case:  # rationals of small integers: the f64 value is the reference's to the last digits it prints
        assert np.allclose([global_score(s) for s in scores], case["expected_ranking_scores"], rtol=0, atol=1e-14), case["source"]
    if "expected_rule_scores" in case:
        for row, want in zip(scores, case["expected_rule_scores"]):
            got = [rk / mx for _, rk, mx in row]
            # the reference reports ExactAttribute + ExactWords as one `exactness` entry (score_details.rs): its score is
            # ExactAttribute's unless the match type is exact
            assert np.allclose(got[:len(want)], want, rtol=0, atol=1e-14), case["source"]
    if "expected_candidates" in case:
        assert n_candidates == case["expected_candidates"], case["source"]


@pytest.mark.parametrize("case", G["cases"], ids=[c["source"].split("/")[-1] + ":" + c["query"][:24] for c in G["cases"]])
def test_reference_golden(case):
    img = _image(case["index"])
    s = case["settings"]
    ix = OracleIndex(img, criteria=s.get("criteria"), authorize_typos=s.get("authorize_typos", True),
                     one_typo=s.get("one_typo", 5), two_typos=s.get("two_typos", 9), weights=s.get("weights"))
    ix.update_settings(exact_words=s.get("exact_words", []), synonyms=s.get("synonyms", {}))
    r = ix.search_batch(TokenBatch([case["query"]], img.stop_words), tms=case["tms"], scoring=case["scoring"],
                        limit=max(case["limit"], 1), offset=case["offset"], threshold=case.get("threshold"))
    check_golden(case, r.ids(0), r.scores(0), int(r.n_candidates[0]))


@pytest.mark.parametrize("case", G.get("count_cases", []), ids=[c["source"].split("/")[-1] + ":" + c["query"] for c in G.get("count_cases", [])])
def test_reference_hit_counts(case):
    """crates/milli/tests/search/typo_tolerance.rs asserts only `documents_ids.len()`."""
    img = _image(case["index"])
    s = case["settings"]
    ix = OracleIndex(img, criteria=s.get("criteria"), authorize_typos=s.get("authorize_typos", True),
                     one_typo=s.get("one_typo", 5), two_typos=s.get("two_typos", 9))
    ix.update_settings(exact_words=s.get("exact_words", []), synonyms=s.get("synonyms", {}))
    r = ix.search_batch(TokenBatch([case["query"]], img.stop_words), tms=case["tms"], scoring=case["scoring"], limit=case["limit"])
    assert len(r.ids(0)) == case["expected_count"], case["source"]


def test_typo_bucketing_scores():
    # crates/milli/src/search/new/tests/snapshots/milli__search__new__tests__typo__typo_bucketing-5.snap:
    # Typo{typo_count, max_typo_count=5} = 0,0,1,1,2,5  
    case = next(c for c in G["cases"] if c["expected_ids"] == [16, 18, 17, 20, 15, 14])
    img = _image(case["index"])
    ix = OracleIndex(img, criteria=["typo"])
    r = ix.search_batch(TokenBatch([case["query"]]), tms="all", scoring="detailed")
    assert r.ids(0) == [16, 18, 17, 20, 15, 14]
    typo_counts = [mx - rk for ((_, rk, mx),) in r.scores(0)]
    assert typo_counts == [0, 0, 1, 1, 2, 5]
    assert all(mx - 1 == 5 for ((_, _, mx),) in r.scores(0))


def _vec_index(vectors, docids=None):
    from corpus.pyindexgen import IndexImage
    img = IndexImage(1)
    n = len(vectors) + 1
    for d in range(n):
        img.add_text(d, 0, "doc")
    img.build()
    ix = OracleIndex(img)
    ix.set_embeddings(np.asarray(vectors, np.float32), docids)
    return ix


def test_vector_golden_cutoff_rs():
    # crates/milli/src/search/new/tests/cutoff.rs:509-600: q=[1,-1]; ids [2,0,3,1], similarities 1.0,0.5,0.5,0.0
    ix = _vec_index([[0.1, 0.1], [-0.1, 0.1], [0.1, -0.1], [-0.1, -0.1]])
    r = ix.search_batch(TokenBatch([""]), vectors=np.asarray([[1.0, -1.0]], np.float32), vector_only=True, limit=4, scoring="detailed")
    assert r.ids(0) == [2, 0, 3, 1]
    sims = [s[0][1] for s in r.scores(0)]
    assert sims == pytest.approx([1.0, 0.5, 0.5, 0.0], abs=1e-6)


def test_vector_golden_hybrid_rs():
    # crates/meilisearch/tests/search/hybrid.rs:262-330: q=[1,1]; [2,3] -> 0.990290343761444, [1,3] -> 0.9472135901451112
    ix = _vec_index([[2.0, 3.0], [1.0, 3.0]])
    ids, dist = ix.nns(np.asarray([1.0, 1.0], np.float32), 2)
    assert list(ids) == [0, 1]
    assert float(np.float32(1.0) - dist[0]) == pytest.approx(0.990290343761444, rel=1e-6)
    assert float(np.float32(1.0) - dist[1]) == pytest.approx(0.9472135901451112, rel=1e-6)


def test_cbo_codec_threshold():
    # cbo_roaring_bitmap_codec.rs:186-256: <=7 ints are raw native-endian u32; above that, portable roaring
    from corpus.pyindexgen import IndexImage
    img = IndexImage(1)
    for d in range(10):
        img.add_text(d, 0, "few " + ("many" if d < 8 else ""))
    img.add_text(3, 0, "few")
    img.build()
    db = img.db("word_docids")
    vals = {db.key(i): db.val(i) for i in range(db.n_keys)}
    assert len(vals[b"many"]) > 28 and int.from_bytes(vals[b"many"][:4], "little") == 12346
    assert list(cbo_decode(vals[b"many"])) == list(range(8))
    few = cbo_decode(vals[b"few"])
    assert list(few) == list(range(10))
    img2 = IndexImage(1)
    for d in range(7):
        img2.add_text(d * 70000, 0, "seven")
    img2.build()
    v = img2.db("word_docids").val(0)
    assert len(v) == 28 and list(cbo_decode(v)) == [d * 70000 for d in range(7)]
