"""Deadline / degraded results: the reference's own known answers (crates/milli/src/search/new/tests/cutoff.rs:20-56,100-407):
5 documents inserted in reverse id order (internal docids 0..4), criteria [words, typo], query "hello puppy kefir", limit 4,
ScoringStrategy::Detailed, Deadline::never().with_stop_after(n) — ids and `ScoreDetails::global_score` per hit.
The oracle is pinned on them here; tests/test_gpu_parity.py::test_cutoff_goldens_on_gpu runs the same table through the C ABI."""
import pytest

DOCS = ["hella puppo kefir", "hella puppy kefir", "hello", "hello puppy", "hello puppy kefir"]
# stop_after -> (documents_ids, global scores rounded to 4 digits, degraded)
CUTOFF_CASES = {
    None: ([4, 1, 0, 3], [1.0, 0.9167, 0.8333, 0.6667], False),
    1: ([0, 1, 4, 2], [0.6667, 0.6667, 0.6667, 0.0], True),
    2: ([4, 0, 1, 2], [1.0, 0.6667, 0.6667, 0.0], True),
    3: ([4, 1, 0, 2], [1.0, 0.9167, 0.6667, 0.0], True),
    4: ([4, 1, 0, 2], [1.0, 0.9167, 0.8333, 0.0], True),
    5: ([4, 1, 0, 3], [1.0, 0.9167, 0.8333, 0.3333], True),
    6: ([4, 1, 0, 3], [1.0, 0.9167, 0.8333, 0.6667], False),  # the search completes before the 7th poll (the reference asserts ids and scores only)
}


def cutoff_image():
    from corpus.pyindexgen import IndexImage

    img = IndexImage(1)
    for d, t in enumerate(DOCS):
        img.add_text(d, 0, t)
    return img.build()


def global_score(details):
    """ScoreDetails::global_score (score_details.rs:133-154) over [(kind, rank, max) | ('vector', sim)]"""
    rk, mx, sem = 1, 1, None
    for s in details:
        if s[0] == "vector":
            sem = s[1] or 0.0
        else:
            rk = max(rk - 1, 0) * s[2] + s[1]
            mx *= s[2]
    return sem if sem is not None else rk / mx


@pytest.mark.parametrize("stop_after", list(CUTOFF_CASES))
def test_oracle_cutoff(stop_after):
    from meilisearch_b200.tokenizer import TokenBatch
    from oracle.pyoracle import OracleIndex

    o = OracleIndex(cutoff_image(), criteria=["words", "typo"])
    r = o.search_batch(TokenBatch(["hello puppy kefir"]), scoring="detailed", limit=4, stop_after=stop_after)
    ids, scores, degraded = CUTOFF_CASES[stop_after]
    assert r.ids(0) == ids
    assert [round(global_score(s), 4) for s in r.scores(0)] == scores
    assert bool(r.degraded[0]) == degraded


def test_oracle_degraded_search_cannot_skip_filter():
    """cutoff.rs:74-98: budget 0, filter id > 2 (external ids 3, 4 = internal docids 1, 0): candidates and hits are [0, 1]"""
    import numpy as np

    from meilisearch_b200.tokenizer import TokenBatch
    from oracle.pyoracle import OracleIndex

    o = OracleIndex(cutoff_image(), criteria=["words", "typo"])
    r = o.search_batch(TokenBatch(["hello puppy kefir"]), limit=100, stop_after=0, universes=[np.array([0b00011], np.uint64)])
    assert r.ids(0) == [0, 1] and int(r.n_candidates[0]) == 2 and r.degraded[0] == 1
