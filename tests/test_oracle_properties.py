"""Size-independent properties of the CPU oracle (the checker itself must be trustworthy beyond the golden vectors):
pagination consistency, strategy containment, scoring-strategy independence of the order, derivation bounds and vector order."""
import numpy as np
import pytest

from meilisearch_b200.tokenizer import TokenBatch
from oracle.pyoracle import OracleIndex
from tests.helpers import synthetic_image


@pytest.fixture(scope="module")
def img():
    return synthetic_image(30000, 9000, seed=0xA11)


@pytest.fixture(scope="module")
def queries(img):
    return img.synthetic_queries(60, seed=5)


def osa(a: str, b: str) -> int:
    """restricted Damerau-Levenshtein (optimal string alignment), the published semantics of levenshtein_automata with transpositions"""
    d = [[0] * (len(b) + 1) for _ in range(len(a) + 1)]
    for i in range(len(a) + 1):
        d[i][0] = i
    for j in range(len(b) + 1):
        d[0][j] = j
    for i in range(1, len(a) + 1):
        for j in range(1, len(b) + 1):
            d[i][j] = min(d[i - 1][j] + 1, d[i][j - 1] + 1, d[i - 1][j - 1] + (a[i - 1] != b[j - 1]))
            if i > 1 and j > 1 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                d[i][j] = min(d[i][j], d[i - 2][j - 2] + 1)
    return d[len(a)][len(b)]


def test_pagination_is_a_slice(img, queries):
    o = OracleIndex(img)
    tb = TokenBatch(queries)
    full = o.search_batch(tb, limit=64)
    for off, lim in ((0, 7), (5, 10), (13, 20), (39, 5)):
        page = o.search_batch(tb, offset=off, limit=lim)
        for q in range(len(queries)):
            assert page.ids(q) == full.ids(q)[off: off + lim], (queries[q], off, lim)


def test_scoring_strategy_does_not_change_the_order(img, queries):
    o = OracleIndex(img)
    tb = TokenBatch(queries)
    a, b = o.search_batch(tb, scoring="skip"), o.search_batch(tb, scoring="detailed")
    for q in range(len(queries)):
        assert a.ids(q) == b.ids(q), queries[q]


def test_all_is_contained_in_last_and_frequency(img, queries):
    o = OracleIndex(img)
    tb = TokenBatch(queries)
    res = {t: o.search_batch(tb, tms=t, limit=1000) for t in ("all", "last", "frequency")}
    for q in range(len(queries)):
        s_all = set(res["all"].ids(q))
        assert s_all <= set(res["last"].ids(q)), queries[q]
        assert s_all <= set(res["frequency"].ids(q)), queries[q]
        # the documents matching every term come first under either optional-words strategy
        n = len(res["all"].ids(q))
        assert set(res["last"].ids(q)[:n]) == s_all or n == 1000


def test_derivations_respect_distance_caps_and_order(img):
    o = OracleIndex(img)
    words = [img.word(i) for i in range(0, img.n_words, max(1, img.n_words // 40))]
    for w in words:
        if len(w) < 5:
            continue
        max_typo = 1 if len(w) < 9 else 2
        one, two = o.derive(w, max_typo, False)
        assert len(one) <= 150 and len(two) <= 50
        assert list(one) == sorted(set(one.tolist())) and list(two) == sorted(set(two.tolist()))
        for r in one:
            x = img.word(int(r))
            assert osa(w, x) == 1 and x[0] == w[0], (w, x)   # a typo on the first letter counts double (typo.rs:1-19)
        for r in two:
            x = img.word(int(r))
            assert osa(w, x) == 2 or (osa(w, x) == 1 and x[0] != w[0]), (w, x)
        if max_typo == 1:
            assert len(two) == 0


def test_vector_hits_are_ordered_by_distance_then_docid(img):
    rng = np.random.default_rng(3)
    n, d = 2000, 48
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb[10] = emb[700]
    o = OracleIndex(img)
    o.set_embeddings(emb, rng.permutation(n).astype(np.uint32))
    for i in range(6):
        q = emb[10] if i == 0 else rng.standard_normal(d).astype(np.float32)
        ids, dist = o.nns(q, 50)
        assert len(ids) == 50 and (np.diff(dist) >= 0).all()
        for a in range(49):
            if dist[a] == dist[a + 1]:
                assert ids[a] < ids[a + 1]
