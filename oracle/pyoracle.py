"""ctypes wrapper over oracle/libmilli_oracle.so — TEST INFRASTRUCTURE ONLY.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CRITERIA = {"words": 0, "typo": 1, "proximity": 2, "attribute": 3, "attributeRank": 4, "wordPosition": 5, "sort": 6, "exactness": 7}
DEFAULT_CRITERIA = ["words", "typo", "proximity", "attributeRank", "sort", "wordPosition", "exactness"]
TMS = {"last": 0, "all": 1, "frequency": 2}
SCORE_KINDS = ["words", "typo", "proximity", "fid", "position", "exactAttribute", "exactWords", "vector", "skipped"]
MAX_SCORES = 12


class _Batch(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("token_begin", C.c_void_p), ("token_kind", C.c_void_p), ("lemma_off", C.c_void_p),
                ("lemma_bytes", C.c_void_p), ("tms", C.c_int), ("scoring", C.c_int), ("offset", C.c_uint32), ("limit", C.c_uint32),
                ("words_limit", C.c_uint32), ("vectors", C.c_void_p), ("semantic_ratio", C.c_float), ("hybrid", C.c_int),
                ("vector_only", C.c_int), ("has_threshold", C.c_int), ("threshold", C.c_double),
                ("universes", C.c_void_p), ("n_universe_words", C.c_uint64), ("stop_after", C.c_int64), ("time_budget_ns", C.c_uint64)]


class _Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("docids", "n_hits", "score_kind", "score_rank", "score_max", "score_sim", "n_scores",
                                          "n_candidates", "semantic_hits", "fetch_bytes", "seconds", "degraded", "used_negative_operator")]


def build_lib():
    so = os.path.join(_HERE, "libmilli_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build_lib())
        l.orc_index_new.restype = C.c_void_p
        l.orc_index_free.argtypes = [C.c_void_p]
        l.orc_set_dictionary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        l.orc_set_db.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_set_documents_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        l.orc_set_settings.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
        l.orc_add_exact_word.argtypes = [C.c_void_p, C.c_char_p]
        l.orc_clear_exact_words.argtypes = [C.c_void_p]
        l.orc_clear_synonyms.argtypes = [C.c_void_p]
        l.orc_add_synonym.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        l.orc_set_embeddings.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        l.orc_set_embeddings_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32]
        l.orc_set_distribution.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        l.orc_search_batch.restype = C.c_int
        l.orc_search_batch.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Out), C.c_uint32, C.c_char_p, C.c_uint32]
        l.orc_derive.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_nns.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_cbo_decode.restype = C.c_uint64
        l.orc_cbo_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        l.orc_cbo_len.restype = C.c_uint64
        l.orc_cbo_len.argtypes = [C.c_void_p, C.c_uint64]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleResults:
    def __init__(self, n, limit):
        self.docids = np.zeros((n, limit), np.uint32)
        self.n_hits = np.zeros(n, np.uint32)
        self.score_kind = np.zeros((n, limit, MAX_SCORES), np.uint8)
        self.score_rank = np.zeros((n, limit, MAX_SCORES), np.uint32)
        self.score_max = np.zeros((n, limit, MAX_SCORES), np.uint32)
        self.score_sim = np.zeros((n, limit, MAX_SCORES), np.float32)
        self.n_scores = np.zeros((n, limit), np.uint8)
        self.n_candidates = np.zeros(n, np.uint64)
        self.semantic_hits = np.zeros(n, np.uint32)
        self.fetch_bytes = np.zeros(n, np.uint64)
        self.seconds = np.zeros(n, np.float64)
        self.degraded = np.zeros(n, np.uint8)
        self.used_negative_operator = np.zeros(n, np.uint8)

    def ids(self, q):
        return [int(x) for x in self.docids[q, : self.n_hits[q]]]

    def scores(self, q):
        """[[(kind, rank, max_rank) | ('vector', similarity)]] per hit"""
        out = []
        for i in range(int(self.n_hits[q])):
            row = []
            for s in range(int(self.n_scores[q, i])):
                k = SCORE_KINDS[self.score_kind[q, i, s]]
                if k == "vector":
                    sim = float(self.score_sim[q, i, s])
                    row.append(("vector", None if sim < 0 else sim))
                else:
                    row.append((k, int(self.score_rank[q, i, s]), int(self.score_max[q, i, s])))
            out.append(row)
        return out


class OracleIndex:
    def __init__(self, image, criteria=None, authorize_typos=True, one_typo=5, two_typos=9, prefix_search=True, weights=None):
        self._l = lib()
        self._h = self._l.orc_index_new()
        self.image = image
        l = self._l
        l.orc_set_dictionary(self._h, _p(image.dict_bytes), _p(image.dict_offsets), image.n_words)
        for i, db in enumerate(image.dbs):
            l.orc_set_db(self._h, i, db.n_keys, _p(db.key_bytes), _p(db.key_offsets), _p(db.val_bytes), _p(db.val_offsets))
        l.orc_set_documents_ids(self._h, _p(image.documents_ids_cbo), len(image.documents_ids_cbo))
        self.n_fields = image.n_fields
        self.weights = list(weights) if weights is not None else list(range(image.n_fields))
        self._settings = dict(criteria=list(DEFAULT_CRITERIA if criteria is None else criteria), authorize_typos=authorize_typos, one_typo=one_typo,
                              two_typos=two_typos, prefix_search=prefix_search)
        self._push_settings()
        self.dim = 0

    def _push_settings(self):
        s = self._settings
        w = np.asarray(self.weights, np.uint16)
        c = np.asarray([CRITERIA[x] for x in s["criteria"]], np.int32)
        self._l.orc_set_settings(self._h, self.n_fields, _p(w), _p(c), len(c), int(s["authorize_typos"]), s["one_typo"],
                                 s["two_typos"], int(s["prefix_search"]))

    def update_settings(self, **kw):
        exact_words = kw.pop("exact_words", None)
        synonyms = kw.pop("synonyms", None)
        self._settings.update(kw)
        self._push_settings()
        if exact_words is not None:
            self._l.orc_clear_exact_words(self._h)
            for w in exact_words:
                self._l.orc_add_exact_word(self._h, w.encode())
        if synonyms is not None:
            self._l.orc_clear_synonyms(self._h)
            for k, vs in synonyms.items():
                for v in vs:
                    self._l.orc_add_synonym(self._h, k.encode(), v.encode())

    def set_embeddings(self, matrix, docids=None, distribution=None):
        ids = np.arange(matrix.shape[0], dtype=np.uint32) if docids is None else np.ascontiguousarray(docids, np.uint32)
        if getattr(matrix, "dtype", None) == np.float16:
            m = np.ascontiguousarray(matrix)
            self._l.orc_set_embeddings_f16(self._h, _p(m), m.shape[0], m.shape[1], _p(ids), os.cpu_count() or 1)
        else:
            m = np.ascontiguousarray(matrix, np.float32)
            self._l.orc_set_embeddings(self._h, _p(m), m.shape[0], m.shape[1], _p(ids))
        self.dim = m.shape[1]
        if distribution:
            self._l.orc_set_distribution(self._h, 1, distribution[0], distribution[1])

    def search_batch(self, tokens, *, tms="last", scoring="skip", offset=0, limit=20, words_limit=10, vectors=None, hybrid=False,
                     semantic_ratio=0.5, vector_only=False, threshold=None, universes=None, stop_after=None, time_budget_ms=None, n_threads=1):
        """tokens: meilisearch_b200.tokenizer.TokenBatch"""
        n = tokens.n_queries
        res = OracleResults(n, limit)
        b = _Batch()
        b.n_queries = n
        b.token_begin, b.token_kind, b.lemma_off, b.lemma_bytes = _p(tokens.token_begin), _p(tokens.token_kind), _p(tokens.lemma_off), _p(tokens.lemma_bytes)
        b.tms, b.scoring = TMS[tms], 1 if scoring == "detailed" else 0
        b.offset, b.limit, b.words_limit = offset, limit, words_limit
        vec = None
        if vectors is not None:
            vec = np.ascontiguousarray(vectors, np.float32)
            b.vectors = _p(vec)
        b.semantic_ratio, b.hybrid, b.vector_only = semantic_ratio, int(hybrid), int(vector_only)
        b.has_threshold, b.threshold = int(threshold is not None), float(threshold or 0.0)
        b.stop_after = -1 if stop_after is None else int(stop_after)
        b.time_budget_ns = 0 if time_budget_ms is None else max(1, int(time_budget_ms * 1e6))
        keep = []
        if universes is not None:  # list of per-query uint64 word arrays (None = all documents); equal objects are shared
            ptrs = (C.c_void_p * n)()
            for i, u in enumerate(universes):
                if u is not None:
                    a = np.ascontiguousarray(u, np.uint64)
                    keep.append(a)
                    ptrs[i] = a.ctypes.data
                    b.n_universe_words = len(a)
            keep.append(ptrs)
            b.universes = C.cast(ptrs, C.c_void_p)
        o = _Out()
        for name, _ in _Out._fields_:
            setattr(o, name, _p(getattr(res, name)))
        err = C.create_string_buffer(512)
        rc = self._l.orc_search_batch(self._h, C.byref(b), C.byref(o), n_threads, err, 512)
        if rc != 0:
            raise RuntimeError("oracle: " + err.value.decode())
        return res

    def derive(self, word, max_typo, is_prefix):
        one = np.zeros(256, np.uint32)
        two = np.zeros(256, np.uint32)
        n1, n2 = C.c_uint32(), C.c_uint32()
        self._l.orc_derive(self._h, word.encode(), max_typo, int(is_prefix), _p(one), C.byref(n1), _p(two), C.byref(n2))
        return one[: n1.value].copy(), two[: n2.value].copy()

    def nns(self, q, limit, cand_words=None):
        q = np.ascontiguousarray(q, np.float32)
        ids = np.zeros(limit, np.uint32)
        dist = np.zeros(limit, np.float32)
        n = C.c_uint32()
        cw = None if cand_words is None else np.ascontiguousarray(cand_words, np.uint64)
        self._l.orc_nns(self._h, _p(q), limit, _p(cw), 0 if cw is None else len(cw), _p(ids), _p(dist), C.byref(n))
        return ids[: n.value].copy(), dist[: n.value].copy()

    def __del__(self):
        try:
            self._l.orc_index_free(self._h)
        except Exception:
            pass


def cbo_decode(b: bytes):
    a = np.frombuffer(b, np.uint8).copy() if len(b) else np.zeros(1, np.uint8)
    n = lib().orc_cbo_len(_p(a), len(b))
    out = np.zeros(max(int(n), 1), np.uint32)
    k = lib().orc_cbo_decode(_p(a), len(b), _p(out), len(out))
    return out[: int(k)]
