"""meilisearch_b200 — B200-native (sm_100a) implementation of milli's query-time scoring path.

Python host-side mirror of the reference interface for this path (crates/milli/src/search/mod.rs:58-86,280-415,526-535):
`Index` (the staged, HBM-resident copy of what `milli::Index` exposes to search), the `Search` builder with
`execute()` / `execute_hybrid()`, and `SearchResult`.  Everything goes through the C ABI of include/b200milli.h
(libb200milli.so, built in-tree by meilisearch_b200/csrc/build.sh).  There is no CPU fallback: without the CUDA
library or without a device, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .tokenizer import TokenBatch, tokenize  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200milli.so")
MAX_SCORES = 12
CRITERIA = {"words": 0, "typo": 1, "proximity": 2, "attribute": 3, "attributeRank": 4, "wordPosition": 5, "sort": 6, "exactness": 7}
DEFAULT_CRITERIA = ["words", "typo", "proximity", "attributeRank", "sort", "wordPosition", "exactness"]  # criterion.rs:121-131
TMS = {"last": 0, "all": 1, "frequency": 2}
SCORE_KINDS = ["words", "typo", "proximity", "fid", "position", "exactAttribute", "exactWords", "vector", "skipped"]
ERRORS = {-1: "NO_DEVICE", -2: "CUDA", -3: "INVALID", -4: "UNSUPPORTED", -5: "CAPACITY", -6: "STATE"}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200milli error {ERRORS.get(code, code)}: {msg}")
        self.code = code


class _Settings(C.Structure):
    _fields_ = [("n_fields", C.c_uint32), ("weights", C.c_void_p), ("criteria", C.c_void_p), ("n_criteria", C.c_uint32),
                ("authorize_typos", C.c_int32), ("min_word_len_one_typo", C.c_uint32), ("min_word_len_two_typos", C.c_uint32),
                ("prefix_search", C.c_int32), ("exact_words", C.c_char_p)]


class _Batch(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("token_begin", C.c_void_p), ("token_kind", C.c_void_p), ("lemma_off", C.c_void_p),
                ("lemma_bytes", C.c_void_p), ("terms_matching_strategy", C.c_int32), ("scoring_strategy", C.c_int32),
                ("offset", C.c_uint32), ("limit", C.c_uint32), ("words_limit", C.c_uint32), ("vectors", C.c_void_p),
                ("mode", C.c_int32), ("semantic_ratio", C.c_float), ("universes", C.c_void_p), ("n_universe_words", C.c_uint64),
                ("time_budget_ns", C.c_uint64), ("stop_after", C.c_int64), ("has_ranking_score_threshold", C.c_int32),
                ("ranking_score_threshold", C.c_double)]


class _Results(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("docids", "n_hits", "n_scores", "score_kind", "score_rank", "score_max", "score_sim",
                                          "n_candidates", "semantic_hits", "status", "degraded", "used_negative_operator", "candidates")] + \
               [("candidates_words", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("device_steps", C.c_uint64), ("posting_bytes", C.c_uint64), ("matrix_bytes", C.c_uint64),
                ("dictionary_bytes", C.c_uint64), ("vector_bytes", C.c_uint64), ("kernel_ms", C.c_double * 10),
                ("kernel_count", C.c_uint64 * 10), ("kernel_bytes", C.c_uint64 * 10), ("device_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("host_ms", C.c_double * 8),
                ("hbm_bytes_staged", C.c_uint64), ("deferred", C.c_uint64), ("arena_peak_bytes", C.c_uint64),
                ("eval_class_launches", C.c_uint64 * 9), ("eval_class_tiles", C.c_uint64 * 9)]


KERNELS = ["lev_match", "act_compact", "pair_probe", "scatter", "eval_paths", "emit", "vec_dist", "topk_select", "vec_gemm_topk", "vec_merge"]


def build_library(force=False):
    """Compile the CUDA extension in-tree for sm_100a (works without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cpp", ".h"))] + [os.path.join(_HERE, "..", "include", "b200milli.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["sh", os.path.join(csrc, "build.sh")])
    return LIB_PATH


_lib = None


def load_library():
    """Load libb200milli.so.  Fails loudly when the extension has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        l = C.CDLL(LIB_PATH)
        l.b200_open.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        l.b200_close.argtypes = [C.c_void_p]
        l.b200_last_error.restype = C.c_char_p
        l.b200_last_error.argtypes = [C.c_void_p]
        l.b200_open_error.restype = C.c_char_p
        l.b200_stage_dictionary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        l.b200_stage_db.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.b200_stage_documents_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        l.b200_stage_settings.argtypes = [C.c_void_p, C.POINTER(_Settings)]
        l.b200_stage_synonyms.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        l.b200_stage_finish.argtypes = [C.c_void_p]
        l.b200_stage_embeddings.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        l.b200_stage_embeddings_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        l.b200_stage_distribution.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        l.b200_derive_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 4
        l.b200_union_postings.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        l.b200_nns_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        l.b200_nns_batch_sharded.argtypes = l.b200_nns_batch.argtypes
        l.b200_comm_unique_id.argtypes = [C.c_void_p, C.c_void_p]
        l.b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        l.b200_search_batch.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Results)]
        l.b200_proximity_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        l.b200_graph_from_tokens.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(C.c_void_p)]
        l.b200_graph_free.argtypes = [C.c_void_p]
        l.b200_rule_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        l.b200_rule_next.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)]
        l.b200_rule_end.argtypes = [C.c_void_p]
        l.b200_get_stats.argtypes = [C.c_void_p, C.POINTER(_Stats)]
        l.b200_reset_stats.argtypes = [C.c_void_p]
        _lib = l
    return _lib


SYMBOLS = ["b200_open", "b200_close", "b200_last_error", "b200_open_error", "b200_stage_dictionary", "b200_stage_db",
           "b200_stage_documents_ids", "b200_stage_settings", "b200_stage_synonyms", "b200_stage_finish", "b200_stage_embeddings", "b200_stage_embeddings_f16", "b200_stage_distribution",
           "b200_derive_batch", "b200_union_postings", "b200_proximity_pairs", "b200_nns_batch", "b200_nns_batch_sharded", "b200_comm_unique_id", "b200_comm_init", "b200_search_batch", "b200_graph_from_tokens",
           "b200_graph_free", "b200_rule_start", "b200_rule_next", "b200_rule_end", "b200_get_stats", "b200_reset_stats"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class SearchResult:
    """milli::SearchResult (search/mod.rs:526-535) for a batch of queries."""

    def __init__(self, n, limit):
        self.limit = max(limit, 1)
        L = self.limit
        self.documents_ids = np.zeros((n, L), np.uint32)
        self.n_hits = np.zeros(n, np.uint32)
        self.n_scores = np.zeros((n, L), np.uint8)
        self.score_kind = np.zeros((n, L, MAX_SCORES), np.uint8)
        self.score_rank = np.zeros((n, L, MAX_SCORES), np.uint32)
        self.score_max = np.zeros((n, L, MAX_SCORES), np.uint32)
        self.score_sim = np.zeros((n, L, MAX_SCORES), np.float32)
        self.n_candidates = np.zeros(n, np.uint64)
        self.semantic_hit_count = np.zeros(n, np.uint32)
        self.status = np.zeros(n, np.int32)
        self.degraded = np.zeros(n, np.uint8)
        self.used_negative_operator = np.zeros(n, np.uint8)
        self.candidates = None  # (n, words) uint64 when requested with Search.with_candidates()

    def ids(self, q):
        return [int(x) for x in self.documents_ids[q, : self.n_hits[q]]]

    def scores(self, q):
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


class Index:
    """The staged index: what milli reads from LMDB at query time, resident in HBM."""

    def __init__(self, image=None, *, device=0, criteria=None, authorize_typos=True, one_typo=5, two_typos=9, prefix_search=True,
                 weights=None, exact_words=(), synonyms=None):
        self._l = load_library()
        h = C.c_void_p()
        rc = self._l.b200_open(device, C.byref(h))
        if rc != 0:
            raise B200Error(rc, self._l.b200_open_error().decode())
        self._h = h
        self.dim = 0
        if image is not None:
            self.stage(image, criteria=criteria, authorize_typos=authorize_typos, one_typo=one_typo, two_typos=two_typos,
                       prefix_search=prefix_search, weights=weights, exact_words=exact_words, synonyms=synonyms)

    def _ck(self, rc):
        if rc != 0:
            raise B200Error(rc, self._l.b200_last_error(self._h).decode())

    def stage(self, image, *, criteria=None, authorize_typos=True, one_typo=5, two_typos=9, prefix_search=True, weights=None, exact_words=(),
              synonyms=None):
        """image: anything with dict_bytes/dict_offsets/n_words, dbs[i].{key_bytes,key_offsets,val_bytes,val_offsets,n_keys},
        documents_ids_cbo, n_fields — i.e. the LMDB databases in their on-disk formats."""
        l = self._l
        self._n_docs = int(image.n_docs)
        self._ck(l.b200_stage_dictionary(self._h, _p(image.dict_bytes), _p(image.dict_offsets), image.n_words))
        for i, db in enumerate(image.dbs):
            self._ck(l.b200_stage_db(self._h, i, db.n_keys, _p(db.key_bytes), _p(db.key_offsets), _p(db.val_bytes), _p(db.val_offsets)))
        self._ck(l.b200_stage_documents_ids(self._h, _p(image.documents_ids_cbo), len(image.documents_ids_cbo)))
        w = np.asarray(weights if weights is not None else list(range(image.n_fields)), np.uint16)
        c = np.asarray([CRITERIA[x] for x in (DEFAULT_CRITERIA if criteria is None else criteria)], np.int32)
        s = _Settings(image.n_fields, _p(w), _p(c), len(c), int(authorize_typos), one_typo, two_typos, int(prefix_search),
                      "\n".join(exact_words).encode() if exact_words else None)
        self._ck(l.b200_stage_settings(self._h, C.byref(s)))
        if synonyms:
            pairs = [(k, v) for k, vs in synonyms.items() for v in vs]
            fr = (C.c_char_p * len(pairs))(*[k.encode() for k, _ in pairs])
            to = (C.c_char_p * len(pairs))(*[v.encode() for _, v in pairs])
            self._ck(l.b200_stage_synonyms(self._h, len(pairs), fr, to))
        self._ck(l.b200_stage_finish(self._h))
        self.n_fields = image.n_fields

    def set_embeddings(self, matrix, docids=None, distribution=None):
        """f32 rows (converted to fp16 on the device), or a float16 matrix staged as it is."""
        ids = None if docids is None else np.ascontiguousarray(docids, np.uint32)
        if getattr(matrix, "dtype", None) == np.float16:
            m = np.ascontiguousarray(matrix)
            self._ck(self._l.b200_stage_embeddings_f16(self._h, _p(m), m.shape[0], m.shape[1], _p(ids)))
        else:
            m = np.ascontiguousarray(matrix, np.float32)
            self._ck(self._l.b200_stage_embeddings(self._h, _p(m), m.shape[0], m.shape[1], _p(ids)))
        self.dim = m.shape[1]
        if distribution:
            self._ck(self._l.b200_stage_distribution(self._h, 1, distribution[0], distribution[1]))

    # S3 — compute_fully_if_needed (compute_derivations.rs:21-37)
    def derive(self, words, max_typo, is_prefix):
        enc = [w.encode() for w in words]
        off = np.zeros(len(enc) + 1, np.uint32)
        off[1:] = np.cumsum([len(b) for b in enc])
        buf = np.frombuffer(b"".join(enc) + b"\0", np.uint8).copy()
        mt = np.asarray(max_typo, np.uint8)
        ip = np.asarray(is_prefix, np.uint8)
        n = len(enc)
        one = np.zeros((n, 150), np.uint32)
        two = np.zeros((n, 50), np.uint32)
        n1 = np.zeros(n, np.uint32)
        n2 = np.zeros(n, np.uint32)
        self._ck(self._l.b200_derive_batch(self._h, n, _p(buf), _p(off), _p(mt), _p(ip), _p(one), _p(n1), _p(two), _p(n2)))
        return [(one[i, : n1[i]].copy(), two[i, : n2[i]].copy()) for i in range(n)]

    # S4 — VectorStore::nns_by_vector (vector/store.rs:638-675)
    def union_postings(self, db, key_indices, universe=None):
        """S2: (OR of the posting lists of database `db` at the given key positions) AND universe, as dense u64 words."""
        keys = np.ascontiguousarray(key_indices, np.uint32)
        n_words = (self._n_docs + 63) // 64
        out = np.zeros(n_words, np.uint64)
        uni = None if universe is None else np.ascontiguousarray(universe, np.uint64)
        self._ck(self._l.b200_union_postings(self._h, int(db), _p(keys), len(keys), _p(uni), 0 if uni is None else len(uni), _p(out)))
        return out

    def proximity_pairs(self, left, right, fwd_prox, bwd_prox, universe=None):
        """S2 for proximity conditions: universe AND the union of word_pair_proximity_docids[(fwd, l, r)] and [(bwd, r, l)]"""
        lw, rw = np.ascontiguousarray(left, np.uint32), np.ascontiguousarray(right, np.uint32)
        n_words = (self._n_docs + 63) // 64
        out = np.zeros(n_words, np.uint64)
        uni = None if universe is None else np.ascontiguousarray(universe, np.uint64)
        self._ck(self._l.b200_proximity_pairs(self._h, _p(lw), len(lw), _p(rw), len(rw), int(fwd_prox), int(bwd_prox), _p(uni),
                                              0 if uni is None else len(uni), _p(out)))
        return out

    def query_graph(self, query, stop_words=frozenset(), terms_matching_strategy="last", words_limit=10):
        """S1: QueryGraph::from_query for one query, as an opaque QueryGraph handle"""
        tokens = query if isinstance(query, TokenBatch) else TokenBatch([query], stop_words)
        b = _Batch(1, _p(tokens.token_begin), _p(tokens.token_kind), _p(tokens.lemma_off), _p(tokens.lemma_bytes), TMS[terms_matching_strategy], 0, 0, 1,
                   words_limit, None, 0, 0.0)
        b.stop_after = -1
        g = C.c_void_p()
        self._ck(self._l.b200_graph_from_tokens(self._h, C.byref(b), C.byref(g)))
        return QueryGraph(self, g)

    def nns_by_vector(self, queries, limit, candidates=None):
        q = np.ascontiguousarray(np.atleast_2d(queries), np.float32)
        n = q.shape[0]
        ids = np.zeros((n, limit), np.uint32)
        dist = np.zeros((n, limit), np.float32)
        cnt = np.zeros(n, np.uint32)
        cw = None if candidates is None else np.ascontiguousarray(candidates, np.uint64)
        self._ck(self._l.b200_nns_batch(self._h, _p(q), n, q.shape[1], limit, _p(cw), 0 if cw is None else len(cw), _p(ids), _p(dist), _p(cnt)))
        return ids, dist, cnt

    def comm_unique_id(self):
        out = np.zeros(128, np.uint8)
        self._ck(self._l.b200_comm_unique_id(self._h, _p(out)))
        return out

    def comm_init(self, rank, world, unique_id):
        """join the NCCL communicator of a corpus partitioned across GPUs (unique_id: 128 bytes drawn by rank 0)"""
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._ck(self._l.b200_comm_init(self._h, int(rank), int(world), _p(uid)))

    def nns_by_vector_sharded(self, queries, limit, candidates=None):
        """every rank: the same queries; returns the merged global top-k (per-shard scan + ncclAllGather + device merge)"""
        q = np.ascontiguousarray(np.atleast_2d(queries), np.float32)
        n = q.shape[0]
        ids = np.zeros((n, limit), np.uint32)
        dist = np.zeros((n, limit), np.float32)
        cnt = np.zeros(n, np.uint32)
        cw = None if candidates is None else np.ascontiguousarray(candidates, np.uint64)
        self._ck(self._l.b200_nns_batch_sharded(self._h, _p(q), n, q.shape[1], limit, _p(cw), 0 if cw is None else len(cw), _p(ids), _p(dist), _p(cnt)))
        return ids, dist, cnt

    def search(self):
        return Search(self)

    def stats(self):
        s = _Stats()
        self._l.b200_get_stats(self._h, C.byref(s))
        d = {n: getattr(s, n) for n, _ in _Stats._fields_ if not n.startswith("kernel_") and not n.startswith("eval_class") and n != "host_ms"}
        d["eval_class_launches"] = list(s.eval_class_launches)
        d["eval_class_tiles"] = list(s.eval_class_tiles)
        d["host_ms"] = dict(zip(["parse", "derive", "terms", "pack", "device_wait", "advance", "results", "total"], list(s.host_ms)))
        d["kernel_launches"] = s.kernel_launches
        d["kernels"] = {KERNELS[i]: {"ms": s.kernel_ms[i], "count": s.kernel_count[i], "bytes": s.kernel_bytes[i]} for i in range(len(KERNELS))}
        return d

    def reset_stats(self):
        self._l.b200_reset_stats(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._l.b200_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QueryGraph:
    """S1: an opaque query graph of the library (QueryGraph + its terms).  `rule()` is RankingRule::start_iteration .. next_bucket ..
    end_iteration for one ranking rule (ranking_rules.rs:26-83)."""

    def __init__(self, index, handle):
        self.index, self._g = index, handle

    def rule(self, kind, universe=None, terms_matching_strategy="last"):
        """yields (candidates bitmap, rank, max_rank, child QueryGraph or None) per bucket in ascending cost order"""
        ix = self.index
        uni = None if universe is None else np.ascontiguousarray(universe, np.uint64)
        r = C.c_void_p()
        ix._ck(ix._l.b200_rule_start(ix._h, SCORE_KINDS.index(kind), TMS[terms_matching_strategy], self._g, _p(uni), 0 if uni is None else len(uni), C.byref(r)))
        try:
            n_words = (ix._n_docs + 63) // 64
            while True:
                out = np.zeros(n_words, np.uint64)
                rank, mx, child = C.c_uint32(), C.c_uint32(), C.c_void_p()
                rc = ix._l.b200_rule_next(r, None, _p(out), n_words, C.byref(rank), C.byref(mx), C.byref(child))
                if rc == 1:
                    return
                if rc != 0:
                    raise B200Error(rc, "b200_rule_next")
                yield out, rank.value, mx.value, (QueryGraph(ix, child) if child.value else None)
        finally:
            ix._l.b200_rule_end(r)

    def close(self):
        if self._g is not None and self._g.value:
            self.index._l.b200_graph_free(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Search:
    """milli::Search builder (search/mod.rs:58-278) for a *batch* of queries sharing the same parameters."""

    def __init__(self, index):
        self.index = index
        self._tokens = None
        self._vectors = None
        self._tms = "last"
        self._scoring = "skip"
        self._offset, self._limit, self._words_limit = 0, 20, 10
        self._universes, self._budget_ms, self._stop_after, self._threshold, self._want_candidates = None, None, None, None, False

    def query(self, queries, stop_words=frozenset()):
        self._tokens = queries if isinstance(queries, TokenBatch) else TokenBatch([queries] if isinstance(queries, str) else list(queries), stop_words)
        return self

    def semantic(self, vectors):
        self._vectors = np.ascontiguousarray(np.atleast_2d(vectors), np.float32)
        return self

    def terms_matching_strategy(self, s):
        self._tms = s
        return self

    def scoring_strategy(self, s):
        self._scoring = s
        return self

    def offset(self, n):
        self._offset = n
        return self

    def limit(self, n):
        self._limit = n
        return self

    def words_limit(self, n):
        self._words_limit = n
        return self

    def universes(self, bitmaps):
        """filtered_universe per query: a list of uint64 word arrays (None = all documents), or one array shared by the batch"""
        self._universes = bitmaps
        return self

    def deadline(self, budget_ms=None, stop_after=None):
        """Search::deadline: a time budget in ms, or the reference's poll-count hook Deadline::never().with_stop_after(n)"""
        self._budget_ms, self._stop_after = budget_ms, stop_after
        return self

    def ranking_score_threshold(self, t):
        self._threshold = t
        return self

    def with_candidates(self):
        self._want_candidates = True
        return self

    def _run(self, mode, ratio=0.0):
        ix = self.index
        tokens = self._tokens
        if tokens is None:
            n = self._vectors.shape[0] if self._vectors is not None else 1
            tokens = TokenBatch([""] * n)
        n = tokens.n_queries
        res = SearchResult(n, self._limit)
        b = _Batch(n, _p(tokens.token_begin), _p(tokens.token_kind), _p(tokens.lemma_off), _p(tokens.lemma_bytes), TMS[self._tms],
                   1 if self._scoring == "detailed" else 0, self._offset, self._limit, self._words_limit,
                   _p(self._vectors) if self._vectors is not None else None, mode, ratio)
        keep = []
        if self._universes is not None:
            us = self._universes
            if isinstance(us, np.ndarray) and us.ndim == 1:
                us = [us] * n
            ptrs = (C.c_void_p * n)()
            cache = {}
            for i, u in enumerate(us):
                if u is None:
                    continue
                if id(u) not in cache:
                    a = np.ascontiguousarray(u, np.uint64)
                    keep.append(a)
                    cache[id(u)] = a
                a = cache[id(u)]
                ptrs[i] = a.ctypes.data
                b.n_universe_words = len(a)
            keep.append(ptrs)
            b.universes = C.cast(ptrs, C.c_void_p)
        b.time_budget_ns = 0 if self._budget_ms is None else max(1, int(self._budget_ms * 1e6))
        b.stop_after = -1 if self._stop_after is None else int(self._stop_after)
        b.has_ranking_score_threshold = int(self._threshold is not None)
        b.ranking_score_threshold = float(self._threshold or 0.0)
        r = _Results(_p(res.documents_ids), _p(res.n_hits), _p(res.n_scores), _p(res.score_kind), _p(res.score_rank), _p(res.score_max),
                     _p(res.score_sim), _p(res.n_candidates), _p(res.semantic_hit_count), _p(res.status), _p(res.degraded),
                     _p(res.used_negative_operator), None, 0)
        if self._want_candidates:
            words = (ix._n_docs + 63) // 64
            res.candidates = np.zeros((n, words), np.uint64)
            r.candidates, r.candidates_words = _p(res.candidates), words
        ix._ck(ix._l.b200_search_batch(ix._h, C.byref(b), C.byref(r)))
        return res

    def execute(self):
        """Search::execute (search/mod.rs:280): keyword search, or semantic search when a vector was given and no query."""
        return self._run(1 if (self._vectors is not None and self._tokens is None) else 0)

    def execute_hybrid(self, semantic_ratio):
        """Search::execute_hybrid (search/hybrid.rs:264)."""
        return self._run(2, float(semantic_ratio))
