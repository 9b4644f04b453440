"""SURVEY §8(d) cfg 1 ("movies, exact single-term, CPU"): 31 944 documents, two searchable fields (title 2-8 words, overview 20-80
words), Zipf(1.07) over a 60 k-word vocabulary; 1024 single in-vocabulary words, authorizeTypos = false, limit 20.  This
configuration is the reference's own CPU-runnable case: it exercises the ranking-rule plumbing of the CPU restatement only (no
GPU, result = oracle by construction) and reports its throughput.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from corpus.pyindexgen import IndexImage
from meilisearch_b200.tokenizer import TokenBatch
from oracle.pyoracle import OracleIndex

img = IndexImage(2)
img.add_synthetic(31944, 60000, zipf_s=1.07, len_lo=2, len_hi=8, seed=0xB200)
img.build()
rng = np.random.default_rng(1)
words = [img.word(int(i)) for i in rng.integers(0, img.n_words, 1024)]
tb = TokenBatch(words)
o = OracleIndex(img, authorize_typos=False)
out = {}
for threads in (1, os.cpu_count() or 1):
    o.search_batch(tb, n_threads=threads)
    t0 = time.perf_counter()
    r = o.search_batch(tb, n_threads=threads)
    dt = time.perf_counter() - t0
    out[f"threads_{threads}"] = {"queries_per_s": 1024 / dt, "p50_ms": 1e3 * float(np.median(r.seconds)), "p95_ms": 1e3 * float(np.percentile(r.seconds, 95))}
hits = int((r.n_hits > 0).sum())
print(json.dumps({"config": "cfg1: 31944 docs x (title 2-8 words, overview 20-80 words), 60k-word Zipf vocabulary, 1024 exact single-word queries, authorizeTypos=false, limit 20",
                  "impl": "CPU restatement of milli (oracle), no GPU", "docs": int(img.n_docs), "dictionary": int(img.n_words), "queries_with_hits": hits,
                  "host_threads": os.cpu_count(), **out}))
