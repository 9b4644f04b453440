"""Deterministic keyword workload for ncu and for the B200_WORK_HIST statistics: the SAME batch of 1024 queries BATCHES times in
single-lane mode, so that launch i of a kernel is the same work in every batch.  DOCS / VOCAB pick the corpus (default cfg 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_SINGLE_LANE"] = "1"
import meilisearch_b200 as mb
from corpus.pyindexgen import synthetic_image
from meilisearch_b200.tokenizer import TokenBatch

img = synthetic_image(int(os.environ.get("DOCS", "1000000")), int(os.environ.get("VOCAB", "400000")), seed=0xB200)
ix = mb.Index(img)
tb = TokenBatch(img.synthetic_queries(1024, seed=0))
scoring = os.environ.get("SCORING", "skip")
for i in range(int(os.environ.get("BATCHES", "3"))):
    ix.reset_stats()
    r = ix.search().query(tb).scoring_strategy(scoring).execute()
    st = ix.stats()
    print("batch", i, "steps", st["device_steps"], "device_ms", round(st["device_ms"], 2), {k: (v["count"], round(v["ms"], 2)) for k, v in st["kernels"].items() if v["count"]}, flush=True)
