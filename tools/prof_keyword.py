"""Deterministic keyword workload for ncu: cfg2 corpus, the SAME batch of 1024 queries three times in single-lane mode, so that
launch i of a kernel is the same work in every batch (tools/profile.sh uses -s to pick the heavy launches of the third batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_SINGLE_LANE"] = "1"
import meilisearch_b200 as mb
from corpus.pyindexgen import IndexImage
from meilisearch_b200.tokenizer import TokenBatch

img = IndexImage(1)
img.add_synthetic(int(os.environ.get("DOCS", "1000000")), int(os.environ.get("VOCAB", "400000")), seed=0xB200)
img.build()
ix = mb.Index(img)
tb = TokenBatch(img.synthetic_queries(1024, seed=0))
for i in range(int(os.environ.get("BATCHES", "3"))):
    ix.reset_stats()
    r = ix.search().query(tb).execute()
    st = ix.stats()
    print("batch", i, "steps", st["device_steps"], {k: (v["count"], round(v["ms"], 2)) for k, v in st["kernels"].items() if v["count"]}, flush=True)
