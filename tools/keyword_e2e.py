"""End-to-end keyword batches (Detailed and Skip) at cfg 3 under lane / thread variants (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meilisearch_b200 as mb
from corpus.pyindexgen import synthetic_image
from meilisearch_b200.tokenizer import TokenBatch

docs, vocab = int(os.environ.get("DOCS", "10000000")), int(os.environ.get("VOCAB", "1500000"))
img = synthetic_image(docs, vocab, seed=0xB200)
batches = [TokenBatch(img.synthetic_queries(1024, seed=i)) for i in range(4)]
os.environ["B200_KERNEL_TIMERS"] = "0"
for cfg in os.environ.get("CFGS", "4:1:32,4:2:32,4:1:64,4:2:64,2:2:32,3:1:48").split(","):
    d, l, t = cfg.split(":")
    os.environ["B200_DRIVERS"], os.environ["B200_LANES_PER_DRIVER"], os.environ["B200_HOST_THREADS"] = d, l, t
    ix = mb.Index(img)   # the worker pools are sized when the handle first searches
    for scoring in ("detailed", "skip"):
        for w in range(2):
            ix.search().query(batches[w]).scoring_strategy(scoring).execute()
        ix.reset_stats()
        t0 = time.perf_counter()
        for i in range(4):
            ix.search().query(batches[i % 4]).scoring_strategy(scoring).execute()
        ms = 1e3 * (time.perf_counter() - t0) / 4
        st = ix.stats()
        print(f"drivers {d} lanes/driver {l} threads {t} {scoring}: {ms:.1f} ms/batch, steps {st['device_steps'] / 4:.0f}, host " +
              ", ".join(f"{k} {v / 4:.1f}" for k, v in st["host_ms"].items() if k in ("derive", "pack", "device_wait", "advance")), flush=True)
    ix.close()
