"""End-to-end hybrid batches at cfg 3 under a few scheduling variants (developer tool): prints ms per 1024-query batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import meilisearch_b200 as mb
from corpus.pyindexgen import synthetic_embeddings_f16, synthetic_image
from meilisearch_b200.tokenizer import TokenBatch

docs, vocab = int(os.environ.get("DOCS", "10000000")), int(os.environ.get("VOCAB", "1500000"))
img = synthetic_image(docs, vocab, seed=0xB200)
ix = mb.Index(img)
ix.set_embeddings(synthetic_embeddings_f16(int(img.n_docs), 768, seed=0xE5BED))
batches = [TokenBatch(img.synthetic_queries(1024, seed=i)) for i in range(4)]
vecs = [np.random.default_rng(77 + i).standard_normal((1024, 768), dtype=np.float32) for i in range(4)]
os.environ["B200_KERNEL_TIMERS"] = "0"


def run(label, reps=6):
    for w in range(2):
        ix.search().query(batches[w]).semantic(vecs[w]).execute_hybrid(0.5)
    t0 = time.perf_counter()
    for i in range(reps):
        ix.search().query(batches[i % 4]).semantic(vecs[i % 4]).execute_hybrid(0.5)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(f"{label}: {ms:.1f} ms/batch = {1024e3 / ms:.0f} q/s", flush=True)


for start in os.environ.get("STARTS", "0,1,2").split(","):
    os.environ["B200_VEC_START"] = start
    os.environ.pop("B200_VEC_SMS", None)
    run(f"vector stage after derivation wave {start}, all SMs")
    for sms in os.environ.get("SMS", "96").split(","):
        os.environ["B200_VEC_SMS"] = sms
        run(f"vector stage after derivation wave {start}, GEMM on {sms} SMs")
os.environ.pop("B200_VEC_SMS", None)
os.environ.pop("B200_VEC_START", None)
os.environ["B200_HYBRID_SERIAL"] = "1"
run("serial")
del os.environ["B200_HYBRID_SERIAL"]
t0 = time.perf_counter()
for i in range(4):
    ix.search().query(batches[i % 4]).scoring_strategy("detailed").execute()
print(f"keyword detailed only: {1e3 * (time.perf_counter() - t0) / 4:.1f} ms/batch", flush=True)
