import os, sys
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/meilisearch_b200") else ".")
import numpy as np
import meilisearch_b200 as mb
from corpus.pyindexgen import IndexImage
img = IndexImage(1); img.add_text(0, 0, "placeholder"); img.build()
ix = mb.Index(img)
rng = np.random.default_rng(0)
n, d = 1000000, 768
ix.set_embeddings(rng.standard_normal((n, d), dtype=np.float32))
for B in (1024,):
    qq = rng.standard_normal((B, d), dtype=np.float32)
    for ss in ("0", "1"):
        os.environ["B200_VEC_GEMM_TS"] = "0" if ss == "1" else "1"
        os.environ.pop("B200_VEC_DEBUG", None)
        for i in range(2): ix.nns_by_vector(qq, 100)
        ix.reset_stats()
        for i in range(3): ix.nns_by_vector(qq, 100)
        s = ix.stats()["kernels"]["vec_gemm_topk"]
        print("B", B, "SS" if ss == "1" else "TS", "%.3f ms" % (s["ms"] / s["count"]), flush=True)
        for mode in ("1", "2"):
            os.environ["B200_VEC_DEBUG"] = mode
            ix.nns_by_vector(qq, 100)
