import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meilisearch_b200 as mb
from oracle.pyoracle import OracleIndex
from tests.helpers import synthetic_image
synth = synthetic_image(60000, 25000, seed=11)
qs = sys.argv[1:] or ["cwtrq rm"]
ix = mb.Index(synth); o = OracleIndex(synth)
for q in qs:
    tb = mb.TokenBatch([q])
    os.environ["B200_DEBUG"] = "1"
    got = ix.search().query(tb).scoring_strategy("detailed").execute()
    want = o.search_batch(tb, scoring="detailed")
    print(q, "got", got.ids(0)); print("want", want.ids(0))
    for k,(a,b) in enumerate(zip(got.scores(0), want.scores(0))): print(k, a, b)
    # position lists of the first-word derivations
    import numpy as np
