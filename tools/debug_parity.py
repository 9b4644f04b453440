"""Ad-hoc GPU debugging: print detailed diffs between the CUDA path and the oracle."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import meilisearch_b200 as mb
from oracle.pyoracle import OracleIndex
from tests.helpers import image_from_corpus, load_goldens, synthetic_image

def show(tag, queries, got, want, limit=6):
    n = 0
    for q in range(len(queries)):
        if got.ids(q) != want.ids(q) or got.scores(q) != want.scores(q):
            n += 1
            if n <= limit:
                print(f"--- {tag} MISMATCH {queries[q]!r} status={got.status[q]} cand {got.n_candidates[q]} vs {want.n_candidates[q]}")
                gi, wi = got.ids(q), want.ids(q)
                gs, ws = got.scores(q), want.scores(q)
                for k in range(max(len(gi), len(wi))):
                    g = (gi[k], gs[k]) if k < len(gi) else None
                    w = (wi[k], ws[k]) if k < len(wi) else None
                    print("   ", "==" if g == w else "!=", g, "|", w)
    print(f"{tag}: {n} mismatches of {len(queries)}")

G = load_goldens()
for case in G["cases"]:
    if case["source"].endswith("exactness.rs:885") or case["source"].endswith("exactness.rs:921"):
        img = image_from_corpus(G["corpora"][case["index"]])
        s = case["settings"]
        for d, doc in enumerate(G["corpora"][case["index"]]["docs"]): print(d, doc)
        ix = mb.Index(img, criteria=s.get("criteria"))
        o = OracleIndex(img, criteria=s.get("criteria"))
        tb = mb.TokenBatch([case["query"]])
        got = ix.search().query(tb).terms_matching_strategy(case["tms"]).scoring_strategy("detailed").execute()
        want = o.search_batch(tb, tms=case["tms"], scoring="detailed")
        print(case["source"], case["query"], s, "expected", case["expected_ids"])
        show("golden", [case["query"]], got, want)
        print("got", got.ids(0), got.scores(0)); print("want", want.ids(0), want.scores(0))

synth = synthetic_image(60000, 25000, seed=11)
queries = synth.synthetic_queries(300, seed=21)
tokens = mb.TokenBatch(queries)
ix = mb.Index(synth)
o = OracleIndex(synth)
for tms in ("last", "all"):
    got = ix.search().query(tokens).terms_matching_strategy(tms).scoring_strategy("detailed").execute()
    want = o.search_batch(tokens, tms=tms, scoring="detailed", n_threads=8)
    show(tms, queries, got, want)
print(json.dumps(ix.stats(), default=str)[:1500])
