"""Shared test helpers: build index images from fixture corpora."""
import json
import os

from corpus.pyindexgen import IndexImage

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_goldens():
    return json.load(open(os.path.join(GOLDEN_DIR, "milli_goldens.json")))


def image_from_corpus(c):
    fields = c["searchable"]
    mask = sum(1 << fields.index(f) for f in c["exact_attributes"] if f in fields)
    img = IndexImage(len(fields), mask, c["stop_words"])
    for d, doc in enumerate(c["docs"]):
        if doc is None:  # a document the extraction dropped (non-ASCII text); its docid stays unused
            continue
        any_text = False
        for f, name in enumerate(fields):
            v = doc.get(name)
            if isinstance(v, str):
                img.add_text(d, f, v)
                any_text = True
        if not any_text:
            img.add_text(d, 0, "")
    return img.build()


_synth_cache = {}


def synthetic_image(n_docs, vocab, seed=0xB200, n_fields=1):
    key = (n_docs, vocab, seed, n_fields)
    if key not in _synth_cache:
        img = IndexImage(n_fields)
        img.add_synthetic(n_docs, vocab, seed=seed)
        _synth_cache[key] = img.build()
    return _synth_cache[key]
