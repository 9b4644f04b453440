"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/b200milli.h declares, and
refuses to run without a device instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import meilisearch_b200 as mb

    mb.build_library()
    lib = mb.load_library()
    hdr = open(os.path.join(ROOT, "include", "b200milli.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200milli.h but not exported"
    assert declared == set(mb.SYMBOLS)


def test_no_cpu_fallback():
    import torch

    import meilisearch_b200 as mb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mb.B200Error) as e:
        mb.Index()
    assert e.value.code == -1


def test_product_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "meilisearch_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".sh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "milli_oracle" not in src and "oracle/" not in src, (dirpath, f)


def test_tokenizer():
    from meilisearch_b200.tokenizer import SEP_HARD, SEP_SOFT, WORD, TokenBatch, tokenize

    assert tokenize("Hello, wor-ld. x") == [(WORD, "hello"), (SEP_HARD, ", "), (WORD, "wor"), (SEP_SOFT, "-"), (WORD, "ld"), (SEP_HARD, ". "), (WORD, "x")]
    tb = TokenBatch(["a b", ""])
    assert list(tb.token_begin) == [0, 3, 3]


def test_integration_doc_declares_every_entry_point():
    """INTEGRATION.md's `extern "C"` block is the binding a maintainer would add: it must name every function of the header"""
    hdr = open(os.path.join(ROOT, "include", "b200milli.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    bound = set(re.findall(r"pub fn (b200_[a-z0-9_]+)\(", doc))
    assert declared - bound == set(), declared - bound
