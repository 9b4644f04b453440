"""Query tokenizer standing in for charabia (crates/milli/src/search/new/mod.rs:918-996) on plain
lowercase Latin text.  The drop-in boundary takes *tokens*: in a deployment the Rust host keeps
calling charabia and hands the normalized token stream across the C ABI; this helper exists so the
Python mirror, the tests and the bench can produce the same stream.

Token kinds follow charabia's TokenKind as used by located_query_terms_from_tokens
(query_term/parse_query.rs:64-181): Word, StopWord, Separator(Soft), Separator(Hard).
Separators are emitted one character at a time, except ". " and ", " which are the hard context
separators (two characters)."""
from __future__ import annotations

import numpy as np

WORD, STOPWORD, SEP_SOFT, SEP_HARD = 0, 1, 2, 3
_HARD_SINGLE = {";", "!", "?"}


def _is_word_char(c: str) -> bool:
    return c.isalnum() or ord(c) >= 0x80


def tokenize(query: str, stop_words=frozenset()):
    """-> list[(kind, lemma)]"""
    s = query  # stop words are matched as written (the reference's set is case sensitive, stop_words.rs:1-10), lemmas are lowercased
    out = []
    i, n = 0, len(s)
    while i < n:
        c = s[i]
        if _is_word_char(c):
            j = i
            while j < n and _is_word_char(s[j]):
                j += 1
            w = s[i:j]
            out.append((STOPWORD if w in stop_words else WORD, w.lower()))
            i = j
        else:
            if c in ".," and i + 1 < n and s[i + 1] == " ":
                out.append((SEP_HARD, s[i : i + 2]))
                i += 2
            elif c in _HARD_SINGLE:
                out.append((SEP_HARD, c))
                i += 1
            else:
                out.append((SEP_SOFT, c))
                i += 1
    return out


class TokenBatch:
    """Flat, C-ABI-ready layout of a batch of tokenized queries (host buffers)."""

    def __init__(self, queries, stop_words=frozenset()):
        begin = [0]
        kinds = []
        offs = [0]
        chunks = []
        total = 0
        for q in queries:
            toks = q if isinstance(q, list) else tokenize(q, stop_words)
            for k, lemma in toks:
                b = lemma.encode("utf-8")
                kinds.append(k)
                chunks.append(b)
                total += len(b)
                offs.append(total)
            begin.append(len(kinds))
        self.n_queries = len(queries)
        self.token_begin = np.asarray(begin, dtype=np.uint32)
        self.token_kind = np.asarray(kinds, dtype=np.uint8) if kinds else np.zeros(1, np.uint8)
        self.lemma_off = np.asarray(offs, dtype=np.uint32)
        self.lemma_bytes = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8).copy()

    def head(self, n):
        """the first n queries as a batch of their own (shares no buffers with self)"""
        n = min(n, self.n_queries)
        t = TokenBatch.__new__(TokenBatch)
        t.n_queries = n
        t.token_begin = self.token_begin[: n + 1].copy()
        nt = int(t.token_begin[-1])
        t.token_kind = self.token_kind[: max(nt, 1)].copy()
        t.lemma_off = self.lemma_off[: nt + 1].copy()
        t.lemma_bytes = self.lemma_bytes.copy()
        return t
