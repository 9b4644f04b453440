#!/usr/bin/env python
"""Extract golden vectors from the reference's own ranking-rule tests into milli_goldens.json.

Run in the build container (the reference checkout is not present on the GPU box):
    python tests/golden/extract_reference_goldens.py

It parses crates/milli/src/search/new/tests/*.rs: the `documents!([...])`
corpus of every `create_*index()` helper, and — walking each `#[test]` body in source order — the settings
updates, `terms_matching_strategy`, `query`, `limit/offset` calls and the inline `insta` snapshot of
`documents_ids` that follows them.  Every emitted case carries file:line of its snapshot assertion.
Cases that need features outside the hot-path scope (sort, geo, distinct, filters, attributesToSearchOn,
proximityPrecision=byAttribute, custom separators/dictionary, non-ASCII text) are skipped and counted.
"""
import json
import os
import re
import sys

REF = "/root/reference/crates/milli/src/search/new/tests"
FILES = ["typo.rs", "proximity.rs", "words_tms.rs", "attribute_fid.rs", "word_position.rs", "exactness.rs",
         "ngram_split_words.rs", "typo_proximity.rs", "proximity_typo.rs", "stop_words.rs"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "milli_goldens.json")

CRIT = {"Words": "words", "Typo": "typo", "Proximity": "proximity", "Attribute": "attribute", "AttributeRank": "attributeRank",
        "WordPosition": "wordPosition", "Sort": "sort", "Exactness": "exactness"}
UNSUPPORTED = ["sort_criteria", "set_sortable", "filter(", "distinct(", "set_distinct", "searchable_attributes(", "ProximityPrecision::ByAttribute",
               "set_separator_tokens", "set_non_separator_tokens", "set_dictionary", "geo", "Criterion::Asc", "Criterion::Desc",
               "set_localized", "locales(", "semantic(", "time_budget", "Deadline", "ranking_score_threshold", "set_prefix_search",
               "exhaustive_number_hits", "set_searchable_fields(vec![S(", "set_filterable"]


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def rust_str(s):
    return bytes(s, "utf-8").decode("unicode_escape").encode("latin-1").decode("utf-8") if "\\" in s else s


def find_block(src, start):
    """src[start] == '{' -> index just after the matching '}' (aware of strings, raw strings, char literals)."""
    depth = 0
    i = start
    n = len(src)
    while i < n:
        c = src[i]
        if c == "r" and re.match(r'r#*"', src[i:i + 8]) and not (src[i - 1].isalnum() or src[i - 1] == "_"):
            hashes = re.match(r'r(#*)"', src[i:i + 8]).group(1)
            j = src.find('"' + hashes, i + len(hashes) + 2)
            i = j + len(hashes) + 1
            continue
        if c == "'" and i + 2 < n and src[i + 2] == "'":
            i += 3
            continue
        if c == '"':
            i += 1
            while src[i] != '"':
                if src[i] == "\\":
                    i += 1
                i += 1
        elif c == "/" and src[i + 1] == "/":
            i = src.find("\n", i)
            continue
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def parse_documents(body):
    m = re.search(r"documents!\(\[", body)
    if not m:
        return None
    i = m.end()
    docs = []
    while True:
        j = body.find("{", i)
        close = body.find("]))", i)
        if j < 0 or (0 <= close < j):
            break
        k = find_block(body, j)
        obj = strip_comments(body[j:k])
        obj = re.sub(r",\s*}", "}", obj)
        try:
            d = json.loads(obj, strict=False)
        except Exception:
            return None
        docs.append(d)
        i = k
    # a later document with the same primary key replaces the earlier one and keeps its internal docid
    seen = {}
    for d in docs:
        seen[json.dumps(d.get('id'))] = d
    out, done = [], set()
    for d in docs:
        k = json.dumps(d.get('id'))
        if k not in done:
            done.add(k)
            out.append(seen[k])
    return out


def str_list(arg):
    return [rust_str(x) for x in re.findall(r'"((?:[^"\\]|\\.)*)"', arg)]


SETTING_RE = re.compile(
    r"s\.set_criteria\(vec!\[(?P<crit>[^\]]*)\]\)|s\.set_searchable_fields\(\s*(?:vec!)?(?P<searchable>\[[^\]]*\])|"
    r"s\.set_authorize_typos\((?P<auth>true|false)\)|s\.set_exact_words\((?P<exw>[^;]*?)\);|"
    r"s\.set_exact_attributes\((?P<exa>[^;]*?)\);|syn\w*\.insert\((?P<syn>[^;]*?)\);|"
    r"s\.set_min_word_len_one_typo\((?P<ot>\d+)\)|s\.set_min_word_len_two_typos\((?P<tt>\d+)\)|"
    r"s\.set_stop_words\((?P<stop>[^;]*?)\);|"
    r"s\.terms_matching_strategy\(TermsMatchingStrategy::(?P<tms>\w+)\)|s\.query\(\"(?P<query>(?:[^\"\\]|\\.)*)\"\)|"
    r"s\.limit\((?P<limit>\d+)\)|s\.offset\((?P<offset>\d+)\)|"
    r"ScoringStrategy::(?P<scoring>Detailed|Skip)|"
    r"let mut s = (?P<news>index\.search|Search::new)|"
    r"assert_snapshot!\(format!\(\"\{(?P<var>\w+):\?\}\"\),\s*@\"(?P<ids>\[[^\"]*\])\"\)|"
    r"(?P<create>create_\w*index\w*)\(\)",
    re.S)


def apply_setting(m, st):
    if m.group("crit") is not None:
        st["criteria"] = [CRIT[c] for c in re.findall(r"Criterion::(\w+)", m.group("crit"))]
    elif m.group("searchable") is not None:
        st["searchable"] = str_list(m.group("searchable"))
    elif m.group("auth") is not None:
        st["authorize_typos"] = m.group("auth") == "true"
    elif m.group("exw") is not None:
        st["exact_words"] = str_list(m.group("exw"))
    elif m.group("exa") is not None:
        st["exact_attributes"] = str_list(m.group("exa"))
    elif m.group("syn") is not None:
        parts = str_list(m.group("syn"))
        st.setdefault("synonyms", {}).setdefault(parts[0], []).extend(parts[1:])
    elif m.group("ot") is not None:
        st["one_typo"] = int(m.group("ot"))
    elif m.group("tt") is not None:
        st["two_typos"] = int(m.group("tt"))
    elif m.group("stop") is not None:
        st["stop_words"] = str_list(m.group("stop"))
    else:
        return False
    return True


def main():
    cases, skipped = [], 0
    for fname in FILES:
        path = os.path.join(REF, fname)
        src = open(path, encoding="utf-8").read()
        # index builders
        builders = {}
        for m in re.finditer(r"fn (create_\w*index\w*)\(\)\s*->\s*TempIndex\s*\{", src):
            end = find_block(src, m.end() - 1)
            body = src[m.end():end]
            st = {}
            for sm in SETTING_RE.finditer(body[: body.find("documents!") if "documents!" in body else len(body)]):
                apply_setting(sm, st)
            docs = parse_documents(body)
            unsupported = any(u in body for u in UNSUPPORTED) or docs is None or "searchable" not in st
            if docs is not None and not all(ord(ch) < 128 for d in docs for v in d.values() if isinstance(v, str) for ch in v):
                unsupported = True
            builders[m.group(1)] = (st, docs, unsupported)
        # tests
        for m in re.finditer(r"#\[test\]\s*fn (\w+)\(\)\s*\{", src):
            end = find_block(src, m.end() - 1)
            body = src[m.end():end]
            base_line = src.count("\n", 0, m.end()) + 1
            if any(u in body for u in UNSUPPORTED):
                skipped += 1
                continue
            st, docs, bad = None, None, True
            cur = {"tms": "Last", "query": None, "limit": 20, "offset": 0, "scoring": "Skip"}
            for sm in SETTING_RE.finditer(body):
                if sm.group("create") is not None:
                    if sm.group("create") not in builders:
                        bad = True
                        continue
                    bst, docs, bad = builders[sm.group("create")]
                    st = json.loads(json.dumps(bst))
                    continue
                if st is None:
                    continue
                if apply_setting(sm, st):
                    continue
                if sm.group("news") is not None:
                    cur = {"tms": "Last", "query": None, "limit": 20, "offset": 0, "scoring": "Skip"}
                elif sm.group("tms") is not None:
                    cur["tms"] = sm.group("tms")
                elif sm.group("query") is not None:
                    cur["query"] = rust_str(sm.group("query"))
                elif sm.group("limit") is not None:
                    cur["limit"] = int(sm.group("limit"))
                elif sm.group("offset") is not None:
                    cur["offset"] = int(sm.group("offset"))
                elif sm.group("scoring") is not None:
                    cur["scoring"] = sm.group("scoring")
                elif sm.group("ids") is not None:
                    if not sm.group("var").startswith(("documents_ids", "ids")) or cur["query"] is None:
                        continue
                    if bad or not all(ord(ch) < 128 for ch in cur["query"]):
                        skipped += 1
                        continue
                    line = base_line + body.count("\n", 0, sm.start())
                    cases.append({
                        "source": f"crates/milli/src/search/new/tests/{fname}:{line}", "test": m.group(1),
                        "index": {"searchable": st["searchable"], "exact_attributes": st.get("exact_attributes", []),
                                  "stop_words": st.get("stop_words", []), "docs": docs},
                        "settings": {k: st[k] for k in ("criteria", "authorize_typos", "exact_words", "synonyms", "one_typo", "two_typos") if k in st},
                        "tms": cur["tms"].lower(), "scoring": cur["scoring"].lower(), "limit": cur["limit"], "offset": cur["offset"],
                        "query": cur["query"], "expected_ids": json.loads(sm.group("ids")),
                    })
    # de-duplicate identical corpora into a table to keep the fixture small
    corpora, keyed = [], {}
    for c in cases:
        k = json.dumps(c["index"], sort_keys=True)
        if k not in keyed:
            keyed[k] = len(corpora)
            corpora.append(c["index"])
        c["index"] = keyed[k]
    json.dump({"generated_by": "tests/golden/extract_reference_goldens.py", "reference": "meilisearch v1.50.0 @ 5cb2f2e",
               "corpora": corpora, "cases": cases}, open(OUT, "w"), indent=0)
    print(f"{len(cases)} cases, {len(corpora)} corpora, {skipped} skipped -> {OUT}")
    by = {}
    for c in cases:
        f = c["source"].split("/")[-1].split(":")[0]
        by[f] = by.get(f, 0) + 1
    print(by)


if __name__ == "__main__":
    sys.exit(main())
