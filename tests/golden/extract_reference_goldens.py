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
SNAPDIR = REF + "/snapshots"
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


def find_paren(src, start):
    """src[start] == '(' -> index just after the matching ')' (aware of strings and raw strings)."""
    depth, i, n = 0, start, len(src)
    while i < n:
        c = src[i]
        if c == "r" and re.match(r'r#*"', src[i:i + 8]) and not (src[i - 1].isalnum() or src[i - 1] == "_"):
            hashes = re.match(r'r(#*)"', src[i:i + 8]).group(1)
            j = src.find('"' + hashes, i + len(hashes) + 2)
            i = j + len(hashes) + 1
            continue
        if c == '"':
            i += 1
            while src[i] != '"':
                if src[i] == "\\":
                    i += 1
                i += 1
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced parens")


SCORE_RE = re.compile(
    r"Words\(\s*Words\s*\{\s*matching_words:\s*(\d+),\s*max_matching_words:\s*(\d+),?\s*\},?\s*\)|"
    r"Typo\(\s*Typo\s*\{\s*typo_count:\s*(\d+),\s*max_typo_count:\s*(\d+),?\s*\},?\s*\)|"
    r"(Proximity|Fid|Position)\(\s*Rank\s*\{\s*rank:\s*(\d+),\s*max_rank:\s*(\d+),?\s*\},?\s*\)|"
    r"ExactAttribute\(\s*(ExactMatch|MatchesStart|NoExactMatch),?\s*\)|"
    r"ExactWords\(\s*ExactWords\s*\{\s*matching_words:\s*(\d+),\s*max_matching_words:\s*(\d+),?\s*\},?\s*\)", re.S)


def parse_score_list(txt):
    """one `[ ScoreDetails, ... ]` -> [[kind, rank, max_rank], ...] using ScoreDetails::rank() (score_details.rs:110-125)"""
    out = []
    for m in SCORE_RE.finditer(txt):
        if m.group(1) is not None:
            out.append(["words", int(m.group(1)), int(m.group(2))])
        elif m.group(3) is not None:
            tc, mx = int(m.group(3)), int(m.group(4))
            out.append(["typo", mx + 1 - tc, mx + 1])
        elif m.group(5) is not None:
            out.append([{"Proximity": "proximity", "Fid": "fid", "Position": "position"}[m.group(5)], int(m.group(6)), int(m.group(7))])
        elif m.group(8) is not None:
            out.append(["exactAttribute", {"ExactMatch": 3, "MatchesStart": 2, "NoExactMatch": 1}[m.group(8)], 3])
        else:
            out.append(["exactWords", int(m.group(9)) + 1, int(m.group(10)) + 1])
    return out


def split_top_level(body):
    """split the inside of `[ a, b, ... ]` at depth-0 commas"""
    parts, depth, cur = [], 0, []
    for ch in body:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    return parts


def parse_snapshot_file(path):
    txt = open(path, encoding="utf-8").read()
    body = txt.split("---", 2)[2].strip()
    inner = body[body.index("[") + 1: body.rindex("]")]
    items = split_top_level(inner)
    ids, scores = [], []
    for it in items:
        it = it.strip()
        if it.startswith("("):  # (docid, [scores])
            m = re.match(r"\(\s*(\d+),", it)
            ids.append(int(m.group(1)))
            scores.append(parse_score_list(it[m.end():]))
        else:
            scores.append(parse_score_list(it))
    return (ids if ids else None), scores


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
    r"(?P<snap>insta::assert_(?:debug_)?snapshot!)\(|"
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



def query_criteria_cases():
    """crates/milli/tests/search/query_criteria.rs:26-101 over crates/milli/tests/assets/test_set.ndjson: 17 documents carrying
    their pre-computed per-rule ranks; the expected order is `expected_order` (crates/milli/tests/search/mod.rs:150-218): a stable
    sort by each criterion's rank, documents with word_rank != 0 dropped under TermsMatchingStrategy::All.  Index settings:
    tests/search/mod.rs:37-68 (searchable title+description, three synonyms).  Only the text criteria are transcribed (no
    asc/desc/sort)."""
    text = open("/root/reference/crates/milli/tests/assets/test_set.ndjson").read()
    dec, i, docs = json.JSONDecoder(), 0, []
    while True:
        while i < len(text) and text[i].isspace():
            i += 1
        if i >= len(text):
            break
        o, i = dec.raw_decode(text, i)
        docs.append(o)
    index = {"searchable": ["title", "description"], "exact_attributes": [], "stop_words": [],
             "docs": [{"id": k, "title": d["title"], "description": d["description"]} for k, d in enumerate(docs)]}
    synonyms = {"hello": ["good morning"], "world": ["earth"], "america": ["the united states"]}
    rank_key = {"words": "word_rank", "typo": "typo_rank", "proximity": "proximity_rank", "attribute": "attribute_rank", "exactness": "exact_rank"}
    tests = [("none", 62, "all", []), ("words", 63, "last", ["words"]), ("attribute", 64, "all", ["attribute"]), ("typo", 65, "all", ["typo"]),
             ("exactness", 66, "all", ["exactness"]), ("proximity", 67, "all", ["proximity"]),
             ("default_criteria_order", 97, "last", ["words", "typo", "proximity", "attribute", "exactness"])]
    # criteria_mixup (query_criteria.rs:104-292) walks the 120 orders of {attribute, desc(asc_desc_rank), exactness, proximity, typo}
    # after `words`, asserting `expected_order` for each.  The Desc criterion is a sort rule (out of this path's scope), so the 24
    # orders of the four text criteria are transcribed instead, each under the test's strategy (Last) and under All, with the
    # expected order computed exactly as the test does (tests/search/mod.rs:150-218, implemented below).
    import itertools
    for perm in itertools.permutations(["attribute", "exactness", "proximity", "typo"]):
        for tms in ("last", "all"):
            tests.append(("criteria_mixup " + ",".join(perm), 104, tms, ["words"] + list(perm)))
    out = []
    for name, line, tms, criteria in tests:
        groups = [list(range(len(docs)))]
        for c in criteria:
            nxt = []
            for g in groups:
                g = sorted(g, key=lambda k: docs[k][rank_key[c]])  # stable, like slice::sort_by_key
                for k in g:  # linear_group_by_key
                    if nxt and nxt[-1][0] in g and docs[nxt[-1][-1]][rank_key[c]] == docs[k][rank_key[c]]:
                        nxt[-1].append(k)
                    else:
                        nxt.append([k])
            groups = nxt
        order = [k for g in groups for k in g]
        if tms == "all":
            order = [k for k in order if docs[k]["word_rank"] == 0]
        out.append({"source": f"crates/milli/tests/search/query_criteria.rs:{line}", "test": name, "index": index,
                    "settings": {"criteria": criteria, "synonyms": synonyms}, "tms": tms, "scoring": "skip", "limit": 17, "offset": 0,
                    "query": "hello world america", "expected_ids": order})
    return out


def matching_strategy_cases():
    """crates/meilisearch/tests/search/matching_strategy.rs:16-140: 7 documents, default settings (every field searchable; serde_json
    maps iterate in key order, so fid 0 = `id`, fid 1 = `title`), three queries under matchingStrategy last / all / frequency with
    the `hits` snapshots.  External ids "1".."7" were inserted in order: internal docid = id - 1."""
    path = "/root/reference/crates/meilisearch/tests/search/matching_strategy.rs"
    src = open(path).read()
    m = re.search(r"static SIMPLE_SEARCH_DOCUMENTS.*?json!\(\[(.*?)\]\)\s*\}\);", src, re.S)
    docs = json.loads("[" + re.sub(r",(\s*[}\]])", r"\1", m.group(1)) + "]")
    # `id` is a searchable text field here; the position in the list is the internal docid
    index = {"searchable": ["id", "title"], "exact_attributes": [], "stop_words": [], "docs": [{"id": d["id"], "title": d["title"]} for d in docs]}
    out = []
    for sm in re.finditer(r'\.search\(json!\(\{"q": "([^"]*)", "matchingStrategy": "(\w+)".*?snapshot!\(response\["hits"\], @(?:r###)?"(.*?)"(?:###)?\);', src, re.S):
        line = src.count("\n", 0, sm.start()) + 1
        if line > 140:
            break
        hits = json.loads(sm.group(3))
        out.append({"source": f"crates/meilisearch/tests/search/matching_strategy.rs:{line}", "test": "matching_strategy", "index": index,
                    "settings": {}, "tms": sm.group(2), "scoring": "skip", "limit": 20, "offset": 0, "query": sm.group(1),
                    "expected_ids": [int(h["id"]) - 1 for h in hits]})
    return out


def http_search_cases():
    """crates/meilisearch/tests/search/mod.rs: known answers of the HTTP search tests that exercise only this path, transcribed by hand
    (source line beside every row).  Fixtures: crates/meilisearch/tests/common/mod.rs:223-256 `DOCUMENTS` and :276-300
    `SCORE_DOCUMENTS`; default settings (searchableAttributes ["*"]: every field searchable with the same weight — `weights` all 0;
    serde_json maps iterate in key order, `_vectors` is not indexed as text).  The title "Gläss" is stored as "Glass": the reference's
    tokenizer folds the diacritic (the test at mod.rs:86 searches "glass"), the tokenizer mirror here is ASCII-only.
    `expected_ranking_scores` = `_rankingScore`, `expected_candidates` = `estimatedTotalHits`, `expected_rule_scores` = the `score`
    of every `_rankingScoreDetails` entry in rule order (exactness = ExactAttribute's rank / max when no word-level detail applies)."""
    src = "crates/meilisearch/tests/search/mod.rs"
    docs = [("287947", "Shazam!", "green blue"), ("299537", "Captain Marvel", "yellow blue"), ("522681", "Escape Room", "yellow red"),
            ("166428", "How to Train Your Dragon: The Hidden World", "green red"), ("450465", "Glass", "blue red")]

    def documents(stop_words=()):
        return {"searchable": ["color", "id", "title"], "exact_attributes": [], "stop_words": list(stop_words),
                "docs": [{"color": c, "id": i, "title": t} for i, t, c in docs]}

    score_docs = {"searchable": ["id", "title"], "exact_attributes": [], "stop_words": [],
                  "docs": [{"id": i, "title": t} for i, t in (("A", "Batman the dark knight returns: Part 1"), ("B", "Batman the dark knight returns: Part 2"),
                                                               ("C", "Batman Returns"), ("D", "Batman"), ("E", "Badman"))]}
    w3, w2 = {"weights": [0, 0, 0]}, {"weights": [0, 0]}
    stop7 = ("the", "The", "a", "an", "to", "in", "of")
    rows = [
        # line, test, index, settings, query, expected ids, extras
        (70, "simple_placeholder_search", documents(), w3, "", [0, 1, 2, 3, 4], {}),
        (89, "simple_search", documents(), w3, "glass", [4], {}),
        (141, "search_with_stop_word", documents(stop7), w3, "to the", [], {}),
        (149, "search_with_stop_word", documents(stop7), w3, "to the ", [0, 1, 2, 3, 4], {}),
        (237, "phrase_search_with_stop_word", documents(("the", "of")), w3, 'how "to" train "the', [3], {}),
        (248, "negative_phrase_search", documents(), w3, '-"train your dragon"', [0, 1, 2, 4], {}),
        (264, "negative_word_search", documents(), w3, "-escape", [0, 1, 3, 4], {}),
        (277, "negative_word_search", documents(), w3, "-escape escape", [], {}),
        (289, "non_negative_search", documents(), w3, "- escape", [2], {}),
        (298, "non_negative_search", documents(), w3, '- "train your dragon"', [3], {}),
        (322, "negative_special_cases_search", documents(), dict(w3, synonyms={"escape": ["glass"]}), "-escape escape", [4], {}),
        (781, "test_score_details", documents(), w3, "train dragon", [3],
         {"expected_rule_scores": [[1.0, 1.0, 0.75, 1.0, 0.8095238095238095, 0.3333333333333333]]}),
        (924, "test_score", score_docs, w2, "Badman the dark knight returns 1", [0, 1, 4, 2, 3],
         {"expected_ranking_scores": [0.9746605609456898, 0.8055252965383685, 0.16666666666666666, 0.07702020202020202, 0.07702020202020202]}),
        (973, "test_score_threshold", score_docs, w2, "Badman dark returns 1", [0, 1, 4, 2, 3],
         {"threshold": 0.0, "expected_candidates": 5,
          "expected_ranking_scores": [0.93430081300813, 0.6685627880184332, 0.25, 0.11553030303030302, 0.11553030303030302]}),
        (1016, "test_score_threshold", score_docs, w2, "Badman dark returns 1", [0, 1, 4],
         {"threshold": 0.2, "expected_candidates": 3, "expected_ranking_scores": [0.93430081300813, 0.6685627880184332, 0.25]}),
        (1049, "test_score_threshold", score_docs, w2, "Badman dark returns 1", [0, 1],
         {"threshold": 0.5, "expected_candidates": 2, "expected_ranking_scores": [0.93430081300813, 0.6685627880184332]}),
        (1077, "test_score_threshold", score_docs, w2, "Badman dark returns 1", [0],
         {"threshold": 0.8, "expected_candidates": 1, "expected_ranking_scores": [0.93430081300813]}),
        (1100, "test_score_threshold", score_docs, w2, "Badman dark returns 1", [], {"threshold": 1.0, "expected_candidates": 0}),
    ]
    out = []
    for line, test, index, st, q, ids, extra in rows:
        c = {"source": f"{src}:{line}", "test": test, "index": index, "settings": dict(st), "tms": "last", "scoring": "detailed", "limit": 20,
             "offset": 0, "query": q, "expected_ids": ids}
        c.update(extra)
        out.append(c)
    # crates/meilisearch/tests/search/multi/mod.rs: multi-search / federated search over the same fixtures.  Federation itself is outside
    # this path, but every merged hit carries the ranking score it got from the query at `queriesPosition` (federation weight 1.0):
    # `expected_doc_scores` = [docid, weightedRankingScore] pairs a single query must reproduce (expected_ids null: the merged list
    # does not show the query's other hits).
    msrc = "crates/meilisearch/tests/search/multi/mod.rs"
    mrows = [
        (164, "simple_search_single_index", documents(), w3, "glass", [4], []),
        (165, "simple_search_single_index", documents(), w3, "captain", [1], []),
        (329, "federation_two_search_single_index", documents(), w3, "glass", None, [[4, 1.0]]),
        (330, "federation_two_search_single_index", documents(), w3, "captain", None, [[1, 0.9848484848484848]]),
        (628, "federation_multiple_search_multiple_indexes", documents(), w3, "Escape", None, [[2, 0.9848484848484848]]),
        (631, "federation_multiple_search_multiple_indexes", documents(), w3, "the bat", None, [[3, 0.4166666666666667]]),
        (258, "federation_multiple_search_single_index", score_docs, w2, "badman returns", None, [[4, 0.5]]),
        (259, "federation_multiple_search_single_index", score_docs, w2, "batman", None, [[3, 1.0], [0, 0.9848484848484848], [1, 0.9848484848484848]]),
        (260, "federation_multiple_search_single_index", score_docs, w2, "batman returns", None, [[2, 1.0]]),
    ]
    for line, test, index, st, q, ids, doc_scores in mrows:
        c = {"source": f"{msrc}:{line}", "test": test, "index": index, "settings": dict(st), "tms": "last", "scoring": "detailed", "limit": 20,
             "offset": 0, "query": q, "expected_ids": ids}
        if doc_scores:
            c["expected_doc_scores"] = doc_scores
        out.append(c)
    return out


def phrase_search_count_cases():
    """crates/milli/tests/search/phrase_search.rs:27-63: test_set.ndjson, stop words a/an/the/of, the phrase query
    "the use of force" under TermsMatchingStrategy::All matches exactly one document, with no criteria and with
    [proximity, attribute, exactness]."""
    text = open("/root/reference/crates/milli/tests/assets/test_set.ndjson").read()
    dec, i, docs = json.JSONDecoder(), 0, []
    while True:
        while i < len(text) and text[i].isspace():
            i += 1
        if i >= len(text):
            break
        o, i = dec.raw_decode(text, i)
        docs.append(o)
    index = {"searchable": ["title", "description"], "exact_attributes": [], "stop_words": ["a", "an", "the", "of"],
             "docs": [{"id": k, "title": d["title"], "description": d["description"]} for k, d in enumerate(docs)]}
    synonyms = {"hello": ["good morning"], "world": ["earth"], "america": ["the united states"]}
    out = []
    for line, criteria in ((54, []), (60, ["proximity", "attribute", "exactness"])):
        out.append({"source": f"crates/milli/tests/search/phrase_search.rs:{line}", "test": "phrase_search_with_stop_words", "index": index,
                    "settings": {"criteria": criteria, "synonyms": synonyms}, "tms": "all", "scoring": "skip", "limit": 10, "offset": 0,
                    "query": '"the use of force"', "expected_count": 1})
    return out


def typo_tolerance_count_cases():
    """crates/milli/tests/search/typo_tolerance.rs:18-357: known-answer tests on `documents_ids.len()` (test_set.ndjson with criteria
    [Typo] and the three synonyms of tests/search/mod.rs:37-68, or two inline documents).  Only the hit COUNT is asserted by the
    reference, so these are kept apart from the ordered goldens (`count_cases`, oracle-only)."""
    text = open("/root/reference/crates/milli/tests/assets/test_set.ndjson").read()
    dec, i, docs = json.JSONDecoder(), 0, []
    while True:
        while i < len(text) and text[i].isspace():
            i += 1
        if i >= len(text):
            break
        o, i = dec.raw_decode(text, i)
        docs.append(o)
    def test_set(exact_attributes=()):
        return {"searchable": ["title", "description"], "exact_attributes": list(exact_attributes), "stop_words": [],
                "docs": [{"id": k, "title": d["title"], "description": d["description"]} for k, d in enumerate(docs)]}
    synonyms = {"hello": ["good morning"], "world": ["earth"], "america": ["the united states"]}
    two_docs = {"searchable": ["id", "data"], "exact_attributes": [], "stop_words": [],
                "docs": [{"id": "1", "data": "zealand"}, {"id": "2", "data": "zearand"}]}
    src = "crates/milli/tests/search/typo_tolerance.rs"
    base = {"criteria": ["typo"], "synonyms": synonyms}
    rows = [
        (43, test_set(), dict(base), "zeal", 1), (60, test_set(), dict(base), "zean", 0),
        (95, test_set(), dict(base, one_typo=4), "zean", 1),
        (123, test_set(), dict(base), "zealand", 1), (140, test_set(), dict(base), "zealemd", 0),
        (175, test_set(), dict(base, two_typos=7), "zealemd", 1),
        (255, two_docs, {}, "zealand", 2), (292, two_docs, {"exact_words": ["zealand"]}, "zealand", 1),
        (321, test_set(), dict(base), "antebelum", 1), (356, test_set(["description"]), dict(base), "antebelum", 0),
    ]
    return [{"source": f"{src}:{line}", "test": "typo_tolerance", "index": index, "settings": st, "tms": "last", "scoring": "skip", "limit": 10,
             "offset": 0, "query": q, "expected_count": n} for line, index, st, q, n in rows]


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
            if docs is not None:
                # documents with non-ASCII text are outside the ASCII tokenizer stand-in: they are dropped (their docids stay
                # reserved) and every case whose expected hits mention one of them is skipped below
                docs = [d if all(ord(ch) < 128 for v in d.values() if isinstance(v, str) for ch in v) else None for d in docs]
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
            snap_no, last_case, last_query = 0, None, None
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
                elif sm.group("snap") is not None:
                    # every insta assertion (inline or not) advances the per-test snapshot counter used in file names
                    snap_no += 1
                    close = find_paren(body, sm.end() - 1)
                    call = body[sm.end():close]
                    var = re.search(r'format!\("\{(\w+):#?\?\}"\)', call)
                    inline = re.search(r'@(r#*)?"', call)
                    line = base_line + body.count("\n", 0, sm.start())
                    if var is None or cur["query"] is None:
                        continue
                    vname = var.group(1)
                    usable = not bad and all(ord(ch) < 128 for ch in cur["query"])

                    def new_case(ids):
                        return {
                            "source": f"crates/milli/src/search/new/tests/{fname}:{line}", "test": m.group(1),
                            "index": {"searchable": st["searchable"], "exact_attributes": st.get("exact_attributes", []),
                                      "stop_words": st.get("stop_words", []), "docs": docs},
                            "settings": {k: st[k] for k in ("criteria", "authorize_typos", "exact_words", "synonyms", "one_typo", "two_typos") if k in st},
                            "tms": cur["tms"].lower(), "scoring": cur["scoring"].lower(), "limit": cur["limit"], "offset": cur["offset"],
                            "query": cur["query"], "expected_ids": ids,
                        }

                    if inline and vname.startswith(("documents_ids", "ids")):
                        lit = re.search(r'@"(\[[^"]*\])"', call)
                        if lit is None:
                            continue
                        if not usable or any(k >= len(docs) or docs[k] is None for k in json.loads(lit.group(1))):
                            skipped += 1
                            continue
                        cases.append(new_case(json.loads(lit.group(1))))
                        last_case = cases[-1]
                        last_query = cur["query"]
                    elif not inline and vname in ("document_scores", "document_ids_scores"):
                        stem = fname[:-3]
                        tname = m.group(1)[5:] if m.group(1).startswith("test_") else m.group(1)
                        snap = os.path.join(SNAPDIR, f"milli__search__new__tests__{stem}__{tname}" + (f"-{snap_no}" if snap_no > 1 else "") + ".snap")
                        if not os.path.exists(snap) or not usable:
                            skipped += 1
                            continue
                        ids, scores = parse_snapshot_file(snap)
                        if ids is not None and any(k >= len(docs) or docs[k] is None for k in ids):
                            skipped += 1
                            continue
                        rel = "crates/milli/src/search/new/tests/snapshots/" + os.path.basename(snap)
                        if vname == "document_ids_scores":
                            c = new_case(ids)
                            c["expected_scores"] = scores
                            c["scores_source"] = rel
                            cases.append(c)
                        elif last_case is not None and last_query == cur["query"] and len(scores) == len(last_case["expected_ids"]):
                            last_case["expected_scores"] = scores
                            last_case["scores_source"] = rel
    cases.extend(query_criteria_cases())
    cases.extend(matching_strategy_cases())
    cases.extend(http_search_cases())
    # de-duplicate identical corpora into a table to keep the fixture small
    corpora, keyed = [], {}
    for c in cases:
        k = json.dumps(c["index"], sort_keys=True)
        if k not in keyed:
            keyed[k] = len(corpora)
            corpora.append(c["index"])
        c["index"] = keyed[k]
    count_cases = typo_tolerance_count_cases() + phrase_search_count_cases()
    for c in count_cases:
        k = json.dumps(c["index"], sort_keys=True)
        if k not in keyed:
            keyed[k] = len(corpora)
            corpora.append(c["index"])
        c["index"] = keyed[k]
    json.dump({"generated_by": "tests/golden/extract_reference_goldens.py", "reference": "meilisearch v1.50.0 @ 5cb2f2e",
               "corpora": corpora, "cases": cases, "count_cases": count_cases}, open(OUT, "w"), indent=0)
    print(f"{len(cases)} cases, {len(corpora)} corpora, {skipped} skipped -> {OUT}")
    by = {}
    for c in cases:
        f = c["source"].split("/")[-1].split(":")[0]
        by[f] = by.get(f, 0) + 1
    print(by)


if __name__ == "__main__":
    sys.exit(main())
