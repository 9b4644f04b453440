// indexgen — see indexgen.h. Test/bench infrastructure standing in for milli's indexer.
#include "indexgen.h"

#include <omp.h>

#include <algorithm>
#include <parallel/algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct Tok {
    uint32_t word;
    uint32_t doc;
    uint16_t fid;
    uint16_t pos;
};

struct Db {
    std::vector<uint8_t> keys;
    std::vector<uint64_t> koff{0};
    std::vector<uint8_t> vals;
    std::vector<uint64_t> voff{0};
    void end_entry() {
        koff.push_back(keys.size());
        voff.push_back(vals.size());
    }
};

// CboRoaringBitmapCodec::serialize_into (cbo_roaring_bitmap_codec.rs:31-51) + roaring 0.10 portable format.
void cbo_encode(const uint32_t *ids, size_t n, std::vector<uint8_t> &out) {
    auto put32 = [&](uint32_t v) {
        uint8_t b[4];
        memcpy(b, &v, 4);
        out.insert(out.end(), b, b + 4);
    };
    auto put16 = [&](uint16_t v) {
        uint8_t b[2];
        memcpy(b, &v, 2);
        out.insert(out.end(), b, b + 2);
    };
    if (n <= 7) {
        for (size_t i = 0; i < n; i++) put32(ids[i]);
        return;
    }
    // containers
    std::vector<std::pair<size_t, size_t>> cont;  // [begin,end)
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        uint32_t key = ids[i] >> 16;
        while (j < n && (ids[j] >> 16) == key) j++;
        cont.push_back({i, j});
        i = j;
    }
    put32(12346);
    put32((uint32_t)cont.size());
    for (auto &c : cont) {
        put16((uint16_t)(ids[c.first] >> 16));
        put16((uint16_t)(c.second - c.first - 1));
    }
    uint32_t off = 8 + 8 * (uint32_t)cont.size();
    for (auto &c : cont) {
        put32(off);
        size_t card = c.second - c.first;
        off += card <= 4096 ? (uint32_t)card * 2 : 8192;
    }
    for (auto &c : cont) {
        size_t card = c.second - c.first;
        if (card <= 4096) {
            for (size_t k = c.first; k < c.second; k++) put16((uint16_t)(ids[k] & 0xffff));
        } else {
            uint64_t bits[1024];
            memset(bits, 0, sizeof bits);
            for (size_t k = c.first; k < c.second; k++) {
                uint32_t lo = ids[k] & 0xffff;
                bits[lo >> 6] |= 1ull << (lo & 63);
            }
            size_t at = out.size();
            out.resize(at + 8192);
            memcpy(out.data() + at, bits, 8192);
        }
    }
}

uint16_t bucketed_position(uint16_t rel) {  // crates/milli/src/lib.rs:248-260
    if (rel < 16) return rel;
    if (rel < 24) return 24;
    uint32_t p = 1;
    while (p < rel) p <<= 1;
    return (uint16_t)p;
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace

struct ig_builder {
    uint32_t n_fields;
    uint32_t exact_mask;
    std::unordered_set<std::string> stop_words;
    std::unordered_map<std::string, uint32_t> intern;
    std::vector<std::string> words;  // by intern id
    std::vector<Tok> toks;
    uint32_t max_doc_plus1 = 0;
    // synthetic docs (field 0) kept for query generation
    std::vector<uint32_t> syn_doc_off{0};
    std::vector<uint32_t> syn_doc_words;
    // built
    std::vector<uint8_t> dict_bytes;
    std::vector<uint64_t> dict_off;
    Db dbs[IG_DB_COUNT];
    std::vector<uint8_t> docids_cbo;
    std::vector<uint32_t> all_docs;

    uint32_t intern_word(const std::string &w) {
        auto it = intern.find(w);
        if (it != intern.end()) return it->second;
        uint32_t id = (uint32_t)words.size();
        words.push_back(w);
        intern.emplace(w, id);
        return id;
    }
};

// Thread count of the generators.  Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; a one-off 10 M-document build or
// a 15 GB embedding fill would then run on one core.  B200_GEN_THREADS wins; an OMP_NUM_THREADS other than 1 is respected; otherwise
// the host's threads are shared out over the ranks of this node (LOCAL_WORLD_SIZE).  The previous setting is restored on return.
struct GenThreads {
    int before;
    GenThreads() : before(omp_get_max_threads()) {
        int n = 0;
        if (const char *e = getenv("B200_GEN_THREADS")) n = atoi(e);
        if (n <= 0) {
            const char *o = getenv("OMP_NUM_THREADS");
            if (o && atoi(o) > 1)
                n = atoi(o);
            else {
                int hw = (int)std::thread::hardware_concurrency();
                const char *lw = getenv("LOCAL_WORLD_SIZE");
                int ranks = lw ? std::max(1, atoi(lw)) : 1;
                n = std::max(1, hw / ranks);
            }
        }
        omp_set_num_threads(n);
    }
    ~GenThreads() { omp_set_num_threads(before); }
};

extern "C" {

ig_builder *ig_new(uint32_t n_fields, uint32_t exact_mask) {
    auto *b = new ig_builder();
    b->n_fields = n_fields;
    b->exact_mask = exact_mask;
    return b;
}
void ig_free(ig_builder *b) { delete b; }

void ig_set_stop_words(ig_builder *b, const char *words) {
    std::string cur;
    for (const char *p = words;; p++) {
        if (*p == ' ' || *p == 0) {
            if (!cur.empty()) b->stop_words.insert(cur);
            cur.clear();
            if (!*p) break;
        } else
            cur.push_back(*p);
    }
}

static bool is_word_byte(unsigned char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c >= 0x80;
}

void ig_add_text(ig_builder *b, uint32_t docid, uint32_t fid, const char *text) {
    // Lowercase ASCII tokenizer standing in for charabia on plain Latin text.
    // Position rules: tokenize_document.rs:128-150 (first word +0, then +1, +8 after a hard separator).
    // stop words are matched on the token as written (the reference's stop-word set is case sensitive:
    // search/new/tests/stop_words.rs:1-10,25-29), everything else on the lowercased token
    std::string s(text);
    size_t i = 0, n = s.size();
    uint32_t pos = 0;
    bool first = true;
    bool hard = false;
    if (docid + 1 > b->max_doc_plus1) b->max_doc_plus1 = docid + 1;
    b->all_docs.push_back(docid);
    while (i < n) {
        if (!is_word_byte((unsigned char)s[i])) {
            // separator run: hard if it contains ". " ", " or one of ; ! ?
            size_t j = i;
            while (j < n && !is_word_byte((unsigned char)s[j])) {
                char c = s[j];
                if (c == ';' || c == '!' || c == '?') hard = true;
                if ((c == '.' || c == ',') && j + 1 < n && s[j + 1] == ' ') hard = true;
                j++;
            }
            i = j;
            continue;
        }
        size_t j = i;
        while (j < n && is_word_byte((unsigned char)s[j])) j++;
        std::string w = s.substr(i, j - i);
        const bool is_stop = b->stop_words.count(w) != 0;
        for (auto &c : w)
            if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
        i = j;
        if (first) {
            first = false;
        } else {
            pos += hard ? 8 : 1;
        }
        hard = false;
        if (pos >= 65536) break;
        if (w.size() > 250) continue;
        if (is_stop) continue;
        b->toks.push_back({b->intern_word(w), docid, (uint16_t)fid, (uint16_t)pos});
    }
}

static std::string random_word(Rng &r) {
    uint32_t len = 3 + r.below(10);
    std::string w(len, 'a');
    for (auto &c : w) c = (char)('a' + r.below(26));
    return w;
}

static std::string mutate(const std::string &w, Rng &r) {
    std::string o = w;
    uint32_t kind = r.below(4);
    if (o.size() < 2) kind = 1;
    switch (kind) {
        case 0: o[r.below((uint32_t)o.size())] = (char)('a' + r.below(26)); break;                 // sub
        case 1: o.insert(o.begin() + r.below((uint32_t)o.size() + 1), (char)('a' + r.below(26))); break;  // ins
        case 2: o.erase(o.begin() + r.below((uint32_t)o.size())); break;                             // del
        default: {
            uint32_t p = r.below((uint32_t)o.size() - 1);
            std::swap(o[p], o[p + 1]);
        }
    }
    return o;
}

void ig_add_synthetic(ig_builder *b, uint32_t n_docs, uint32_t vocab, double zipf_s, uint32_t len_lo,
                      uint32_t len_hi, uint64_t seed) {
    GenThreads gen_threads;
    Rng r(seed);
    // vocabulary: rank -> interned id
    std::vector<uint32_t> vid(vocab);
    std::unordered_set<std::string> seen;
    std::vector<std::string> vw;
    vw.reserve(vocab);
    for (uint32_t i = 0; i < vocab; i++) {
        std::string w;
        for (;;) {
            if (i > 100 && r.below(100) < 30) {
                w = mutate(vw[r.below(i)], r);
                if (w.size() < 3 || w.size() > 14) continue;
            } else
                w = random_word(r);
            if (seen.insert(w).second) break;
        }
        vw.push_back(w);
        vid[i] = b->intern_word(w);
    }
    // Zipf CDF
    std::vector<double> cdf(vocab);
    double acc = 0;
    for (uint32_t i = 0; i < vocab; i++) {
        acc += 1.0 / std::pow((double)(i + 1), zipf_s);
        cdf[i] = acc;
    }
    auto draw = [&]() -> uint32_t {
        double u = r.unit() * acc;
        uint32_t k = (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
        return k >= vocab ? vocab - 1 : k;
    };
    (void)draw;
    // Every document has its own generator (seeded by seed and docid), so the corpus does not depend on the number of threads:
    // pass 1 draws the field lengths, pass 2 fills the preallocated token arrays in parallel.
    uint32_t base = b->max_doc_plus1;
    const uint32_t nf = std::min<uint32_t>(b->n_fields, 2);
    auto doc_rng = [&](uint32_t d) { return Rng(seed ^ (0xD6E8FEB86659FD93ull * ((uint64_t)d + 1))); };
    std::vector<uint64_t> tok_off((size_t)n_docs + 1, 0), f0_off((size_t)n_docs + 1, 0);
#pragma omp parallel for schedule(static)
    for (uint32_t d = 0; d < n_docs; d++) {
        Rng rd = doc_rng(d);
        uint32_t l0 = len_lo + rd.below(len_hi - len_lo + 1), l1 = nf > 1 ? 20 + rd.below(61) : 0;
        tok_off[d + 1] = l0 + l1;
        f0_off[d + 1] = l0;
    }
    for (uint32_t d = 0; d < n_docs; d++) {
        tok_off[d + 1] += tok_off[d];
        f0_off[d + 1] += f0_off[d];
    }
    const size_t tok_base = b->toks.size(), f0_base = b->syn_doc_words.size();
    b->toks.resize(tok_base + tok_off[n_docs]);
    b->syn_doc_words.resize(f0_base + f0_off[n_docs]);
    b->syn_doc_off.resize(b->syn_doc_off.size() + n_docs);
    const size_t sdo_base = b->syn_doc_off.size() - n_docs;
    b->all_docs.resize(b->all_docs.size() + n_docs);
    const size_t ad_base = b->all_docs.size() - n_docs;
#pragma omp parallel for schedule(static)
    for (uint32_t d = 0; d < n_docs; d++) {
        Rng rd = doc_rng(d);
        const uint32_t doc = base + d;
        uint32_t lens[2] = {len_lo + rd.below(len_hi - len_lo + 1), nf > 1 ? 20 + rd.below(61) : 0};
        Tok *tk = b->toks.data() + tok_base + tok_off[d];
        uint32_t *sw = b->syn_doc_words.data() + f0_base + f0_off[d];
        for (uint32_t f = 0; f < nf; f++) {
            uint32_t pos = 0;
            for (uint32_t k = 0; k < lens[f]; k++) {
                double u = rd.unit() * acc;
                uint32_t rank = (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
                if (rank >= vocab) rank = vocab - 1;
                if (k > 0) pos += (rd.below(100) < 5) ? 8 : 1;
                *tk++ = Tok{vid[rank], doc, (uint16_t)f, (uint16_t)pos};
                if (f == 0) *sw++ = vid[rank];
            }
        }
        b->syn_doc_off[sdo_base + d] = (uint32_t)(f0_base + f0_off[d + 1]);
        b->all_docs[ad_base + d] = doc;
    }
    b->max_doc_plus1 = base + n_docs;
}

char *ig_synthetic_queries(ig_builder *b, uint32_t n, uint64_t seed, int with_typos) {
    Rng r(seed ^ 0xC0FFEEull);
    std::string out;
    uint32_t nd = (uint32_t)b->syn_doc_off.size() - 1;
    for (uint32_t q = 0; q < n; q++) {
        uint32_t d, len;
        do {
            d = r.below(nd);
            len = b->syn_doc_off[d + 1] - b->syn_doc_off[d];
        } while (len < 2);
        uint32_t want = 2 + r.below(3);
        if (want > len) want = len;
        uint32_t start = r.below(len - want + 1);
        for (uint32_t k = 0; k < want; k++) {
            std::string w = b->words[b->syn_doc_words[b->syn_doc_off[d] + start + k]];
            if (with_typos) {
                uint32_t e = r.below(100);
                uint32_t edits = e < 40 ? 0 : (e < 80 ? 1 : 2);
                for (uint32_t t = 0; t < edits; t++) {
                    std::string m = mutate(w, r);
                    if (!m.empty()) w = m;
                }
                if (k + 1 == want && r.below(100) < 30 && w.size() > 2) {
                    w = w.substr(0, 2 + r.below((uint32_t)w.size() - 2));
                }
            }
            if (k) out.push_back(' ');
            out += w;
        }
        out.push_back('\n');
    }
    char *p = (char *)malloc(out.size() + 1);
    memcpy(p, out.c_str(), out.size() + 1);
    return p;
}
void ig_free_str(char *p) { free(p); }

extern "C++" {
static void put_be16(std::vector<uint8_t> &k, uint16_t v) {
    k.push_back((uint8_t)(v >> 8));
    k.push_back((uint8_t)(v & 0xff));
}

// Emit a db from (key tuple -> sorted docs) given tuples sorted by (a, b, doc); key writer gets (a,b).
template <class KeyFn>
static void emit_db(Db &db, std::vector<std::pair<uint64_t, uint32_t>> &tuples, KeyFn key_fn) {
    __gnu_parallel::sort(tuples.begin(), tuples.end());
    tuples.erase(std::unique(tuples.begin(), tuples.end()), tuples.end());
    const size_t n = tuples.size();
    // the key ranges are encoded in parallel into per-thread databases and concatenated in order
    const int T = std::max(1, std::min<int>(omp_get_max_threads(), (int)(n / 65536) + 1));
    std::vector<size_t> cut(T + 1, n);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        size_t c = n * (size_t)t / (size_t)T;
        while (c < n && c > 0 && tuples[c].first == tuples[c - 1].first) c++;
        cut[t] = std::max(c, cut[t - 1]);
    }
    std::vector<Db> parts(T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        Db &part = parts[t];
        std::vector<uint32_t> ids;
        size_t i = cut[t];
        const size_t e = cut[t + 1];
        while (i < e) {
            size_t j = i;
            ids.clear();
            while (j < e && tuples[j].first == tuples[i].first) ids.push_back(tuples[j++].second);
            key_fn(tuples[i].first, part.keys);
            cbo_encode(ids.data(), ids.size(), part.vals);
            part.end_entry();
            i = j;
        }
    }
    for (auto &part : parts) {
        const uint64_t kb = db.keys.size(), vb = db.vals.size();
        db.keys.insert(db.keys.end(), part.keys.begin(), part.keys.end());
        db.vals.insert(db.vals.end(), part.vals.begin(), part.vals.end());
        for (size_t k = 1; k < part.koff.size(); k++) {
            db.koff.push_back(kb + part.koff[k]);
            db.voff.push_back(vb + part.voff[k]);
        }
        Db().keys.swap(part.keys);
        Db().vals.swap(part.vals);
    }
}
}  // extern "C++"

void ig_build(ig_builder *b) {
    GenThreads gen_threads;
    // 1. sorted dictionary, remap word ids to ranks
    size_t nw = b->words.size();
    std::vector<uint32_t> order(nw);
    for (uint32_t i = 0; i < nw; i++) order[i] = i;
    __gnu_parallel::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return b->words[x] < b->words[y]; });
    // only words that actually occur are in the dictionary
    std::vector<uint8_t> used(nw, 0);
    for (auto &t : b->toks) used[t.word] = 1;
    std::vector<uint32_t> rank(nw, 0xffffffffu);
    std::vector<const std::string *> sorted_words;
    for (uint32_t i = 0; i < nw; i++)
        if (used[order[i]]) {
            rank[order[i]] = (uint32_t)sorted_words.size();
            sorted_words.push_back(&b->words[order[i]]);
        }
    b->dict_off.assign(1, 0);
    for (auto *w : sorted_words) {
        b->dict_bytes.insert(b->dict_bytes.end(), w->begin(), w->end());
        b->dict_off.push_back(b->dict_bytes.size());
    }
    size_t W = sorted_words.size();
    auto wkey = [&](uint32_t r, std::vector<uint8_t> &k) { k.insert(k.end(), sorted_words[r]->begin(), sorted_words[r]->end()); };

    // documents ids
    std::sort(b->all_docs.begin(), b->all_docs.end());
    b->all_docs.erase(std::unique(b->all_docs.begin(), b->all_docs.end()), b->all_docs.end());
    cbo_encode(b->all_docs.data(), b->all_docs.size(), b->docids_cbo);

    // 2. word_docids / exact_word_docids / word_fid / word_position / fid_word_count
    std::vector<std::pair<uint64_t, uint32_t>> t_word, t_exact, t_fid, t_pos, t_cnt;
    // sort tokens by (doc, fid, pos) for pair extraction and counts
    {
        auto tok_less = [](const Tok &x, const Tok &y) {
            if (x.doc != y.doc) return x.doc < y.doc;
            if (x.fid != y.fid) return x.fid < y.fid;
            return x.pos < y.pos;
        };
        if (!std::is_sorted(b->toks.begin(), b->toks.end(), tok_less)) __gnu_parallel::stable_sort(b->toks.begin(), b->toks.end(), tok_less);
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < b->toks.size(); i++) b->toks[i].word = rank[b->toks[i].word];
    std::vector<std::pair<uint64_t, uint32_t>> t_pair;  // key = prox<<42 | w1<<21 | w2
    if (W >= (1u << 21)) {
        fprintf(stderr, "indexgen: vocabulary too large for packed pair keys\n");
        abort();
    }
    const int T = std::max(1, omp_get_max_threads());
    std::vector<std::vector<std::pair<uint64_t, uint32_t>>> p_word(T), p_exact(T), p_fid(T), p_pos(T), p_cnt(T), p_pair(T);
#pragma omp parallel num_threads(T)
    {
        const int tid = omp_get_thread_num();
        auto &t_word = p_word[tid];
        auto &t_exact = p_exact[tid];
        auto &t_fid = p_fid[tid];
        auto &t_pos = p_pos[tid];
        auto &t_cnt = p_cnt[tid];
        auto &t_pair = p_pair[tid];
        const size_t ntok = b->toks.size();
        // this thread's token range, moved to document boundaries
        size_t i = ntok * (size_t)tid / (size_t)T, n = ntok * (size_t)(tid + 1) / (size_t)T;
        while (i > 0 && i < ntok && b->toks[i].doc == b->toks[i - 1].doc) i++;
        while (n > 0 && n < ntok && b->toks[n].doc == b->toks[n - 1].doc) n++;
        std::vector<std::pair<uint64_t, uint8_t>> docpairs;  // (w1<<21|w2, prox)
        while (i < n) {
            size_t dj = i;
            while (dj < n && b->toks[dj].doc == b->toks[i].doc) dj++;
            docpairs.clear();
            size_t f0 = i;
            while (f0 < dj) {
                size_t f1 = f0;
                while (f1 < dj && b->toks[f1].fid == b->toks[f0].fid) f1++;
                uint32_t cnt = (uint32_t)(f1 - f0);
                uint16_t fid = b->toks[f0].fid;
                if (cnt <= 30) t_cnt.push_back({((uint64_t)fid << 8) | cnt, b->toks[i].doc});
                for (size_t a = f0; a < f1; a++) {
                    const Tok &ta = b->toks[a];
                    bool exact = (b->exact_mask >> fid) & 1;
                    (exact ? t_exact : t_word).push_back({ta.word, ta.doc});
                    t_fid.push_back({((uint64_t)ta.word << 16) | fid, ta.doc});
                    t_pos.push_back({((uint64_t)ta.word << 16) | bucketed_position(ta.pos), ta.doc});
                    for (size_t c = a + 1; c < f1; c++) {
                        uint32_t dist = (uint32_t)b->toks[c].pos - ta.pos;
                        if (dist >= 4) break;
                        if (dist > 0) docpairs.push_back({((uint64_t)ta.word << 21) | b->toks[c].word, (uint8_t)dist});
                    }
                }
                f0 = f1;
            }
            // keep the smallest proximity per ordered pair (sort + dedup_by key, :471-483)
            std::sort(docpairs.begin(), docpairs.end());
            for (size_t k = 0; k < docpairs.size(); k++) {
                if (k && docpairs[k].first == docpairs[k - 1].first) continue;
                t_pair.push_back({((uint64_t)docpairs[k].second << 42) | docpairs[k].first, b->toks[i].doc});
            }
            i = dj;
        }
    }
    auto gather = [&](std::vector<std::vector<std::pair<uint64_t, uint32_t>>> &parts, std::vector<std::pair<uint64_t, uint32_t>> &out) {
        size_t tot = 0;
        std::vector<size_t> at(parts.size() + 1, 0);
        for (size_t t = 0; t < parts.size(); t++) at[t + 1] = (tot += parts[t].size());
        out.resize(tot);
#pragma omp parallel for schedule(static, 1)
        for (size_t t = 0; t < parts.size(); t++) {
            std::copy(parts[t].begin(), parts[t].end(), out.begin() + at[t]);
            std::vector<std::pair<uint64_t, uint32_t>>().swap(parts[t]);
        }
    };
    gather(p_word, t_word);
    gather(p_exact, t_exact);
    gather(p_fid, t_fid);
    gather(p_pos, t_pos);
    gather(p_cnt, t_cnt);
    gather(p_pair, t_pair);
    auto key_word = [&](uint64_t k, std::vector<uint8_t> &out) { wkey((uint32_t)k, out); };
    auto key_word_u16 = [&](uint64_t k, std::vector<uint8_t> &out) {
        wkey((uint32_t)(k >> 16), out);
        out.push_back(0);
        put_be16(out, (uint16_t)(k & 0xffff));
    };
    // keep decoded copies of word lists for the prefix dbs
    emit_db(b->dbs[IG_DB_WORD_DOCIDS], t_word, key_word);
    emit_db(b->dbs[IG_DB_EXACT_WORD_DOCIDS], t_exact, key_word);
    emit_db(b->dbs[IG_DB_WORD_FID_DOCIDS], t_fid, key_word_u16);
    emit_db(b->dbs[IG_DB_WORD_POSITION_DOCIDS], t_pos, key_word_u16);
    emit_db(b->dbs[IG_DB_FIELD_ID_WORD_COUNT_DOCIDS], t_cnt, [&](uint64_t k, std::vector<uint8_t> &out) {
        put_be16(out, (uint16_t)(k >> 8));
        out.push_back((uint8_t)(k & 0xff));
    });
    emit_db(b->dbs[IG_DB_WORD_PAIR_PROXIMITY_DOCIDS], t_pair, [&](uint64_t k, std::vector<uint8_t> &out) {
        out.push_back((uint8_t)(k >> 42));
        wkey((uint32_t)((k >> 21) & 0x1fffff), out);
        out.push_back(0);
        wkey((uint32_t)(k & 0x1fffff), out);
    });
    std::vector<std::pair<uint64_t, uint32_t>>().swap(t_pair);

    // 3. prefix dbs: every 1..4-byte prefix shared by >= 100 dictionary words (word_fst_builder.rs:100-131)
    struct Pfx {
        std::string p;
        uint32_t lo, hi;  // word rank range
    };
    std::vector<Pfx> pfx;
    for (uint32_t n = 1; n <= 4; n++) {
        size_t i = 0;
        while (i < W) {
            const std::string &w = *sorted_words[i];
            if (w.size() < n) {
                i++;
                continue;
            }
            // utf-8 boundary check: a prefix must end on a char boundary
            if (w.size() > n && ((unsigned char)w[n] & 0xC0) == 0x80) {
                i++;
                continue;
            }
            std::string p = w.substr(0, n);
            size_t j = i;
            while (j < W && sorted_words[j]->size() >= n && sorted_words[j]->compare(0, n, p) == 0) j++;
            if (j - i >= 100) pfx.push_back({p, (uint32_t)i, (uint32_t)j});
            i = j;
        }
    }
    std::sort(pfx.begin(), pfx.end(), [](const Pfx &a, const Pfx &c) { return a.p < c.p; });
    // union lists per prefix from the tuple arrays (already sorted by key)
    auto build_prefix = [&](std::vector<std::pair<uint64_t, uint32_t>> &tuples, int shift, Db &out, bool with_u16) {
        // tuples sorted by (word<<shift | x, doc)
        std::vector<std::pair<uint64_t, uint32_t>> pt;
        for (size_t pi = 0; pi < pfx.size(); pi++) {
            uint64_t lo = (uint64_t)pfx[pi].lo << shift, hi = (uint64_t)pfx[pi].hi << shift;
            auto it0 = std::lower_bound(tuples.begin(), tuples.end(), std::make_pair(lo, 0u));
            auto it1 = std::lower_bound(tuples.begin(), tuples.end(), std::make_pair(hi, 0u));
            for (auto it = it0; it != it1; ++it) {
                uint64_t x = shift ? (it->first & 0xffff) : 0;
                pt.push_back({((uint64_t)pi << 16) | x, it->second});
            }
        }
        emit_db(out, pt, [&](uint64_t k, std::vector<uint8_t> &o) {
            const std::string &p = pfx[k >> 16].p;
            o.insert(o.end(), p.begin(), p.end());
            if (with_u16) {
                o.push_back(0);
                put_be16(o, (uint16_t)(k & 0xffff));
            }
        });
    };
    build_prefix(t_word, 0, b->dbs[IG_DB_WORD_PREFIX_DOCIDS], false);
    build_prefix(t_exact, 0, b->dbs[IG_DB_EXACT_WORD_PREFIX_DOCIDS], false);
    build_prefix(t_fid, 16, b->dbs[IG_DB_WORD_PREFIX_FID_DOCIDS], true);
    build_prefix(t_pos, 16, b->dbs[IG_DB_WORD_PREFIX_POSITION_DOCIDS], true);
    std::vector<Tok>().swap(b->toks);
}

// Query generation from a persisted corpus (the arrays ig_query_source exposes): same stream as ig_synthetic_queries.
static char *queries_from_arrays(const uint8_t *word_bytes, const uint64_t *word_off, const uint32_t *doc_off, uint32_t n_docs,
                                 const uint32_t *doc_words, uint32_t n, uint64_t seed, int with_typos) {
    Rng r(seed ^ 0xC0FFEEull);
    std::string out;
    for (uint32_t q = 0; q < n; q++) {
        uint32_t d, len;
        do {
            d = r.below(n_docs);
            len = doc_off[d + 1] - doc_off[d];
        } while (len < 2);
        uint32_t want = 2 + r.below(3);
        if (want > len) want = len;
        uint32_t start = r.below(len - want + 1);
        for (uint32_t k = 0; k < want; k++) {
            uint32_t wid = doc_words[doc_off[d] + start + k];
            std::string w((const char *)word_bytes + word_off[wid], (size_t)(word_off[wid + 1] - word_off[wid]));
            if (with_typos) {
                uint32_t e = r.below(100);
                uint32_t edits = e < 40 ? 0 : (e < 80 ? 1 : 2);
                for (uint32_t t = 0; t < edits; t++) {
                    std::string m = mutate(w, r);
                    if (!m.empty()) w = m;
                }
                if (k + 1 == want && r.below(100) < 30 && w.size() > 2) w = w.substr(0, 2 + r.below((uint32_t)w.size() - 2));
            }
            if (k) out.push_back(' ');
            out += w;
        }
        out.push_back('\n');
    }
    char *p = (char *)malloc(out.size() + 1);
    memcpy(p, out.c_str(), out.size() + 1);
    return p;
}
char *ig_queries_from_arrays(const uint8_t *word_bytes, const uint64_t *word_off, const uint32_t *doc_off, uint32_t n_docs,
                             const uint32_t *doc_words, uint32_t n, uint64_t seed, int with_typos) {
    return queries_from_arrays(word_bytes, word_off, doc_off, n_docs, doc_words, n, seed, with_typos);
}
// The arrays ig_queries_from_arrays needs, for the on-disk cache of a synthetic corpus: interned words (by intern id) and the
// field-0 word ids of every synthetic document.  The byte/offset buffers are malloc'ed; free them with ig_free_str.
void ig_query_source(const ig_builder *b, uint8_t **word_bytes, uint64_t **word_off, uint64_t *n_words, const uint32_t **doc_off,
                     uint64_t *n_docs, const uint32_t **doc_words, uint64_t *n_doc_words) {
    uint64_t tot = 0;
    for (auto &w : b->words) tot += w.size();
    uint8_t *wb = (uint8_t *)malloc(tot + 1);
    uint64_t *wo = (uint64_t *)malloc((b->words.size() + 1) * 8);
    uint64_t at = 0;
    for (size_t i = 0; i < b->words.size(); i++) {
        wo[i] = at;
        memcpy(wb + at, b->words[i].data(), b->words[i].size());
        at += b->words[i].size();
    }
    wo[b->words.size()] = at;
    *word_bytes = wb;
    *word_off = wo;
    *n_words = b->words.size();
    *doc_off = b->syn_doc_off.data();
    *n_docs = b->syn_doc_off.size() - 1;
    *doc_words = b->syn_doc_words.data();
    *n_doc_words = b->syn_doc_words.size();
}

// cfg 4 embeddings: rows i.i.d. N(0,1) then L2-normalised, stored fp16.  Row r has its own generator (seed, r), so
// any row range can be produced by any number of threads; `out` holds n x d IEEE binary16 values for rows [first_row, first_row+n).
void ig_fill_embeddings_f16(uint16_t *out, uint64_t first_row, uint64_t n, uint32_t d, uint64_t seed) {
    GenThreads gen_threads;
#pragma omp parallel
    {
        std::vector<float> v(d + 1);
#pragma omp for schedule(static)
        for (uint64_t r = 0; r < n; r++) {
            Rng g(seed ^ (0xA24BAED4963EE407ull * (first_row + r + 1)));
            double ss = 0;
            for (uint32_t i = 0; i < d; i += 2) {  // Box-Muller, two values per pair of uniforms
                double u1 = 1.0 - g.unit(), u2 = g.unit();
                double rad = std::sqrt(-2.0 * std::log(u1)), ang = 6.283185307179586 * u2;
                v[i] = (float)(rad * std::cos(ang));
                v[i + 1] = (float)(rad * std::sin(ang));
            }
            for (uint32_t i = 0; i < d; i++) ss += (double)v[i] * v[i];
            const float inv = ss > 0 ? (float)(1.0 / std::sqrt(ss)) : 0.f;
            _Float16 *o = reinterpret_cast<_Float16 *>(out + r * d);
            for (uint32_t i = 0; i < d; i++) o[i] = (_Float16)(v[i] * inv);
        }
    }
}

uint32_t ig_n_docs(const ig_builder *b) { return b->max_doc_plus1; }
uint64_t ig_n_words(const ig_builder *b) { return b->dict_off.size() - 1; }
void ig_dictionary(const ig_builder *b, const uint8_t **bytes, const uint64_t **offsets) {
    *bytes = b->dict_bytes.data();
    *offsets = b->dict_off.data();
}
void ig_db(const ig_builder *b, int id, ig_db_view *out) {
    const Db &d = b->dbs[id];
    out->n_keys = d.koff.size() - 1;
    out->key_bytes = d.keys.data();
    out->key_offsets = d.koff.data();
    out->val_bytes = d.vals.data();
    out->val_offsets = d.voff.data();
}
void ig_documents_ids(const ig_builder *b, const uint8_t **bytes, uint64_t *len) {
    *bytes = b->docids_cbo.data();
    *len = b->docids_cbo.size();
}
}
