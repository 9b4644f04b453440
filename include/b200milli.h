/* b200milli — B200-native (sm_100a) implementation of milli's query-time scoring path.
 *
 * C ABI of the drop-in boundary.  The reference has no FFI for this path; its
 * seams are Rust-internal.  Each entry point names the reference interface a Rust shim would
 * replace with a call to it (paths relative to the meilisearch checkout, v1.50.0 @ 5cb2f2e).
 * Conventions: caller allocates every output; the library never frees caller memory; handles are
 * opaque; every function returns 0 on success or a negative B200_ERR_* code, with a message
 * available from b200_last_error(); no exceptions/panics cross the boundary; one handle may be
 * used from many threads (calls are serialised per handle; each call runs on the handle's stream).
 * There is NO CPU fallback: without a CUDA device b200_open fails with B200_ERR_NO_DEVICE.
 */
#ifndef B200MILLI_H
#define B200MILLI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_NO_DEVICE (-1)   /* no CUDA device / driver (the product never falls back to the CPU) */
#define B200_ERR_CUDA (-2)        /* a CUDA call failed; see b200_last_error */
#define B200_ERR_INVALID (-3)     /* bad argument */
#define B200_ERR_UNSUPPORTED (-4) /* query feature outside the implemented scope (sort, filters, more than 12 terms, ...) */
#define B200_ERR_CAPACITY (-5)    /* a device work queue / arena overflowed */
#define B200_ERR_STATE (-6)       /* call order (e.g. search before b200_stage_finish) */

typedef struct b200_index b200_index;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Replaces: opening the LMDB env for search (crates/milli/src/index.rs:129-…) — here a device-resident copy. */
int b200_open(int device_ordinal, b200_index **out);
void b200_close(b200_index *);
const char *b200_last_error(const b200_index *); /* valid until the next call on the handle */
/* message of the last failure of b200_open itself */
const char *b200_open_error(void);

/* ---- staging: done once, copies everything (LMDB pages are only borrowed for a RoTxn) --- */
/* database ids, in the key/value formats of crates/milli/src/index.rs:97-124 and
 * heed_codec/{str_beu32_codec.rs (StrBEU16Codec), str_str_u8_codec.rs (U8StrStrCodec)};
 * values are CboRoaringBitmapCodec bytes (heed_codec/roaring_bitmap/cbo_roaring_bitmap_codec.rs:15-85). */
enum b200_db {
    B200_DB_WORD_DOCIDS = 0,               /* key: word */
    B200_DB_EXACT_WORD_DOCIDS = 1,         /* key: word */
    B200_DB_WORD_PREFIX_DOCIDS = 2,        /* key: prefix */
    B200_DB_EXACT_WORD_PREFIX_DOCIDS = 3,  /* key: prefix */
    B200_DB_WORD_PAIR_PROXIMITY_DOCIDS = 4,/* key: u8 prox | w1 | 0x00 | w2 */
    B200_DB_WORD_POSITION_DOCIDS = 5,      /* key: word | 0x00 | u16 BE bucketed position */
    B200_DB_WORD_FID_DOCIDS = 6,           /* key: word | 0x00 | u16 BE field id */
    B200_DB_WORD_PREFIX_POSITION_DOCIDS = 7,
    B200_DB_WORD_PREFIX_FID_DOCIDS = 8,
    B200_DB_FIELD_ID_WORD_COUNT_DOCIDS = 9,/* key: u16 BE fid | u8 count */
    B200_DB_COUNT = 10
};
/* Replaces Index::words_fst (index.rs:1238): the FST enumerated once on the host into its sorted word list. */
int b200_stage_dictionary(b200_index *, const uint8_t *word_bytes, const uint64_t *word_offsets, uint64_t n_words);
/* Replaces the typed `Database<..., CboRoaringBitmapCodec>` handles read by search/new/db_cache.rs:183-719.
 * Keys must be in LMDB (bytewise) order. */
int b200_stage_db(b200_index *, int db, uint64_t n_keys, const uint8_t *key_bytes, const uint64_t *key_offsets,
                  const uint8_t *val_bytes, const uint64_t *val_offsets);
/* Replaces Index::documents_ids (index.rs, main["documents-ids"]); CBO bytes. */
int b200_stage_documents_ids(b200_index *, const uint8_t *cbo, uint64_t len);

/* criteria (crates/milli/src/criterion.rs:121-131) */
enum b200_criterion { B200_C_WORDS = 0, B200_C_TYPO = 1, B200_C_PROXIMITY = 2, B200_C_ATTRIBUTE = 3, B200_C_ATTRIBUTE_RANK = 4,
                      B200_C_WORD_POSITION = 5, B200_C_SORT = 6, B200_C_EXACTNESS = 7 };
typedef struct {
    uint32_t n_fields;              /* searchable fields; fid = 0..n_fields-1 */
    const uint16_t *weights;        /* fieldids_weights_map: fid -> weight */
    const int32_t *criteria;        /* index `criteria` setting */
    uint32_t n_criteria;
    int32_t authorize_typos;        /* index.rs authorize-typos */
    uint32_t min_word_len_one_typo; /* index.rs:46 (5) */
    uint32_t min_word_len_two_typos;/* index.rs:47 (9) */
    int32_t prefix_search;          /* PrefixSearch::IndexingTime (1) / Disabled (0) */
    const char *exact_words;        /* '\n'-joined exact_words set (may be NULL) */
} b200_settings;
int b200_stage_settings(b200_index *, const b200_settings *);
/* The index `synonyms` database (crates/milli/src/index.rs synonyms; read by compute_derivations.rs:221-239 and
 * parse_query.rs:277-285): entry i maps the word sequence from_words[i] to the word sequence to_words[i], both already
 * tokenised and joined by single spaces.  Several entries may share the same `from`.  Replaces any previous set. */
int b200_stage_synonyms(b200_index *, uint32_t n, const char *const *from_words, const char *const *to_words);
/* Uploads everything to HBM and builds the device directories.  Must follow the stage_* calls. */
int b200_stage_finish(b200_index *);
/* Replaces the arroy/hannoy item nodes read by VectorStore (crates/milli/src/vector/store.rs:1427-1434):
 * one f32[d] vector per row + its docid.  Stored on device as fp16 rows + f32 inverse norms. */
int b200_stage_embeddings(b200_index *, const float *vectors, uint64_t n, uint32_t d, const uint32_t *docids);
/* Same store from rows that are already IEEE binary16 (a quantised store, or a corpus generated as fp16): no f32 round trip;
 * the inverse norms are those of the fp16 values. */
int b200_stage_embeddings_f16(b200_index *, const uint16_t *rows_fp16, uint64_t n, uint32_t d, const uint32_t *docids);
/* Embedder `distribution` (crates/milli/src/vector/distribution.rs): enabled=0 disables the shift. */
int b200_stage_distribution(b200_index *, int enabled, float mean, float sigma);

/* ---- S3: term derivation --------------------------------------------------------------- */
/* Replaces Interned<QueryTerm>::compute_fully_if_needed -> find_one_typo_derivations /
 * find_one_two_typo_derivations (crates/milli/src/search/new/query_term/compute_derivations.rs:21-168),
 * i.e. `fst.search_with_state(Intersection/Union(StartsWith, LevenshteinDFA))`.
 * For word i (bytes words[word_off[i]..word_off[i+1]]), max_typo[i] in {1,2}, is_prefix[i] in {0,1}:
 * one_out[i*150..] gets up to 150 dictionary ranks at distance 1, two_out[i*50..] up to 50 at distance 2
 * (ascending rank order, caps and first-letter rule exactly as the reference). */
#define B200_MAX_ONE_TYPO 150
#define B200_MAX_TWO_TYPOS 50
int b200_derive_batch(b200_index *, uint32_t n_words, const char *words, const uint32_t *word_off, const uint8_t *max_typo,
                      const uint8_t *is_prefix, uint32_t *one_out, uint32_t *n_one, uint32_t *two_out, uint32_t *n_two);

/* ---- S2: condition resolver (union-shaped conditions) ---------------------------------- */
/* Replaces the posting-list part of G::resolve_condition as called by ConditionDocIdsCache::get_computed_condition
 * (crates/milli/src/search/new/ranking_rule_graph/condition_docids_cache.rs:34-57; the unions are
 * compute_query_term_subset_docids, crates/milli/src/search/new/resolve_query_graph.rs:33-59): out = (OR of the posting lists of
 * `db` at key indices key_index[0..n_keys)) AND universe.  db: the b200_stage_db id; a key index is the position of the key in the
 * database as staged (LMDB order).  universe: dense little-endian u64 words over docids, NULL = all documents; out: the same
 * shape, ceil((max docid + 1) / 64) words. */
int b200_union_postings(b200_index *, int db, const uint32_t *key_index, uint32_t n_keys, const uint64_t *universe, uint64_t n_universe_words,
                        uint64_t *out);

/* Replaces ProximityGraph::resolve_condition for one edge (crates/milli/src/search/new/ranking_rule_graph/proximity/compute_docids.rs:15-108,
 * the non-prefix lookups :172-211): out = universe AND the union, over every l in `left` and r in `right` (dictionary ranks, i.e.
 * positions in the staged dictionary), of word_pair_proximity_docids[(fwd_prox, l, r)] and word_pair_proximity_docids[(bwd_prox, r, l)];
 * a proximity of 0 disables that direction.  universe / out as in b200_union_postings. */
int b200_proximity_pairs(b200_index *, const uint32_t *left, uint32_t n_left, const uint32_t *right, uint32_t n_right, uint32_t fwd_prox,
                         uint32_t bwd_prox, const uint64_t *universe, uint64_t n_universe_words, uint64_t *out);

/* ---- S4: vector store ------------------------------------------------------------------ */
/* Replaces VectorStore::nns_by_vector (crates/milli/src/vector/store.rs:638-675) for a batch of queries:
 * exact scan, ascending distance (1 - cos)/2, ties by ascending docid.
 * queries: n_q x d f32 (host).  cand_bitmap: optional dense little-endian u64 words over docids (filter),
 * shared by the batch (NULL = all).  ids_out/dist_out: n_q x limit; n_out: n_q. */
int b200_nns_batch(b200_index *, const float *queries, uint32_t n_q, uint32_t d, uint32_t limit, const uint64_t *cand_bitmap,
                   uint64_t n_cand_words, uint32_t *ids_out, float *dist_out, uint32_t *n_out);

/* ---- corpus partitioned across GPUs (cfg 5) --------------------------------- */
/* One process per GPU, each with the rows of its docid range staged (b200_stage_embeddings with global docids).  The library
 * binds NCCL at run time (libnccl.so.2).  Rank 0 draws a unique id, the host application carries its 128 bytes to the other
 * ranks, every rank calls b200_comm_init. */
int b200_comm_unique_id(b200_index *, uint8_t *out128);
int b200_comm_init(b200_index *, int rank, int world, const uint8_t *unique_id128);
/* b200_nns_batch over the partitioned store: every rank passes the SAME queries (and candidate bitmap over global docids); each
 * scans its shard, the per-shard top-`limit` lists are exchanged with one ncclAllGather on the library's vector stream and merged
 * on the device by (distance, docid); every rank receives the global result. */
int b200_nns_batch_sharded(b200_index *, const float *queries, uint32_t n_q, uint32_t d, uint32_t limit, const uint64_t *cand_bitmap,
                           uint64_t n_cand_words, uint32_t *ids_out, float *dist_out, uint32_t *n_out);

/* ---- S0: whole search ------------------------------------------------------------------ */
/* Replaces milli::Search::execute / execute_hybrid (crates/milli/src/search/mod.rs:280-415,
 * search/hybrid.rs:264-366) for a batch of queries against one index, as called from
 * search_from_kind (crates/meilisearch/src/search/mod.rs:2126-2148) / SearchByIndex::execute
 * (crates/meilisearch/src/search/federated/perform.rs:1544).
 * Queries arrive tokenised (charabia stays on the host): tokens of query i are
 * [token_begin[i], token_begin[i+1]); kind: 0 Word, 1 StopWord, 2 Separator(Soft), 3 Separator(Hard). */
enum b200_tms { B200_TMS_LAST = 0, B200_TMS_ALL = 1, B200_TMS_FREQUENCY = 2 };
typedef struct {
    uint32_t n_queries;
    const uint32_t *token_begin;  /* n_queries + 1 */
    const uint8_t *token_kind;
    const uint32_t *lemma_off;    /* n_tokens + 1 */
    const char *lemma_bytes;
    int32_t terms_matching_strategy; /* b200_tms */
    int32_t scoring_strategy;        /* 0 Skip, 1 Detailed (score_details.rs:431-438) */
    uint32_t offset, limit;          /* Search::offset / limit */
    uint32_t words_limit;            /* Search::words_limit (default 10) */
    const float *vectors;            /* n_queries x d, or NULL: the `semantic` vector per query */
    int32_t mode;                    /* 0 keyword (execute), 1 semantic (execute with vector), 2 hybrid (execute_hybrid) */
    float semantic_ratio;            /* hybrid only */
    /* filtered_universe (search/new/mod.rs:719: documents_ids & filter), what Search::filter / candidates produce on the host:
     * n_queries pointers to dense little-endian u64 words over docids (n_universe_words each), NULL entry = all documents,
     * NULL array = no filter anywhere.  Queries may share a bitmap (equal pointers are uploaded once). */
    const uint64_t *const *universes;
    uint64_t n_universe_words;
    /* Search::deadline (crates/milli/src/lib.rs:154-226).  time_budget_ns > 0: Deadline::from_budget, counted from the start of
     * the call; 0: Deadline::never.  stop_after >= 0: the reference's poll-count hook (Deadline::with_stop_after(n): exceeded from
     * the n-th poll on, the clock is then ignored); -1: unused.  When the deadline is exceeded the remaining universe of every
     * rule is returned unsorted with a Skipped score and the result is marked degraded (bucket_sort.rs:206-264). */
    uint64_t time_budget_ns;
    int64_t stop_after;
    /* Search::ranking_score_threshold (bucket_sort.rs:188,221-224,293-296) */
    int32_t has_ranking_score_threshold;
    double ranking_score_threshold;
} b200_query_batch;
#define B200_MAX_SCORES 12
/* score kinds: ScoreDetails variants (score_details.rs:9-32) */
enum b200_score_kind { B200_S_WORDS = 0, B200_S_TYPO = 1, B200_S_PROXIMITY = 2, B200_S_FID = 3, B200_S_POSITION = 4,
                       B200_S_EXACT_ATTRIBUTE = 5, B200_S_EXACT_WORDS = 6, B200_S_VECTOR = 7, B200_S_SKIPPED = 8 };
typedef struct {                  /* SearchResult (search/mod.rs:526-535), flattened; all caller-allocated */
    uint32_t *docids;             /* n_queries x limit      documents_ids */
    uint32_t *n_hits;             /* n_queries */
    uint8_t *n_scores;            /* n_queries x limit      len of document_scores[i] (Detailed only) */
    uint8_t *score_kind;          /* n_queries x limit x B200_MAX_SCORES */
    uint32_t *score_rank;         /* idem: Rank.rank  (score_details.rs:512-522) */
    uint32_t *score_max;          /* idem: Rank.max_rank */
    float *score_sim;             /* idem: Vector.similarity, -1 when None */
    uint64_t *n_candidates;       /* n_queries: candidates.len() */
    uint32_t *semantic_hits;      /* n_queries: execute_hybrid's semantic_hit_count (may be NULL) */
    int32_t *status;              /* n_queries: 0 or a B200_ERR_* for that query (e.g. UNSUPPORTED) */
    uint8_t *degraded;            /* n_queries: SearchResult::degraded (may be NULL) */
    uint8_t *used_negative_operator; /* n_queries: SearchResult::used_negative_operator (may be NULL) */
    uint64_t *candidates;         /* optional (may be NULL): n_queries x candidates_words dense u64 words, SearchResult::candidates
                                     for keyword searches without a ranking-score threshold (others: B200_ERR_UNSUPPORTED) */
    uint64_t candidates_words;    /* words per query in `candidates` (>= ceil((max docid + 1) / 64)) */
} b200_results;
int b200_search_batch(b200_index *, const b200_query_batch *, b200_results *);

/* ---- S1: the RankingRule seam ---------------------------------------------------------- */
/* Replaces `dyn RankingRule` as driven by bucket_sort (crates/milli/src/search/new/ranking_rules.rs:26-83, bucket_sort.rs:123,266,323)
 * for the graph-based rules and ExactAttribute.  Query graphs are opaque library objects (a QueryGraph plus the terms it refers
 * to): the first one comes from b200_graph_from_tokens (QueryGraph::from_query, query_graph.rs:96-187, with every term's
 * derivations computed), the next ones from b200_rule_next (RankingRuleOutput::query: the graph rebuilt from the paths that
 * produced the bucket, graph_based_ranking_rule.rs:340-353).  Every graph handed out must be released with b200_graph_free. */
typedef struct b200_graph b200_graph;
typedef struct b200_rule b200_rule;
/* one_query: a b200_query_batch with n_queries == 1 (tokens, words_limit, terms_matching_strategy) */
int b200_graph_from_tokens(b200_index *, const b200_query_batch *one_query, b200_graph **out);
void b200_graph_free(b200_graph *);
/* RankingRule::start_iteration(universe, query).  rule_kind: B200_S_WORDS, _TYPO, _PROXIMITY, _FID, _POSITION, _EXACT_ATTRIBUTE,
 * _EXACT_WORDS (= Exactness); terms_matching_strategy matters for B200_S_WORDS only.  universe: dense u64 words, NULL = all
 * documents.  All buckets of the rule are evaluated here, in one device step. */
int b200_rule_start(b200_index *, int rule_kind, int terms_matching_strategy, const b200_graph *query, const uint64_t *universe,
                    uint64_t n_universe_words, b200_rule **out);
/* RankingRule::next_bucket(universe): returns 0 and the next bucket in ascending cost order (empty ones included, like the
 * reference, graph_based_ranking_rule.rs:231-236), or 1 when the rule is exhausted (None).  out_bitmap (n_words words) =
 * RankingRuleOutput::candidates = bucket AND universe (NULL = the start universe); rank / max_rank = the bucket's score
 * (score_details.rs Rank); out_query = RankingRuleOutput::query (NULL for an empty bucket; caller frees). */
int b200_rule_next(b200_rule *, const uint64_t *universe, uint64_t *out_bitmap, uint64_t n_words, uint32_t *rank, uint32_t *max_rank,
                   b200_graph **out_query);
/* RankingRule::end_iteration */
void b200_rule_end(b200_rule *);

/* ---- introspection for measurement ----------------------------------------------------- */
/* kernel classes for the per-kernel accounting below */
enum b200_kernel { B200_K_LEV = 0, B200_K_COMPACT = 1, B200_K_PAIR_PROBE = 2, B200_K_SCATTER = 3, B200_K_EVAL_PATHS = 4, B200_K_EMIT = 5,
                   B200_K_VEC_DIST = 6, B200_K_TOPK = 7, B200_K_VEC_GEMM = 8, B200_K_VEC_MERGE = 9, B200_K_COUNT = 10 };
typedef struct {
    uint64_t kernel_launches;     /* kernels launched by the library since the last reset */
    uint64_t device_steps;        /* host<->device round trips since the last reset */
    uint64_t posting_bytes;       /* algorithmic bytes: stored bytes of posting lists read */
    uint64_t matrix_bytes;        /* algorithmic bytes: condition/bucket matrix words read+written */
    uint64_t dictionary_bytes;    /* algorithmic bytes of the term-derivation sweeps */
    uint64_t vector_bytes;        /* algorithmic bytes of the distance scans */
    double kernel_ms[10];         /* CUDA-event time accumulated per kernel class (events on the library's stream) */
    uint64_t kernel_count[10];    /* launches per kernel class */
    uint64_t kernel_bytes[10];    /* algorithmic bytes attributed to each kernel class */
    double device_ms;             /* CUDA-event time from the first to the last kernel of every step */
    uint64_t h2d_bytes, d2h_bytes; /* bytes copied across PCIe/NVLink-C2C by search/derive/nns calls */
    double host_ms[8];            /* wall time of the host phases of b200_search_batch: 0 parse, 1 derive (incl. device), 2 term finalisation,
                                     3 step packing, 4 step device wait, 5 bucket-sort advance, 6 result copy, 7 total */
    uint64_t hbm_bytes_staged;
    uint64_t deferred;            /* ranking-rule activations that had to wait for a later device step (scratch / arena full) */
    uint64_t arena_peak_bytes;    /* high-water mark of the per-batch level storage (universes + bucket columns) */
    uint64_t eval_class_launches[9]; /* eval_dp launches per DP-table class: <= 16 / 24 / 40 / 56 / 80 / 112 / 160 / 216 slots in shared memory, [8] = global matrices */
    uint64_t eval_class_tiles[9];    /* 128-row tiles evaluated per class */
} b200_stats;
int b200_get_stats(b200_index *, b200_stats *out);
int b200_reset_stats(b200_index *);

#ifdef __cplusplus
}
#endif
#endif
