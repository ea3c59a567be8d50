/*
 * ottocov.h -- C ABI of libottocov.so, the B200-native co-visitation (co-event) counting engine.
 *
 * The reference (nicolaivicol/otto-recommender) has no FFI: its stage boundary is the Python
 * module model/count_co_events.py plus parquet files.  Each entry point below names the
 * reference code it replaces (file:line into the reference tree); INTEGRATION.md shows the
 * ctypes stub a maintainer would add to model/count_co_events.py.
 *
 * Conventions
 *   - every function returns 0 (OTTOCOV_OK) or a negative ottocov_status; nothing throws;
 *     ottocov_last_error(ctx) gives the message of the last failure on that context.
 *   - plain pointers and sizes only.  `where` says whether a caller buffer is host
 *     (OTTOCOV_HOST) or device (OTTOCOV_DEVICE) memory; the caller owns every buffer it passes.
 *   - the library owns its scratch and every ottocov_table; one context per device, one thread
 *     per context.  All work is enqueued on the context's stream (ottocov_set_stream); only
 *     the *_fetch / *_info calls and calls that must size an allocation synchronise it.
 *   - there is no CPU fallback: without a CUDA device ottocov_create fails with
 *     OTTOCOV_ERR_CUDA.
 *
 * Key format of a table row: key = (uint64)aid << 32 | (uint32)aid_next, count = uint32.
 * Rows of a table are always sorted by key and keys are distinct.
 */
#ifndef OTTOCOV_H
#define OTTOCOV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define OTTOCOV_VERSION 200

typedef enum {
    OTTOCOV_OK = 0,
    OTTOCOV_ERR_CUDA = -1,        /* a CUDA runtime call failed (message has the CUDA error)   */
    OTTOCOV_ERR_ARG = -2,         /* bad argument (NULL, negative size, k out of range ...)     */
    OTTOCOV_ERR_DATA = -3,        /* input violates the schema (type not in 0..2, aid < 0 ...)  */
    OTTOCOV_ERR_STATE = -4,       /* call order wrong (count before load_events ...)            */
    OTTOCOV_ERR_CAPACITY = -5,    /* caller buffer too small                                    */
    OTTOCOV_ERR_NOMEM = -6        /* device allocation failed                                   */
} ottocov_status;

enum { OTTOCOV_HOST = 0, OTTOCOV_DEVICE = 1 };
enum { OTTOCOV_ORDER_KEY = 0,          /* (aid, aid_next) ascending                              */
       OTTOCOV_ORDER_COUNT_DESC = 1 }; /* count desc, then aid asc, aid_next asc (file order)   */

typedef struct ottocov_ctx ottocov_ctx;
typedef struct ottocov_table ottocov_table;

/* One co-event kind = one entry of config.MAP_NAME_COUNT_TYPE + MAP_MAX_TIME_TO_NEXT
 * (reference config.py:43-49, 81-88). */
enum { OTTOCOV_SYM_OFF = 1, OTTOCOV_SYM_ON = 2,
       OTTOCOV_HASH_OFF = 4, OTTOCOV_HASH_ON = 8,     /* see ottocov_spec.flags */
       OTTOCOV_DT_RANGE = 16 };                       /* dt_min / dt_max of the spec are set */
/* bits of ottocov_reduce_pairs' strip_dest argument above bit 0 */
enum { OTTOCOV_REDUCE_HASH_ON = 2, OTTOCOV_REDUCE_HASH_OFF = 4 };

typedef struct {
    int32_t type_this;       /* source event type: 0 click, 1 cart, 2 order                       */
    uint32_t next_mask;      /* bit t set <=> events of type t are "next" events                  */
    int64_t window;          /* |ts_next - ts| <= window, seconds, inclusive (config.MAP_MAX_TIME_TO_NEXT); */
                             /* intersected with the pre-filter dt_min <= ts_next - ts <= dt_max     */
    int64_t pair_budget;     /* max co-event pairs expanded at once (HBM footprint); 0 = auto     */
    uint32_t min_count;      /* keep only pairs with count >= min_count (0 or 1 = keep all); fused */
                             /* into the run-length reduce: filter(count >= ...) of :131-132, :172  */
    uint32_t flags;          /* 0 = auto.  Kinds whose source and target type agree are symmetric  */
                             /* (count(a,b) == count(b,a)): the engine may expand each unordered    */
                             /* event pair once and mirror the reduced table.  Auto does so when    */
                             /* min_count > 1.  OTTOCOV_SYM_OFF / OTTOCOV_SYM_ON force the choice.   */
                             /* Reduce-by-key strategy: full radix sort + run-length reduce, or the  */
                             /* bucketed hash reduce (keys written through a bijective mix, sorted   */
                             /* on the top hash bits only, counted in shared-memory tables).  Auto   */
                             /* takes the hash reduce when min_count > 1 (few rows left to bring     */
                             /* back into key order); OTTOCOV_HASH_OFF / OTTOCOV_HASH_ON force it.    */
                             /* Both produce the same table bit for bit.                              */
    int64_t dt_min, dt_max;  /* the pre-filter of self_merge (count_co_events.py:33-36):               */
                             /* config.MIN_TIME_TO_NEXT <= ts_next - ts <= config.MAX_TIME_TO_NEXT.     */
                             /* Read only when flags has OTTOCOV_DT_RANGE; otherwise -86400 / +86400    */
                             /* (config.py:41-42).  An asymmetric range switches the symmetric shortcut */
                             /* off (count(a,b) != count(b,a) then).                                     */
} ottocov_spec;

typedef struct {
    int64_t n_rows_in;       /* rows handed to ottocov_load_events                                 */
    int64_t n_events;        /* after exact-duplicate removal (count_co_events.py:92)              */
    int64_t n_by_type[3];    /* deduplicated events per type                                       */
    int32_t session_min, session_max, ts_min, ts_max, aid_max;
    int32_t aid_bits;        /* significant bits of aid_max                                        */
    int32_t was_sorted;      /* input was already ordered by (session, ts): sort skipped           */
} ottocov_events_info;

typedef struct {
    int64_t n_pairs;         /* co-event pairs emitted by the last ottocov_count (sum of counts)   */
    int64_t n_unique;        /* rows of the table it produced                                      */
    int32_t n_chunks;        /* pair-budget chunks it ran                                          */
    int32_t sort_passes;     /* distribution passes per chunk (a pass fused into the expansion counts) */
    int32_t fused;           /* 1 = the first pass of the bucketed hash reduce ran inside the expansion;
                              * 2 = likewise, and whole buckets were counted per CTA (one pass fewer)     */
    int32_t reserved;
    int64_t h2d_bytes;       /* ottocov_count_parts: bytes copied host -> device (the session column travels  */
                             /* run-length encoded when it compresses); 0 for the other calls                */
} ottocov_count_info;

/* Per-kernel-family accounting for roofline reports (bench.py). */
enum { OTTOCOV_K_LOAD = 0, OTTOCOV_K_WINDOW, OTTOCOV_K_EXPAND, OTTOCOV_K_HIST, OTTOCOV_K_SORT_PASS,
       OTTOCOV_K_RLE, OTTOCOV_K_FILTER, OTTOCOV_K_TOPK, OTTOCOV_K_ORDER, OTTOCOV_K_PARTITION,
       OTTOCOV_K_MISC, OTTOCOV_K_FAMILIES };
typedef struct {
    int64_t launches;        /* kernel launches since the last reset                               */
    double ms;               /* summed CUDA-event time (only while profiling is on)                */
    double algo_bytes;       /* algorithmic bytes those launches moved            */
} ottocov_kernel_stat;

/* ---- context ------------------------------------------------------------------------------ */
int ottocov_version(void);
int ottocov_create(int device, ottocov_ctx** out);
int ottocov_destroy(ottocov_ctx* ctx);
const char* ottocov_last_error(const ottocov_ctx* ctx);      /* ctx may be NULL (create failures) */
int ottocov_set_stream(ottocov_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = default       */
int ottocov_synchronize(ottocov_ctx* ctx);
/* The context keeps freed device blocks for re-use (no driver allocation in steady state).
 * ottocov_trim hands them back to the driver; ottocov_memory_info reports the footprint. */
int ottocov_trim(ottocov_ctx* ctx);
int ottocov_memory_info(ottocov_ctx* ctx, int64_t* live_bytes, int64_t* cached_bytes, int64_t* peak_bytes);
int ottocov_set_profiling(ottocov_ctx* ctx, int on);         /* 0 off, 1 all families, else a bit  */
                                                             /* mask (1 << family): CUDA events    */
int ottocov_kernel_stats(ottocov_ctx* ctx, ottocov_kernel_stat* out /*[OTTOCOV_K_FAMILIES]*/, int reset);
const char* ottocov_kernel_family_name(int family);

/* ---- (1) loader: replaces pl.read_parquet(...).unique() and the session grouping the self-join
 * does implicitly (model/count_co_events.py:91-92, :19, :41-57).  Rows may come in any order.
 * Sorts by (session, ts), drops exact duplicates, splits by type into sorted columnar arrays. */
int ottocov_load_events(ottocov_ctx* ctx, const int32_t* session, const int32_t* aid,
                        const int32_t* ts, const int8_t* type, int64_t n, int where);
int ottocov_get_events_info(ottocov_ctx* ctx, ottocov_events_info* out);

/* ---- (2)+(3) pair expansion + reduce-by-key: replaces self_merge + the filter/groupby of
 * count_co_events (model/count_co_events.py:17-38, :60-77) for ONE co-event kind over the loaded
 * events.  The n^2 join is never materialised.  *out is a new table (caller frees it). */
int ottocov_count(ottocov_ctx* ctx, const ottocov_spec* spec, ottocov_table** out);
int ottocov_get_count_info(ottocov_ctx* ctx, ottocov_count_info* out);

/* ---- streamed ingest + count: replaces the whole per-file loop of count_co_events_all_files
 * (model/count_co_events.py:83-100: read part, unique(), self-merge, count) PLUS the merge of the part tables
 * (concat_files_w_stats, :112-168) for one population, with the host -> device copy of the parts overlapped with the
 * counting.  Input = the parquet parts as the ETL wrote them (etl/jsonl_to_parquet.py:23-29, 59-84: columns
 * session i32, aid i32, ts i32 seconds, type i8; a session never spans two parts, :35-40), one host buffer per
 * column per part -- nothing is concatenated on the host.  One contiguous host array cut at session boundaries is
 * the special case.  Parts are copied on a second stream (all aid columns first: the key width must be known before
 * the first key is made) and processed in groups as they land: loader (any row order inside a part), window ranges,
 * expansion fused with the first bucket pass; the remaining passes, the hash reduce with the fused threshold and
 * the symmetric mirror run once over all groups.  tables_out[k] is the table of specs[k] (same result as
 * ottocov_load_events of the concatenation + ottocov_count).  Afterwards no events are loaded (ottocov_count
 * needs a new ottocov_load_events); ottocov_get_events_info reports the totals over all parts. */
int ottocov_count_parts(ottocov_ctx* ctx, int n_parts, const int32_t* const* session, const int32_t* const* aid,
                        const int32_t* const* ts, const int8_t* const* type, const int64_t* rows,
                        const ottocov_spec* specs, int n_specs, ottocov_table** tables_out);

/* ---- tables: the (aid, aid_next, count) frames the reference writes per part and re-reads in
 * concat_files_w_stats (model/count_co_events.py:97-100, :112-115). */
int ottocov_table_from_arrays(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next,
                              const uint32_t* count, int64_t n, int where, ottocov_table** out);
int ottocov_table_from_packed(ottocov_ctx* ctx, const uint64_t* keys, const uint32_t* count,
                              int64_t n, int where, ottocov_table** out);
int ottocov_table_free(ottocov_ctx* ctx, ottocov_table* t);
int ottocov_table_rows(const ottocov_table* t, int64_t* n_rows);
int ottocov_table_total(ottocov_ctx* ctx, const ottocov_table* t, int64_t* sum_of_counts);

/* groupby(['aid','aid_next']).sum('count') over the concatenation of tables
 * (model/count_co_events.py:168; also :154-155).  Inputs stay valid. */
int ottocov_table_merge(ottocov_ctx* ctx, ottocov_table* const* tabs, int n_tabs, ottocov_table** out);

/* filter(count >= min_count) (model/count_co_events.py:131-132, :156, :172). */
int ottocov_table_filter(ottocov_ctx* ctx, const ottocov_table* t, uint32_t min_count, ottocov_table** out);

/* Copy rows out, optionally in the file order of model/count_co_events.py:173-175
 * (count descending, first `head` rows, count as int32).  Returns rows written in *n_out. */
int ottocov_table_fetch(ottocov_ctx* ctx, const ottocov_table* t, int order, int64_t head,
                        int32_t* aid, int32_t* aid_next, int32_t* count, int64_t cap, int where,
                        int64_t* n_out);
/* Raw packed view (device pointers owned by the table; valid until it is freed). */
int ottocov_table_device_ptrs(const ottocov_table* t, const uint64_t** keys, const uint32_t** count);

/* ---- (4) segmented top-K per aid: replaces sort('aid') + rank('ordinal', reverse=True).over('aid')
 * <= first_n of model/retrieve.py:41-47.  Order inside an aid: count desc, then aid_next asc
 * (canonical rule; the reference's tie order is unspecified).  1 <= k <= 32.
 * Result: n_aids segments; row i has aid_x[i], n_valid[i] = min(k, segment size) and k slots
 * aid_y[i*k..], cnt[i*k..] (unused slots are -1 / 0). */
int ottocov_table_topk(ottocov_ctx* ctx, const ottocov_table* t, int k, int64_t* n_aids);
int ottocov_topk_fetch(ottocov_ctx* ctx, int32_t* aid_x, int32_t* n_valid, int32_t* aid_y,
                       int32_t* cnt, int64_t cap_aids, int where);

/* Candidate lookup on the last top-K result: for each query aid its top-K row (n_valid = 0 when the aid
 * has none).  Replaces the consumer's join of session aids with the top-N rows,
 * df_aids[['aid']].unique().join(df_count[['aid','aid_next']], on='aid')  (model/retrieve.py:75-91). */
int ottocov_topk_lookup(ottocov_ctx* ctx, const int32_t* aids, int64_t n, int where, int32_t* n_valid,
                        int32_t* aid_y /*[n*k]*/, int32_t* cnt /*[n*k]*/);

/* ---- derived co-count features: replaces get_df_count_for_co_event_type
 * (model/retrieve.py:18-63) on a count table given in FILE order (count descending, as count_co_events.py:173-179
 * wrote it): per-aid top-first_n by (count desc, aid_next asc) with, for every kept row,
 *   count_pop = int16(min((count - min) / (q - min), 1) * 10000), q = the 0.9999 quantile of count (:33-35); the caller
 *               passes quantile_row = the row of the ASCENDING count order that its quantile rule selects
 *               (numpy 'nearest': round((n - 1) * 0.9999)), the device sorts the column and reads it
 *   perc_pop  = int16(row_nr / n * 10000), row_nr from 1 in file order (:36-38)
 *   rank      = 1-based ordinal rank inside the aid (:41-44),  count_rel = int8(count / max count of the aid * 100) (:45-49)
 * Rows come back ordered by (aid, rank).  IEEE double arithmetic, bit-identical to the host restatement. */
int ottocov_count_features(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next, const int32_t* count, int64_t n,
                           int where, int first_n, int64_t quantile_row, int64_t* n_rows);
int ottocov_count_features_fetch(ottocov_ctx* ctx, int32_t* aid, int32_t* aid_next, int32_t* count, int16_t* count_pop,
                                 int16_t* perc_pop, int16_t* rank, int8_t* count_rel, int64_t cap_rows, int where);

/* ---- popularity of aids inside session clusters: replaces, for ONE clustering, the body of
 * model/count_popularity.py:56-85 -- groupby([cluster, aid]) with six counts (clicks, carts, orders; all time and
 * ts > ts_recent, the reference's 7-day horizon :54), rank('ordinal', reverse=True).over(cluster) clipped to 999
 * for each count (:73-75), filter(min(ranks) <= keep_top_k) (:81).  `cluster` is the per-EVENT cluster id
 * (-1 = session without a cluster, the fill_null(-1) of :51); events are not de-duplicated (the reference does
 * not either).  Ties: count descending, then aid ascending (the reference's tie order is unspecified).
 * Rows come back ordered by (cluster, aid); ranks is [6][cap_rows] int16 in the reference's column order
 * rank_clicks, rank_carts, rank_orders, rank_clicks_7d, rank_carts_7d, rank_orders_7d. */
int ottocov_count_popularity(ottocov_ctx* ctx, const int32_t* cluster, const int32_t* aid, const int32_t* ts,
                             const int8_t* type, int64_t n, int where, int32_t ts_recent, int keep_top_k,
                             int64_t* n_rows);
int ottocov_popularity_fetch(ottocov_ctx* ctx, int32_t* aid, int32_t* cluster, int16_t* ranks /*[6][cap_rows]*/,
                             int64_t cap_rows, int where);

/* ---- multi-GPU exchange support: stable partition of a table's rows by
 * dest = ottocov_hash_dest(aid, n_ranks) into caller-owned DEVICE buffers (the send buffers of
 * the all-to-all); rows_per_dest is a HOST array [n_ranks].  No reference counterpart (the
 * reference is single-process); counts shard by key because sum is associative. */
int ottocov_table_partition(ottocov_ctx* ctx, const ottocov_table* t, int n_ranks,
                            uint64_t* keys_out_dev, uint32_t* count_out_dev, int64_t* rows_per_dest);
uint32_t ottocov_hash_dest(uint32_t aid, uint32_t n_ranks);

/* Exchange-before-reduce path (used for N > 1): raw co-event keys are expanded, grouped by destination
 * rank (dest = ottocov_hash_dest(aid of the key, n_ranks), stamped into key bits 56..63, one stable
 * radix pass), exchanged by the caller (NCCL all-to-all on its own tensors) and reduced where they land.
 *   ottocov_expand_prepare  window pass for `spec` on the loaded events: how many keys will be emitted,
 *                           and whether they are canonical half pairs of a symmetric kind
 *   ottocov_expand_run      emits them into caller-owned DEVICE buffers (each >= n_keys); the grouped
 *                           keys end in buf_b when *result_in_b, else in buf_a; rows_per_dest is HOST
 *   ottocov_reduce_pairs    sort + run-length count of received keys (keys_dev is used as scratch);
 *                           symmetric = keys are half pairs (diagonal counts are doubled), strip_dest
 *                           bit 0 = clear key bits 56..63 first; OTTOCOV_REDUCE_HASH_ON / _OFF in the
 *                           same argument force the reduce strategy (auto: hash reduce iff min_count > 1)
 *   ottocov_table_mirror    half table (rows a <= b) -> full symmetric table, or only the transposed
 *                           off-diagonal rows (transpose_only), which belong to rank hash(b) */
int ottocov_expand_prepare(ottocov_ctx* ctx, const ottocov_spec* spec, int64_t* n_keys, int* symmetric);
int ottocov_expand_run(ottocov_ctx* ctx, int n_ranks, uint64_t* buf_a_dev, uint64_t* buf_b_dev,
                       int* result_in_b, int64_t* rows_per_dest);
/* Fused partition + exchange: with buf_b_dev == NULL ottocov_expand_run only expands the stamped keys into
 * buf_a_dev and counts them per destination; ottocov_push_keys then runs the ONE distribution pass with
 * each destination's run starting at dest_ptrs[rank] -- a device BYTE address inside that rank's receive
 * buffer, mapped into this process (CUDA IPC / torch symmetric memory), so the keys cross NVLink as the
 * coalesced stores of the partition kernel itself.  dest_ptrs is a HOST array [n_ranks]; the caller
 * exchanged the per-destination counts first and synchronises the ranks around the call. */
int ottocov_push_keys(ottocov_ctx* ctx, const uint64_t* keys_dev, int64_t n, int n_ranks, const uint64_t* dest_ptrs);
int ottocov_reduce_pairs(ottocov_ctx* ctx, uint64_t* keys_dev, int64_t n, int aid_bits, uint32_t min_count,
                         int symmetric, int strip_dest, ottocov_table** out);
int ottocov_table_mirror(ottocov_ctx* ctx, const ottocov_table* t, int transpose_only, ottocov_table** out);

/* ---- fused expansion + exchange (default multi-GPU path, round 2) --------------------------------------------
 * The first distribution pass of the bucketed hash reduce needs no stability, so it runs INSIDE the pair expansion
 * (see ottocov_count); with n_ranks > 1 its digit is (owner rank of the key, low hash-bucket bits) and every digit's
 * region is a stripe of the OWNER's receive area, mapped into this process (CUDA IPC / torch symmetric memory): the
 * keys cross NVLink as the stores of the expansion kernel and land already partitioned for the owner's remaining
 * passes.  No key is written to local HBM first, no NCCL data collective, no per-destination counts on the host.
 *
 *   ottocov_xplan_make      pure host arithmetic, identical on every rank given the same arguments: stripe
 *                           capacities and the byte layout of a rank's receive area
 *                             counts [src][sub] u64 | status [src][4] u64 | hist [src][passes][256] u64 |
 *                             key stripes [src][sub][stripe_cap] u64 | mirrored rows [src]: count, keys, counts
 *                           max_local_keys / total_keys: largest per-rank and summed key count of
 *                           ottocov_expand_prepare over the ranks; mirror_rows_hint: expected thresholded half-table
 *                           rows per rank (0 = derive from total_keys)
 *   ottocov_expand_scatter  after ottocov_expand_prepare: expands this rank's keys into the peers' stripes
 *                           (peer_base[r] = device address of rank r's receive area as seen from this process) and
 *                           publishes per-stripe counts, pass histograms and its status (overflow flag + the
 *                           capacity it would have needed) to EVERY rank.  Only enqueues work.
 *   ottocov_reduce_received after the ranks synchronised: reads the published status; *need_cap > 0 (and no table)
 *                           when some rank overflowed a stripe -- every rank sees the same value, grows the plan
 *                           and repeats -- else the remaining passes + hash reduce over this rank's stripes.
 *                           symmetric: keys are canonical half pairs; the table then holds rows a <= b only, in NO
 *                           particular order: it is the input of ottocov_mirror_push / _collect (which sorts).
 *   ottocov_mirror_push     half table -> its transposed off-diagonal rows (b, a, c) stored into the stripes of
 *                           their owners hash(b) (+ status as above); ottocov_mirror_collect merges what this rank
 *                           received with its own half rows into the full sorted table (*need_rows > 0: grow). */
typedef struct {
    int32_t n_ranks, aid_bits, bucket_bits, sub_bits, rest_passes, reserved;
    int64_t stripe_cap, mirror_cap;
    int64_t off_counts, off_status, off_hist, off_keys, off_mstatus, off_mkeys, off_mcnt, total_bytes;
} ottocov_xplan;
int ottocov_xplan_make(int n_ranks, int aid_bits, int64_t max_local_keys, int64_t total_keys, int64_t stripe_cap,
                       int64_t mirror_cap, ottocov_xplan* out);     /* stripe_cap / mirror_cap: 0 = derive */
int ottocov_expand_scatter(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const uint64_t* peer_base /*HOST [n_ranks]*/);
int ottocov_reduce_received(ottocov_ctx* ctx, const ottocov_xplan* plan, uint64_t recv_area_dev, uint32_t min_count,
                            int symmetric, ottocov_table** out, int64_t* need_cap);
int ottocov_mirror_push(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                        const uint64_t* peer_base /*HOST [n_ranks]*/);
int ottocov_mirror_collect(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                           uint64_t recv_area_dev, ottocov_table** out, int64_t* need_rows);

/* ---- EXTENSION: time-decay weighted co-event scores (north_star config 4) ------------------------------------------
 * The reference has NO weighting in its counts (every in-window pair counts 1, model/count_co_events.py:70-71; SURVEY 0,
 * App. A.6), so this mode has no reference counterpart and its parity is pinned only by this repo's float64 oracle
 * (oracle/cov_oracle.c::cov_oracle_score).  Every pair the integer path counts contributes
 *     w = max(0.10, 1 - |ts_next - ts| / window)        (the decay-with-floor shape of model/kmeans_sessions.py:59)
 * to score(aid, aid_next).  Weights are quantised to 24 fractional bits (|error| <= 3e-8 per pair, <= 3e-7 relative)
 * and summed as integers, so the result is exact in fixed point and independent of the summation order; it agrees
 * with a float64 evaluation to well under the 1e-5 relative error north_star asks for.  Rows keep the integer count
 * too; spec->min_count thresholds that count exactly as in ottocov_count.  At most 2^20 pairs per (aid, aid_next) and
 * 2^32-2 pairs per call (OTTOCOV_ERR_CAPACITY otherwise).  Single GPU. */
typedef struct ottocov_wtable ottocov_wtable;
int ottocov_count_weighted(ottocov_ctx* ctx, const ottocov_spec* spec, ottocov_wtable** out);
int ottocov_wtable_rows(const ottocov_wtable* t, int64_t* n_rows);
int ottocov_wtable_free(ottocov_ctx* ctx, ottocov_wtable* t);
/* rows in (aid, aid_next) order */
int ottocov_wtable_fetch(ottocov_ctx* ctx, const ottocov_wtable* t, int32_t* aid, int32_t* aid_next, double* score,
                         int32_t* count, int64_t cap, int where, int64_t* n_out);
/* per-aid top-k by (score desc, aid_next asc), long format ordered by (aid, rank); rank is 1-based.  Two calls:
 * with cap == 0 only *n_out is set (rows the result has), then with buffers of that size. */
int ottocov_wtable_topk(ottocov_ctx* ctx, const ottocov_wtable* t, int k, int32_t* aid, int32_t* aid_next, double* score,
                        int32_t* rank, int64_t cap, int where, int64_t* n_out);

/* ---- building blocks exposed for tests and micro-benchmarks ---------------------------------- */
/* The bijective key mix of the bucketed hash reduce (host code, needs no GPU): a pair (aid, aid_next), both below
 * 2^aid_bits, <-> a 2*aid_bits-bit mixed key.  ottocov_key_unmix returns aid << 32 | aid_next.  1 <= aid_bits <= 28. */
uint64_t ottocov_key_mix(int aid_bits, uint32_t aid, uint32_t aid_next);
uint64_t ottocov_key_unmix(int aid_bits, uint64_t mixed);
/* LSD radix sort of device-resident 64-bit keys on bits [lo_bit, hi_bit), optional 32-bit
 * payload (vals may be NULL).  Sorted data ends in keys/vals (in place from the caller's view). */
int ottocov_sort_u64(ottocov_ctx* ctx, uint64_t* keys_dev, uint32_t* vals_dev, int64_t n,
                     int lo_bit, int hi_bit);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OTTOCOV_H */
