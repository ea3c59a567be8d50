// internal.cuh -- shared declarations of libottocov.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <stdexcept>
#include "../../include/ottocov.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- errors: C++ exceptions inside, status codes at the ABI ------------------------------------
struct CovError {
    int code;
    std::string msg;
};

#define COV_THROW(code_, ...)                                           \
    do {                                                                \
        char _b[512];                                                   \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                          \
        throw CovError{(code_), std::string(_b)};                       \
    } while (0)

#define CUDA_CHECK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            int _c = (_e == cudaErrorMemoryAllocation) ? OTTOCOV_ERR_NOMEM : OTTOCOV_ERR_CUDA; \
            COV_THROW(_c, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                      cudaGetErrorString(_e));                                             \
        }                                                                                  \
    } while (0)

// OTTOCOV_TRACE=1: print host wall time between marks (each mark synchronises the stream) -- a poor
// man's timeline for finding host-side gaps; off by default, never used in timed runs.
void cov_trace(struct ottocov_ctx* ctx, const char* what);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

// ---- per-type sorted event arrays (the CSR-by-session layout, SoA) -------------------------------
// Events of one type, ordered by skey = (session - session_min) << 32 | (ts - ts_min).
struct TypeArray {
    u64* skey = nullptr;     // [n] search key
    u32* aid = nullptr;      // [n]
    u32* xrank[2] = {nullptr, nullptr};  // [n] insertion rank of this event in type (t+1)%3 / (t+2)%3
    int64_t n = 0;
};

struct ProfEvent { int family; cudaEvent_t a, b; };
struct CacheBlock { void* p; size_t bytes; };

struct ottocov_table {
    u64* keys = nullptr;     // [n] sorted, distinct (buffers may be larger than n)
    u32* count = nullptr;    // [n]
    int64_t n = 0;
    int aid_bits = 32;       // significant bits of both key halves (bounds the radix passes)
};

struct ottocov_wtable {      // EXTENSION: time-decay weighted scores (expand.cu, "weighted")
    u64* keys = nullptr;     // [n] sorted, distinct
    u32* count = nullptr;    // [n] integer pair counts
    u64* score_fx = nullptr; // [n] sum of weights, 24 fractional bits
    int64_t n = 0;
    int aid_bits = 32;
};

struct ottocov_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = 0;
    std::string err;
    // profiling / accounting
    unsigned profiling = 0;            // bit f set: CUDA-event timing of kernel family f
    ottocov_kernel_stat stats[OTTOCOV_K_FAMILIES];
    std::vector<ProfEvent> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    // caching allocator: freed blocks are kept and re-used (all work is ordered on one stream), so
    // the steady-state pipeline performs no driver allocations at all
    std::vector<CacheBlock> cache;
    std::unordered_map<void*, size_t> live;
    size_t cached_bytes = 0, live_bytes = 0, peak_bytes = 0;
    // events
    bool loaded = false;
    bool info_only = false;            // `info` holds the totals of an ottocov_count_parts run (no events are loaded)
    ottocov_events_info info;
    TypeArray ta[3];
    ottocov_count_info last_count;
    // onesweep look-back state (grow-only, epoch-tagged so it is never re-zeroed per pass)
    u64* sweep_status = nullptr;
    size_t sweep_status_words = 0;
    u32 sweep_epoch = 0;
    u32* sweep_ticket = nullptr;       // [64] tickets, one per pass slot
    // generic look-back scan state
    u64* scan_status = nullptr;
    size_t scan_status_words = 0;
    u32 scan_epoch = 0;
    u32* scan_ticket = nullptr;
    u64* scan_totals = nullptr;        // [8] grand totals of the last scan launch
    void* host_stage = nullptr;        // page-locked staging for data the library prepares on the host (session runs)
    size_t host_stage_bytes = 0;
    int64_t h2d_bytes_last = 0;        // bytes the last ottocov_count_parts copied host -> device
    void* pinned = nullptr;            // 4 KB page-locked landing pad for small device -> host read-backs
    cudaStream_t copy_stream = nullptr;        // host -> device column copies overlapped with the loader (events.cu)
    std::vector<cudaEvent_t> sync_events;      // pool of timing-less events for cross-stream ordering
    std::unordered_map<const void*, size_t> func_smem;   // kernels opted in to > 48 KB dynamic shared memory on THIS device
    u64 budget_cache = 0;              // pair budget derived from free HBM (expand.cu::auto_budget); 0 = not computed
    void* plan = nullptr;              // ExpandPlan between ottocov_expand_prepare and ottocov_expand_run
    // top-k result
    int topk_k = 0;
    int64_t topk_n = 0;
    int32_t* topk_aid_x = nullptr;
    int32_t* topk_nvalid = nullptr;
    int32_t* topk_aid_y = nullptr;
    int32_t* topk_cnt = nullptr;
    // derived count features of the last ottocov_count_features (features.cu): kept rows, struct of arrays
    bool feat_valid = false;
    int64_t feat_n = 0, feat_cap = 0;
    int32_t* feat_i32 = nullptr;       // [3][feat_cap] aid | aid_next | count
    int16_t* feat_i16 = nullptr;       // [3][feat_cap] count_pop | perc_pop | rank
    int8_t* feat_i8 = nullptr;         // [feat_cap] count_rel
    // popularity result (popularity.cu): rows kept by the last ottocov_count_popularity
    bool pop_valid = false;
    int64_t pop_n = 0, pop_stride = 0;
    int32_t* pop_aid = nullptr;
    int32_t* pop_cl = nullptr;
    int16_t* pop_rank = nullptr;       // [6][pop_stride]

    void begin(int family);
    void end(int family, double algo_bytes);
};

// Small device -> host read-back (totals, statistics) through the context's page-locked pad, then a
// stream synchronise: a pageable destination would make the runtime stage and block on its own terms.
void cov_readback(ottocov_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
void* cov_alloc(ottocov_ctx* ctx, size_t bytes);      // api.cu; throws CovError
void cov_free(ottocov_ctx* ctx, void* p);             // returns the block to the context cache
void cov_trim(ottocov_ctx* ctx);                      // releases every cached block to the driver

// RAII device buffer on the context's caching allocator.
template <class T>
struct DevBuf {
    ottocov_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(ottocov_ctx* c, size_t n_) { alloc(c, n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void alloc(ottocov_ctx* c, size_t n_) {
        release();
        ctx = c; n = n_;
        size_t bytes = (n_ ? n_ : 1) * sizeof(T);
        p = (T*)cov_alloc(c, bytes);
    }
    T* take() { T* q = p; p = nullptr; n = 0; return q; }   // give ownership away
    void release() {
        if (p) { cov_free(ctx, p); p = nullptr; }
        n = 0;
    }
    ~DevBuf() { release(); }
};

static inline void dev_free(ottocov_ctx* ctx, void* p) { if (p) cov_free(ctx, p); }

// Launch helper: counts the launch, brackets it with CUDA events when profiling is on.
#define COV_LAUNCH(ctx_, fam_, bytes_, kern_, grid_, block_, smem_, ...)                      \
    do {                                                                                      \
        (ctx_)->begin(fam_);                                                                  \
        kern_<<<(grid_), (block_), (smem_), (ctx_)->stream>>>(__VA_ARGS__);                  \
        (ctx_)->end(fam_, (double)(bytes_));                                                  \
        CUDA_CHECK(cudaGetLastError());                                                       \
    } while (0)

// ---- bijective key mix (bucketed hash reduce, hash_reduce.cu) --------------------------------------
// A pair key (aid << 32 | aid_next, both < 2^ab) is compacted to kb = 2 ab bits and scrambled by an
// invertible multiply / xor-shift / multiply / xor-shift on kb bits.  Equal keys stay equal, the top
// bits of the result are uniform whatever the popularity skew of the aids, and the plain key is
// recovered exactly by running the steps backwards with the modular inverses of the multipliers.
struct KeyMix {
    int ab = 0, kb = 0, s = 0;   // aid bits, key bits, xor-shift distance (2 s >= kb: self-inverse)
    u64 mask = 0;
    u64 m1 = 0, m2 = 0, m1inv = 0, m2inv = 0;
};

static inline u64 inv_odd_u64(u64 m) {          // Newton iteration: exact inverse mod 2^64 of an odd m
    u64 x = m;                                   // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2ull - m * x;
    return x;
}

static inline KeyMix make_key_mix(int aid_bits) {
    KeyMix m;
    m.ab = aid_bits;
    m.kb = 2 * aid_bits;
    m.s = (m.kb + 1) / 2;
    m.mask = m.kb >= 64 ? ~0ull : ((1ull << m.kb) - 1ull);
    m.m1 = 0x9E3779B97F4A7C15ull;
    m.m2 = 0xBF58476D1CE4E5B9ull;
    m.m1inv = inv_odd_u64(m.m1);
    m.m2inv = inv_odd_u64(m.m2);
    return m;
}

#ifdef __CUDACC__
#define COV_HD __host__ __device__ __forceinline__
#else
#define COV_HD static inline
#endif
COV_HD u64 key_mix_fwd(const KeyMix& m, u32 x, u32 y) {
    u64 h = (((u64)x << m.ab) | (u64)y);
    h = (h * m.m1) & m.mask;
    h ^= h >> m.s;
    h = (h * m.m2) & m.mask;
    h ^= h >> m.s;
    return h;
}
COV_HD u64 key_mix_inv(const KeyMix& m, u64 h) {     // -> plain key aid << 32 | aid_next
    h ^= h >> m.s;
    h = (h * m.m2inv) & m.mask;
    h ^= h >> m.s;
    h = (h * m.m1inv) & m.mask;
    const u64 x = h >> m.ab, y = h & ((1ull << m.ab) - 1ull);
    return (x << 32) | y;
}

constexpr u32 HR_FLAG_FUSED_OVERFLOW = 16u;      // device flag: a region of the fused first pass was too small

// ---- device building blocks (implemented in the .cu files) -------------------------------------
// radix_sort.cu
struct BitField { int lo, hi; };   // sort on key bits [lo, hi)
constexpr int RS_MAX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_MAX_BITS;
constexpr int RS_MAX_PASSES = 16;
struct PassList {                  // the distribution passes a field list expands to (<= 8 bits each)
    int n;
    int shift[RS_MAX_PASSES];
    int bits[RS_MAX_PASSES];
};
PassList make_pass_list(const BitField* fields, int n_fields);
// Sorts n keys (+ optional payload) on the given fields, least significant field first.
// keys/alt (and vals/valt) are a double buffer; on return `keys`/`vals` point at the sorted data.
// Returns the number of radix passes it ran.  pre_hist: optional device array [passes][RS_RADIX] of raw
// digit counts of these keys (accumulated by whoever wrote them); saves the histogram read.
int radix_sort_pairs(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n,
                     const BitField* fields, int n_fields, u64* pre_hist = nullptr);
// seg_cnt (optional, keys only, needs pre_hist): the n input keys sit in n_a * n_b regions inside `keys`: region
// a * n_b + b starts at key offset seg_off[.] and holds seg_cnt[.] keys (device arrays); logical order: b-major.
int radix_sort_passes(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n, const PassList& pl,
                      u64* pre_hist = nullptr, const u64* seg_cnt = nullptr, const u64* seg_off = nullptr, int n_a = 1,
                      int n_b = 1, const u32* abort_flag = nullptr,      // abort_flag: see rs_onesweep_kernel
                      u64* bucket_bounds = nullptr);                     // [2^bits * n_b] pre-set to ~0: see RsSegs (one pass only)
// cudaFuncAttributeMaxDynamicSharedMemorySize opt-in, once per (context = device, kernel)
void cov_func_smem(ottocov_ctx* ctx, const void* func, size_t bytes);

void radix_partition_push(ottocov_ctx* ctx, const u64* keys, int64_t n, int shift, int bits,
                          const u64* ptr_base_host, int n_digits);

// events.cu
void load_events_impl(ottocov_ctx* ctx, const int32_t* session, const int32_t* aid,
                      const int32_t* ts, const int8_t* type, int64_t n, int where);
void free_events(ottocov_ctx* ctx);

// expand.cu
ottocov_table* count_impl(ottocov_ctx* ctx, const ottocov_spec* spec);
void free_plan(ottocov_ctx* ctx);
void count_parts_impl(ottocov_ctx* ctx, int n_parts, const int32_t* const* session, const int32_t* const* aid,
                      const int32_t* const* ts, const int8_t* const* type, const int64_t* rows, const ottocov_spec* specs,
                      int n_specs, ottocov_table** tables_out);
void expand_prepare_impl(ottocov_ctx* ctx, const ottocov_spec* spec, int64_t* n_keys, int* symmetric);
void expand_run_impl(ottocov_ctx* ctx, int n_ranks, u64* buf_a, u64* buf_b, int* result_in_b, int64_t* rows_per_dest);
void push_keys_impl(ottocov_ctx* ctx, const u64* keys, int64_t n, int n_ranks, const u64* dest_ptrs_host);
ottocov_table* reduce_pairs_impl(ottocov_ctx* ctx, u64* keys, int64_t n, int aid_bits, u32 min_count, int sym,
                                 int strip_dest);

// weighted extension (expand.cu)
ottocov_wtable* count_weighted_impl(ottocov_ctx* ctx, const ottocov_spec* spec);
void wtable_fetch_impl(ottocov_ctx* ctx, const ottocov_wtable* t, int32_t* aid, int32_t* aid_next, double* score,
                       int32_t* count, int64_t cap, int where, int64_t* n_out);
void wtable_topk_impl(ottocov_ctx* ctx, const ottocov_wtable* t, int k, int32_t* aid, int32_t* aid_next, double* score,
                      int32_t* rank, int64_t cap, int where, int64_t* n_out);

// fused expansion + exchange (expand.cu, exchange.cu)
void xplan_make_impl(int n_ranks, int aid_bits, int64_t max_local_keys, int64_t total_keys, int64_t stripe_cap,
                     int64_t mirror_cap, ottocov_xplan* out);
void expand_scatter_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const u64* peer_base_host);
ottocov_table* reduce_received_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, u64 recv_area, u32 min_count, int sym,
                                    int64_t* need_cap);
void mirror_push_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half, const u64* peer_base_host);
ottocov_table* mirror_collect_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                                   u64 recv_area, int64_t* need_rows);
constexpr int XCH_MAX_RANKS = 64;
struct PeerBases { u64 p[XCH_MAX_RANKS]; };       // device addresses of every rank's receive area, passed by value

// reduce.cu
// sorted keys -> distinct keys + run lengths (vals == nullptr) or summed payload (vals != nullptr),
// keeping only rows whose count is >= min_count
// `sym`: keys are canonical (min aid, max aid) pairs counted once per unordered event pair; a
// diagonal key (a, a) then stands for two ordered pairs, so its total is doubled before the threshold.
void reduce_sorted(ottocov_ctx* ctx, const u64* keys, const u32* vals, int64_t n, u32 min_count, bool sym,
                   u64** out_keys, u32** out_count, int64_t* n_out);
// half table (rows a <= b) -> full symmetric table
// transpose_only: return just the mirrored rows (b, a, c) of the off-diagonal rows, sorted
ottocov_table* mirror_table_impl(ottocov_ctx* ctx, const ottocov_table* half, bool transpose_only);
ottocov_table* merge_tables_impl(ottocov_ctx* ctx, ottocov_table* const* tabs, int n_tabs);
ottocov_table* filter_table_impl(ottocov_ctx* ctx, const ottocov_table* t, u32 min_count);
void fetch_table_impl(ottocov_ctx* ctx, const ottocov_table* t, int order, int64_t head,
                      int32_t* aid, int32_t* aid_next, int32_t* count, int64_t cap, int where,
                      int64_t* n_out);
int64_t table_total_impl(ottocov_ctx* ctx, const ottocov_table* t);
ottocov_table* table_from_arrays_impl(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next,
                                      const u32* count, int64_t n, int where);
ottocov_table* table_from_packed_impl(ottocov_ctx* ctx, const u64* keys, const u32* count,
                                      int64_t n, int where);
void partition_table_impl(ottocov_ctx* ctx, const ottocov_table* t, int n_ranks, u64* keys_out,
                          u32* count_out, int64_t* rows_per_dest);

// hash_reduce.cu
// Bucketed hash aggregation of n MIXED keys (key_mix_fwd; bits above mix.kb clear).  keys/alt are a
// double buffer of n keys each and are used as scratch.  Returns the table of plain keys whose count is
// >= min_count, sorted by key.  sym: keys are canonical half pairs (diagonal totals doubled); mirror
// (sym only): also emit the transposed off-diagonal rows, i.e. return the full symmetric table.
// The keys are first sorted on their top hashed_bucket_bits(n, mix.kb) bits, i.e. on the field
// [mix.kb - bb, mix.kb); pre_hist (optional) = raw digit counts of exactly those passes (see radix_sort_pairs).
bool hashed_reduce_supported(int aid_bits);
int hashed_bucket_bits(int64_t n, int kb);
// Fused first pass: whoever wrote the keys (expand_scatter_kernel) already distributed them on the lowest digit of
// the bucket field [kb - bb, kb) (make_pass_list's pass 0) into n_a * n_b slack regions of seg_stride keys inside
// `keys`; seg_cnt = their fill counters (device), ctr = device [2] (rows written, flags), zeroed before the keys were
// written: a writer that ran out of room in a region sets HR_FLAG_FUSED_OVERFLOW there and hashed_reduce throws
// FusedOverflow (the caller re-expands the unfused way).  pre_hist then covers the REMAINING passes only.
struct FusedOverflow {};
struct HashPre {
    int bb;
    int first_bits;          // width of the digit the writer partitioned on (0 = make_pass_list's pass 0 of the field)
    const u64* seg_cnt;      // [n_a * n_b] fill counters of the regions (device)
    const u64* seg_off;      // [n_a * n_b] key offset of each region inside `keys` (device)
    int n_a, n_b;
    unsigned long long* ctr;
    bool skip_sort = false;  // leave the surviving rows in bucket order (the caller sorts them later anyway)
    bool big = false;        // whole-bucket reduce (hash_reduce_buckets_kernel): bb = first_bits + ONE more pass
};
// Bucket bits of the whole-bucket reduce for n keys, or 0 when it does not save a distribution pass over
// hashed_bucket_bits (or cannot serve keys this wide).
int hashed_big_bucket_bits(int64_t n, int kb);
ottocov_table* hashed_reduce(ottocov_ctx* ctx, u64* keys, u64* alt, int64_t n, const KeyMix& mix, u32 min_count,
                             bool sym, bool mirror, int* passes_out, u64* pre_hist = nullptr, const HashPre* pre = nullptr);
// Reduce over several bucket-sorted arrays (streamed ingest): see hash_reduce.cu, "the same reduce over SEVERAL ...".
u32 hashed_range_buckets(int64_t n_est, int bb);                 // buckets per CTA so that a CTA counts ~2048 keys
int64_t hashed_n_ranges(int bb, u32 rb);
void hashed_range_bounds(ottocov_ctx* ctx, const u64* keys, int64_t n, int rem_bits, u32 rb, int64_t n_ranges, u32* bounds,
                         const u32* abort_flag);
ottocov_table* hashed_reduce_groups(ottocov_ctx* ctx, const u64* const* keys, const u32* const* bounds, int n_groups,
                                    int64_t n_total, int bb, u32 rb, const KeyMix& mix, u32 min_count, bool sym, bool mirror,
                                    unsigned long long* ctr);
// plain keys (optionally with a destination stamp in bits 56..63) -> mixed keys, in place; also fills ghist
// (device, [pl.n][RS_RADIX]) with the digit counts of the passes in pl (hashed_reduce's pre_hist)
void mix_keys_inplace(ottocov_ctx* ctx, u64* keys, int64_t n, const KeyMix& mix, bool strip_dest, const PassList& pl,
                      u64* ghist);

// topk.cu
void topk_impl(ottocov_ctx* ctx, const ottocov_table* t, int k);
void free_topk(ottocov_ctx* ctx);
void topk_lookup_impl(ottocov_ctx* ctx, const int32_t* aids, int64_t n, int where, int32_t* n_valid, int32_t* aid_y,
                      int32_t* cnt);

// features.cu
void count_features_impl(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next, const int32_t* count, int64_t n,
                         int where, int first_n, int64_t quantile_row);
void count_features_fetch_impl(ottocov_ctx* ctx, int32_t* aid, int32_t* aid_next, int32_t* count, int16_t* count_pop,
                               int16_t* perc_pop, int16_t* rank, int8_t* count_rel, int64_t cap, int where);
void free_features(ottocov_ctx* ctx);

// popularity.cu
void count_popularity_impl(ottocov_ctx* ctx, const int32_t* cluster, const int32_t* aid, const int32_t* ts,
                           const int8_t* type, int64_t n, int where, int32_t ts_recent, int keep_top_k);
void popularity_fetch_impl(ottocov_ctx* ctx, int32_t* aid, int32_t* cluster, int16_t* ranks, int64_t cap_rows, int where);
void free_popularity(ottocov_ctx* ctx);

// ---- small device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// streaming (evict-first) accessors for data touched once per pass
__device__ __forceinline__ u64 ld_stream_u64(const u64* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream_u64(u64* p, u64 v) { __stcs(p, v); }

__device__ __forceinline__ u32 hash_dest(u32 aid, u32 n_ranks) {
    u32 h = aid * 0x9E3779B1u;
    h ^= h >> 15;
    return h % n_ranks;
}

// Block-wide exclusive scan of one value per thread (blockDim.x == THREADS, multiple of 32).
// `s_warp` must hold THREADS/32 + 1 elements of T.  Returns the exclusive prefix; *total gets the
// block sum (same value in every thread).  Contains three __syncthreads().
template <class T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                                // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = THREADS / 32;
        T w = (lane < NW) ? s_warp[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NW) s_warp[lane] = winc - w;     // exclusive warp offsets
        if (lane == NW - 1) s_warp[NW] = winc;      // block total
    }
    __syncthreads();
    T res = s_warp[warp] + inc - v;
    *total = s_warp[THREADS / 32];
    return res;
}
#endif
