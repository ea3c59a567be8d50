// popularity.cu -- popularity of aids inside session clusters: counts by event type and 7-day horizon,
// ordinal ranks per cluster, keep the top ranks.  SURVEY 8(f) rank 3: "the same reduce-by-key + segmented
// ranking kernels with a different key".
//
// Replaces the body of the reference's model/count_popularity.py:56-85 for ONE clustering:
//     groupby([cluster, aid]).agg(n_clicks, n_carts, n_orders, n_clicks_7d, n_carts_7d, n_orders_7d)   :61-70
//     rank('ordinal', reverse=True).over(cluster).clip_max(999).cast(Int16) for each of the six counts  :73-75
//     filter(min(ranks) <= keep_top_k)                                                                    :81
// Events are NOT de-duplicated here (the reference does not call unique() in this stage).  The reference's
// ordinal rank breaks ties by the row order its hash group-by happened to leave; the canonical rule here is
// count descending, then aid ascending, which makes the result a pure function of the input.
//
//   pack        key = (cluster + 1) << 32 | aid << 3 | class,  class = type + 3 * (ts > ts_recent)
//   sort + RLE  the engine's radix sort and run-length reduce: one row per (cluster, aid, class)
//   pivot       chained scan over the rows: the <= 6 class rows of a (cluster, aid) group become six counters
//   rank x 6    stable radix sort of the groups by (cluster, ~count) -- groups arrive ordered by (cluster, aid),
//               so ties keep aid order -- then rank = position - first position of the cluster, clipped to 999
//   filter      chained-scan compaction of the groups whose best rank is <= keep_top_k
#include "internal.cuh"
#include "scan.cuh"

constexpr int POP_COLS = 6;
constexpr int POP_AID_SHIFT = 3;          // class lives in key bits [0, 3)
constexpr int POP_GROUP_CL_SHIFT = 29;    // group key = key >> 3: cluster + 1 in bits [29, ...), aid in [0, 29)
constexpr int POP_MAX_AID_BITS = 28;

__global__ void __launch_bounds__(256) pop_pack_kernel(const int32_t* __restrict__ cluster, const int32_t* __restrict__ aid,
                                                       const int32_t* __restrict__ ts, const int8_t* __restrict__ type,
                                                       int64_t n, int32_t ts_recent, u64* __restrict__ keys,
                                                       u32* __restrict__ stat /*[3]: bad flags, OR of aids, max(cluster + 1)*/) {
    u32 bad = 0, aor = 0, cmax = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = cluster[i], a = aid[i], y = type[i];
        if (c < -1) bad |= 1u;
        if (a < 0 || a >= (1 << POP_MAX_AID_BITS)) bad |= 2u;
        if (y < 0 || y > 2) bad |= 4u;
        const u32 cls = (u32)(y & 3) + (ts[i] > ts_recent ? 3u : 0u);
        const u32 c1 = (u32)(c + 1);
        keys[i] = ((u64)c1 << 32) | ((u64)(u32)a << POP_AID_SHIFT) | (u64)cls;
        aor |= (u32)a;
        cmax = c1 > cmax ? c1 : cmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        aor |= __shfl_xor_sync(0xffffffffu, aor, o);
        const u32 m = __shfl_xor_sync(0xffffffffu, cmax, o);
        cmax = m > cmax ? m : cmax;
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicOr(&stat[0], bad);
        if (aor) atomicOr(&stat[1], aor);
        atomicMax(&stat[2], cmax);
    }
}

// rows (cluster, aid, class, count) sorted by key -> one group per (cluster, aid) with its six counters
struct PopPivot {
    static constexpr int NC = 1;
    const u64* k;
    const u32* c;
    int64_t m;
    u64* gkey;       // [cap]
    u32* cnt;        // [POP_COLS][cap]
    int64_t cap;
    __device__ u64 value(int64_t i) const {
        return (i == 0 || (k[i] >> POP_AID_SHIFT) != (k[i - 1] >> POP_AID_SHIFT)) ? 1ull : 0ull;
    }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        const u64 g = k[i] >> POP_AID_SHIFT;
        u32 base3[3] = {0, 0, 0}, recent3[3] = {0, 0, 0};
        for (int64_t j = i; j < m && j < i + 6 && (k[j] >> POP_AID_SHIFT) == g; ++j) {
            const u32 cls = (u32)(k[j] & 7ull);
            const u32 x = c[j];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (cls == (u32)t) base3[t] = x;
                if (cls == (u32)t + 3u) recent3[t] = x;
            }
        }
        const int64_t o = (int64_t)pre[0];
        gkey[o] = g;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            cnt[(int64_t)t * cap + o] = base3[t] + recent3[t];          // n_clicks, n_carts, n_orders
            cnt[(int64_t)(t + 3) * cap + o] = recent3[t];               // n_*_7d
        }
    }
};

__global__ void __launch_bounds__(256) pop_order_keys_kernel(const u64* __restrict__ gkey, const u32* __restrict__ cnt,
                                                             int64_t n, u64* __restrict__ okeys, u32* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    okeys[i] = ((gkey[i] >> POP_GROUP_CL_SHIFT) << 32) | (u64)(0xFFFFFFFFu - cnt[i]);
    idx[i] = (u32)i;
}

// sorted by (cluster, count desc, aid asc): rank = position inside the cluster, 1-based, clipped to 999
__global__ void __launch_bounds__(256) pop_rank_kernel(const u64* __restrict__ okeys, const u32* __restrict__ perm,
                                                       int64_t n, int16_t* __restrict__ rank) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const u64 first = okeys[p] & 0xFFFFFFFF00000000ull;      // smallest possible sort key of this cluster
    int64_t lo = 0, hi = p;                                  // first position with okeys >= first
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (okeys[mid] >= first) hi = mid; else lo = mid + 1;
    }
    const int64_t r = p - lo + 1;
    rank[perm[p]] = (int16_t)(r > 999 ? 999 : r);
}

struct PopKeep {
    static constexpr int NC = 1;
    const u64* gkey;
    const int16_t* rank;     // [POP_COLS][cap]
    int64_t cap;
    int keep_top_k;
    int32_t* out_aid;
    int32_t* out_cl;
    int16_t* out_rank;       // [POP_COLS][cap]
    __device__ u64 value(int64_t i) const {
        int best = 32767;
#pragma unroll
        for (int c = 0; c < POP_COLS; ++c) { const int r = rank[(int64_t)c * cap + i]; best = r < best ? r : best; }
        return best <= keep_top_k ? 1ull : 0ull;
    }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        const int64_t o = (int64_t)pre[0];
        const u64 g = gkey[i];
        out_aid[o] = (int32_t)(g & ((1ull << POP_GROUP_CL_SHIFT) - 1ull));
        out_cl[o] = (int32_t)(g >> POP_GROUP_CL_SHIFT) - 1;
#pragma unroll
        for (int c = 0; c < POP_COLS; ++c) out_rank[(int64_t)c * cap + o] = rank[(int64_t)c * cap + i];
    }
};

void free_popularity(ottocov_ctx* ctx) {
    dev_free(ctx, ctx->pop_aid); dev_free(ctx, ctx->pop_cl); dev_free(ctx, ctx->pop_rank);
    ctx->pop_aid = nullptr; ctx->pop_cl = nullptr; ctx->pop_rank = nullptr;
    ctx->pop_n = 0; ctx->pop_stride = 0; ctx->pop_valid = false;
}

static int bits_of_u32(u32 v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

void count_popularity_impl(ottocov_ctx* ctx, const int32_t* cluster, const int32_t* aid, const int32_t* ts,
                           const int8_t* type, int64_t n, int where, int32_t ts_recent, int keep_top_k) {
    free_popularity(ctx);
    ctx->pop_valid = true;
    if (n == 0) return;
    // ---- columns onto the device ---------------------------------------------------------------------------
    DevBuf<int32_t> d_cl, d_aid, d_ts;
    DevBuf<int8_t> d_type;
    if (where == OTTOCOV_HOST) {
        d_cl.alloc(ctx, n); d_aid.alloc(ctx, n); d_ts.alloc(ctx, n); d_type.alloc(ctx, n);
        CUDA_CHECK(cudaMemcpyAsync(d_cl.p, cluster, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_aid.p, aid, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_ts.p, ts, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_type.p, type, n, cudaMemcpyHostToDevice, ctx->stream));
        cluster = d_cl.p; aid = d_aid.p; ts = d_ts.p; type = d_type.p;
    }
    // ---- pack, validate ----------------------------------------------------------------------------------------
    DevBuf<u64> keys(ctx, n), kalt(ctx, n);
    DevBuf<u32> stat(ctx, 3);
    CUDA_CHECK(cudaMemsetAsync(stat.p, 0, 3 * sizeof(u32), ctx->stream));
    {
        const int grid = (int)imin64(ceil_div64(n, 256), (int64_t)ctx->num_sms * 16);
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 21.0 * n, pop_pack_kernel, grid, 256, 0, cluster, aid, ts, type, n, ts_recent, keys.p,
                   stat.p);
    }
    u32 hs[3];
    cov_readback(ctx, hs, stat.p, sizeof(hs));
    if (hs[0] & 1u) COV_THROW(OTTOCOV_ERR_DATA, "cluster id below -1");
    if (hs[0] & 2u) COV_THROW(OTTOCOV_ERR_DATA, "aid outside [0, 2^%d)", POP_MAX_AID_BITS);
    if (hs[0] & 4u) COV_THROW(OTTOCOV_ERR_DATA, "event type outside {0,1,2}");
    d_cl.release(); d_aid.release(); d_ts.release(); d_type.release();
    const int aid_bits = bits_of_u32(hs[1]) > 0 ? bits_of_u32(hs[1]) : 1;
    const int cl_bits = bits_of_u32(hs[2]) > 0 ? bits_of_u32(hs[2]) : 1;

    // ---- one row per (cluster, aid, class) ----------------------------------------------------------------------
    BitField kf[2] = {{0, POP_AID_SHIFT + aid_bits}, {32, 32 + cl_bits}};
    u64* k = keys.p; u64* ka = kalt.p; u32* v = nullptr; u32* va = nullptr;
    radix_sort_pairs(ctx, k, ka, v, va, n, kf, 2);
    u64* rkeys = nullptr; u32* rcount = nullptr; int64_t m = 0;
    reduce_sorted(ctx, k, nullptr, n, 1, false, &rkeys, &rcount, &m);
    keys.release(); kalt.release();
    struct RowsGuard { ottocov_ctx* c; u64* a; u32* b; ~RowsGuard() { dev_free(c, a); dev_free(c, b); } } rows_guard{ctx, rkeys, rcount};

    // ---- pivot: six counters per (cluster, aid) ------------------------------------------------------------------
    const int64_t cap = m;
    DevBuf<u64> gkey(ctx, cap);
    DevBuf<u32> cnt(ctx, (size_t)POP_COLS * cap);
    PopPivot pv;
    pv.k = rkeys; pv.c = rcount; pv.m = m; pv.gkey = gkey.p; pv.cnt = cnt.p; pv.cap = cap;
    u64 tot[1];
    scan_apply(ctx, OTTOCOV_K_RLE, pv, m, tot, 12.0 * m + 32.0 * m);
    const int64_t G = (int64_t)tot[0];

    // ---- ordinal rank per cluster for each counter ------------------------------------------------------------------
    DevBuf<int16_t> rank(ctx, (size_t)POP_COLS * cap);
    {
        DevBuf<u64> okeys(ctx, G), okalt(ctx, G);
        DevBuf<u32> idx(ctx, G), idxalt(ctx, G);
        BitField of[2] = {{0, 32}, {32, 32 + cl_bits}};
        const unsigned g1 = (unsigned)ceil_div64(G, 256);
        for (int c = 0; c < POP_COLS; ++c) {
            COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 24.0 * G, pop_order_keys_kernel, g1, 256, 0, gkey.p, cnt.p + (size_t)c * cap, G,
                       okeys.p, idx.p);
            u64* sk = okeys.p; u64* ska = okalt.p; u32* sv = idx.p; u32* sva = idxalt.p;
            radix_sort_pairs(ctx, sk, ska, sv, sva, G, of, 2);          // stable: ties keep (cluster, aid) order
            COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 14.0 * G, pop_rank_kernel, g1, 256, 0, sk, sv, G, rank.p + (size_t)c * cap);
        }
    }

    // ---- keep the groups whose best rank is <= keep_top_k -------------------------------------------------------------
    DevBuf<int32_t> o_aid(ctx, G), o_cl(ctx, G);
    DevBuf<int16_t> o_rank(ctx, (size_t)POP_COLS * cap);
    PopKeep kp;
    kp.gkey = gkey.p; kp.rank = rank.p; kp.cap = cap; kp.keep_top_k = keep_top_k;
    kp.out_aid = o_aid.p; kp.out_cl = o_cl.p; kp.out_rank = o_rank.p;
    u64 kept[1];
    scan_apply(ctx, OTTOCOV_K_FILTER, kp, G, kept, 20.0 * G);
    ctx->pop_n = (int64_t)kept[0];
    ctx->pop_stride = cap;
    ctx->pop_aid = o_aid.take();
    ctx->pop_cl = o_cl.take();
    ctx->pop_rank = o_rank.take();
}

void popularity_fetch_impl(ottocov_ctx* ctx, int32_t* aid, int32_t* cluster, int16_t* ranks, int64_t cap_rows, int where) {
    if (!ctx->pop_valid) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_popularity_fetch before ottocov_count_popularity");
    const int64_t n = ctx->pop_n;
    if (cap_rows < n) COV_THROW(OTTOCOV_ERR_CAPACITY, "popularity fetch needs room for %lld rows", (long long)n);
    if (n > 0) {
        if (!aid || !cluster || !ranks) COV_THROW(OTTOCOV_ERR_ARG, "NULL output");
        const cudaMemcpyKind kind = (where == OTTOCOV_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        CUDA_CHECK(cudaMemcpyAsync(aid, ctx->pop_aid, n * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(cluster, ctx->pop_cl, n * 4, kind, ctx->stream));
        for (int c = 0; c < POP_COLS; ++c)       // caller layout: [6][cap_rows]
            CUDA_CHECK(cudaMemcpyAsync(ranks + (size_t)c * cap_rows, ctx->pop_rank + (size_t)c * ctx->pop_stride, n * 2, kind,
                                       ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
