// features.cu -- get_df_count_for_co_event_type (model/retrieve.py:18-63) on the device.
//
// The consumer reads {count_type}.parquet (rows in FILE order = count descending, count_co_events.py:173) and derives
//   count_pop  = int16( min((count - min) / (quantile_0.9999 - min), 1) * 10_000 )      retrieve.py:33-35  (population)
//   perc_pop   = int16( row_nr / n * 10_000 ), row_nr from 1 in file order              retrieve.py:36-38
//   rank       = ordinal rank of count, descending, over aid; keep rank <= first_n      retrieve.py:41-47
//   count_rel  = int8( count / max(count over aid) * 100 )                              retrieve.py:45-49
// The reference sorts the whole table by aid and runs window functions over it.  Here: the rows are sorted by
// (aid, aid_next) with their file row number as payload, the per-aid top-N comes from topk.cu (canonical tie rule:
// count desc, aid_next asc), the two population statistics from one sort of the count column, and one compaction
// pass writes the kept rows with their four features.  The float arithmetic is IEEE double, operation for
// operation what numpy does on the host, so the truncated integers are bit-identical to the restatement's.
#include "internal.cuh"
#include "scan.cuh"

__global__ void __launch_bounds__(256) feat_pack_kernel(const int32_t* __restrict__ aid, const int32_t* __restrict__ aid_next,
                                                        const int32_t* __restrict__ count, int64_t n, u64* __restrict__ keys,
                                                        u32* __restrict__ idx, u64* __restrict__ cnt64, u32* __restrict__ stat) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t a = aid[i], b = aid_next[i], c = count[i];
    if (a < 0 || b < 0 || c < 0) atomicOr(&stat[0], 1u);
    atomicOr(&stat[1], (u32)a | (u32)b);            // OR of the ids: bounds the sort passes
    atomicMax(&stat[2], (u32)c);
    keys[i] = ((u64)(u32)a << 32) | (u64)(u32)b;
    idx[i] = (u32)i;
    cnt64[i] = (u64)(u32)c;
}

__global__ void __launch_bounds__(256) feat_gather_kernel(const u32* __restrict__ idx, const int32_t* __restrict__ count, int64_t n,
                                                          u32* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)count[idx[i]];
}

// one element per aid segment of the top-N result; writes its n_valid kept rows
struct FeatureRows {
    static constexpr int NC = 1;
    const int32_t* aid_x; const int32_t* nvalid; const int32_t* aid_y; const int32_t* cnt;
    int k;
    const u64* keys; const u32* idx; int64_t n;                // key-sorted table + file row numbers
    double cmin, denom;
    int32_t* o_aid; int32_t* o_next; int32_t* o_cnt; int16_t* o_pop; int16_t* o_perc; int16_t* o_rank; int8_t* o_rel;
    __device__ u64 value(int64_t s) const { return (u64)nvalid[s]; }
    __device__ void apply(int64_t s, u64 v, const u64* pre) const {
        const u32 x = (u32)aid_x[s];
        const double mx = (double)cnt[s * k];
        for (int l = 0; l < (int)v; ++l) {
            const u32 y = (u32)aid_y[s * k + l];
            const int32_t c = cnt[s * k + l];
            const u64 key = ((u64)x << 32) | y;
            int64_t lo = 0, hi = n;                        // the row (x, y) of the key-sorted table
            while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < key) lo = mid + 1; else hi = mid; }
            const double row_nr = (double)idx[lo] + 1.0;
            const int64_t o = (int64_t)pre[0] + l;
            o_aid[o] = (int32_t)x; o_next[o] = (int32_t)y; o_cnt[o] = c;
            o_pop[o] = (int16_t)(int)(fmin(((double)c - cmin) / denom, 1.0) * 10000.0);
            o_perc[o] = (int16_t)(int)(row_nr / (double)n * 10000.0);
            o_rank[o] = (int16_t)(l + 1);
            o_rel[o] = (int8_t)(int)((double)c / mx * 100.0);
        }
    }
};

void free_features(ottocov_ctx* ctx) {
    dev_free(ctx, ctx->feat_i32); dev_free(ctx, ctx->feat_i16); dev_free(ctx, ctx->feat_i8);
    ctx->feat_i32 = nullptr; ctx->feat_i16 = nullptr; ctx->feat_i8 = nullptr;
    ctx->feat_n = 0; ctx->feat_cap = 0; ctx->feat_valid = false;
}

void count_features_impl(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next, const int32_t* count, int64_t n,
                         int where, int first_n, int64_t quantile_row) {
    free_features(ctx);
    ctx->feat_valid = true;
    if (n == 0) return;
    if (first_n < 1 || first_n > 32) COV_THROW(OTTOCOV_ERR_ARG, "first_n must be 1..32");
    if (n >= (int64_t)0xFFFFFFFFll) COV_THROW(OTTOCOV_ERR_ARG, "at most 2^32-2 rows");
    if (quantile_row < 0 || quantile_row >= n) COV_THROW(OTTOCOV_ERR_ARG, "quantile row out of range");
    DevBuf<int32_t> da, db, dc;
    if (where == OTTOCOV_HOST) {
        da.alloc(ctx, n); db.alloc(ctx, n); dc.alloc(ctx, n);
        CUDA_CHECK(cudaMemcpyAsync(da.p, aid, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(db.p, aid_next, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(dc.p, count, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        aid = da.p; aid_next = db.p; count = dc.p;
    }
    DevBuf<u64> keys(ctx, n), kalt(ctx, n), c64(ctx, n);
    DevBuf<u32> idx(ctx, n), ialt(ctx, n), stat(ctx, 4);
    CUDA_CHECK(cudaMemsetAsync(stat.p, 0, 16, ctx->stream));
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 28.0 * n, feat_pack_kernel, (unsigned)ceil_div64(n, 256), 256, 0, aid, aid_next, count, n,
               keys.p, idx.p, c64.p, stat.p);
    u32 hs[4];
    cov_readback(ctx, hs, stat.p, 16);
    if (hs[0]) COV_THROW(OTTOCOV_ERR_DATA, "negative aid or count in the count table");
    int aid_bits = 0, cnt_bits = 0;
    for (u32 v = hs[1]; v; v >>= 1) ++aid_bits;
    for (u32 v = hs[2]; v; v >>= 1) ++cnt_bits;
    if (aid_bits == 0) aid_bits = 1;
    // (aid, aid_next) order, file row number as payload
    u64* k = keys.p; u64* ka = kalt.p; u32* v = idx.p; u32* va = ialt.p;
    BitField fields[2] = {{0, aid_bits}, {32, 32 + aid_bits}};
    radix_sort_pairs(ctx, k, ka, v, va, n, fields, 2);
    DevBuf<u32> cs(ctx, n);
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 12.0 * n, feat_gather_kernel, (unsigned)ceil_div64(n, 256), 256, 0, v, count, n, cs.p);
    // population statistics: min and the 0.9999 quantile ("nearest": the caller passes the row of the ascending order)
    {
        u64* q = c64.p; u64* qa = (k == keys.p) ? kalt.p : keys.p;      // the spare half of the key double buffer
        u32* nv = nullptr; u32* nva = nullptr;
        BitField cf[1] = {{0, cnt_bits}};
        if (cnt_bits > 0) radix_sort_pairs(ctx, q, qa, nv, nva, n, cf, 1);
        u64 two[2];
        CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, q, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync((char*)ctx->pinned + 8, q + quantile_row, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        memcpy(two, ctx->pinned, 16);
        ottocov_table tab;
        tab.keys = k; tab.count = cs.p; tab.n = n; tab.aid_bits = aid_bits;
        topk_impl(ctx, &tab, first_n);
        const int64_t A = ctx->topk_n;
        const int64_t cap = A * first_n < n ? A * first_n : n;
        ctx->feat_i32 = (int32_t*)cov_alloc(ctx, (size_t)(cap > 0 ? cap : 1) * 3 * 4);
        ctx->feat_i16 = (int16_t*)cov_alloc(ctx, (size_t)(cap > 0 ? cap : 1) * 3 * 2);
        ctx->feat_i8 = (int8_t*)cov_alloc(ctx, (size_t)(cap > 0 ? cap : 1));
        ctx->feat_cap = cap;
        FeatureRows f;
        f.aid_x = ctx->topk_aid_x; f.nvalid = ctx->topk_nvalid; f.aid_y = ctx->topk_aid_y; f.cnt = ctx->topk_cnt; f.k = first_n;
        f.keys = k; f.idx = v; f.n = n;
        f.cmin = (double)two[0];
        const double d = (double)two[1] - (double)two[0];
        f.denom = d > 1e-12 ? d : 1e-12;                       // retrieve.py:34 guards the division the same way
        f.o_aid = ctx->feat_i32; f.o_next = ctx->feat_i32 + cap; f.o_cnt = ctx->feat_i32 + 2 * cap;
        f.o_pop = ctx->feat_i16; f.o_perc = ctx->feat_i16 + cap; f.o_rank = ctx->feat_i16 + 2 * cap;
        f.o_rel = ctx->feat_i8;
        u64 tot[1];
        scan_apply(ctx, OTTOCOV_K_TOPK, f, A, tot, 8.0 * first_n * A + 19.0 * (double)cap);
        ctx->feat_n = (int64_t)tot[0];
    }
}

void count_features_fetch_impl(ottocov_ctx* ctx, int32_t* aid, int32_t* aid_next, int32_t* count, int16_t* count_pop,
                               int16_t* perc_pop, int16_t* rank, int8_t* count_rel, int64_t cap, int where) {
    if (!ctx->feat_valid) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_count_features_fetch before ottocov_count_features");
    const int64_t m = ctx->feat_n, c = ctx->feat_cap;
    if (cap < m) COV_THROW(OTTOCOV_ERR_CAPACITY, "feature fetch needs room for %lld rows", (long long)m);
    if (m > 0) {
        const cudaMemcpyKind kind = (where == OTTOCOV_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        CUDA_CHECK(cudaMemcpyAsync(aid, ctx->feat_i32, m * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_next, ctx->feat_i32 + c, m * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(count, ctx->feat_i32 + 2 * c, m * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(count_pop, ctx->feat_i16, m * 2, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(perc_pop, ctx->feat_i16 + c, m * 2, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(rank, ctx->feat_i16 + 2 * c, m * 2, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(count_rel, ctx->feat_i8, m, kind, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
