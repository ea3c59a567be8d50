// topk.cu -- segmented top-K per aid over a key-sorted (aid, aid_next, count) table.
//
// Replaces sort(['aid']) + rank('ordinal', reverse=True).over('aid') <= first_n of the consumer
// (model/retrieve.py:41-47; first_n = 10/10/20/20/20, config.py:90-96).  The reference's order among
// equal counts is whatever its upstream hash group-by left; the canonical rule here
// is count descending, then aid_next ascending, which makes the result a pure function of the table.
//
// Rows of one aid are contiguous (the table is sorted by key).  One warp owns one aid segment and
// streams it 32 rows at a time, keeping the running top-32 sorted across its lanes: a batch is
// looked at only if some row beats the current K-th best (one ballot); then the batch is bitonic-
// sorted across lanes and bitonic-merged into the running list.  The composite compare key
// (count << 32 | ~aid_next) is unique inside a segment, so there are no ties to break.
#include "internal.cuh"
#include "scan.cuh"

struct SegmentHeads {
    static constexpr int NC = 1;
    const u64* keys;
    u64* seg_start;
    __device__ u64 value(int64_t i) const {
        return (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1ull : 0ull;
    }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (v && seg_start) seg_start[pre[0]] = (u64)i;
    }
};

__device__ __forceinline__ u64 umax64(u64 a, u64 b) { return a > b ? a : b; }
__device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }

// full bitonic sort of one value per lane, descending by lane
__device__ __forceinline__ u64 warp_sort_desc(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 p = __shfl_xor_sync(0xffffffffu, v, j);
            const bool desc_block = ((lane & k) == 0);     // k == 32: always true -> final order descending
            const bool lower = ((lane & j) == 0);
            v = (lower == desc_block) ? umax64(v, p) : umin64(v, p);
        }
    }
    return v;
}

// lanes hold a bitonic sequence -> descending by lane
__device__ __forceinline__ u64 warp_bitonic_merge_desc(u64 v, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const u64 p = __shfl_xor_sync(0xffffffffu, v, j);
        v = ((lane & j) == 0) ? umax64(v, p) : umin64(v, p);
    }
    return v;
}

__global__ void __launch_bounds__(256) topk_kernel(const u64* __restrict__ keys, const u32* __restrict__ count,
                                                   const u64* __restrict__ seg_start, const u64* __restrict__ n_seg_dev,
                                                   int64_t n_rows, int k, int32_t* __restrict__ out_aid_x,
                                                   int32_t* __restrict__ out_nvalid, int32_t* __restrict__ out_aid_y,
                                                   int32_t* __restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_seg = (int64_t)*n_seg_dev;          // written by the segment-head scan just before this launch
    for (int64_t s = warp0; s < n_seg; s += n_warps) {
        const int64_t a = (int64_t)seg_start[s];
        const int64_t b = (s + 1 < n_seg) ? (int64_t)seg_start[s + 1] : n_rows;
        u64 best = 0;                                   // lane l: l-th largest so far (0 = empty)
        for (int64_t i = a; i < b; i += 32) {
            const int64_t r = i + lane;
            u64 c = 0;
            if (r < b) c = ((u64)count[r] << 32) | (u64)(0xFFFFFFFFu - (u32)(keys[r] & 0xFFFFFFFFu));
            const u64 kth = __shfl_sync(0xffffffffu, best, k - 1);
            if (__ballot_sync(0xffffffffu, c > kth) == 0) continue;
            c = warp_sort_desc(c, lane);
            const u64 rev = __shfl_sync(0xffffffffu, c, 31 - lane);
            best = warp_bitonic_merge_desc(umax64(best, rev), lane);
        }
        const int64_t len = b - a;
        const int nv = (int)(len < k ? len : k);
        if (lane == 0) {
            out_aid_x[s] = (int32_t)(keys[a] >> 32);
            out_nvalid[s] = nv;
        }
        if (lane < k) {
            const bool ok = lane < nv;
            out_aid_y[s * k + lane] = ok ? (int32_t)(0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu)) : -1;
            out_cnt[s * k + lane] = ok ? (int32_t)min((u64)0x7FFFFFFFull, best >> 32) : 0;
        }
    }
}

void free_topk(ottocov_ctx* ctx) {
    dev_free(ctx, ctx->topk_aid_x); dev_free(ctx, ctx->topk_nvalid);
    dev_free(ctx, ctx->topk_aid_y); dev_free(ctx, ctx->topk_cnt);
    ctx->topk_aid_x = ctx->topk_nvalid = ctx->topk_aid_y = ctx->topk_cnt = nullptr;
    ctx->topk_n = 0; ctx->topk_k = 0;
}

void topk_impl(ottocov_ctx* ctx, const ottocov_table* t, int k) {
    if (k < 1 || k > 32) COV_THROW(OTTOCOV_ERR_ARG, "top-k supports 1 <= k <= 32 (got %d)", k);
    free_topk(ctx);
    ctx->topk_k = k;
    if (t->n == 0) return;
    // number of distinct aids is bounded by min(rows, 2^aid_bits)
    int64_t cap = t->n;
    if (t->aid_bits < 40 && ((int64_t)1 << t->aid_bits) < cap) cap = (int64_t)1 << t->aid_bits;
    DevBuf<u64> seg_start(ctx, cap);
    SegmentHeads f;
    f.keys = t->keys; f.seg_start = seg_start.p;
    // The number of segments stays on the device: the top-K kernel is enqueued right behind the scan with a grid
    // and output buffers sized for the bound, and reads the count itself -- one host round trip less per call.
    scan_apply(ctx, OTTOCOV_K_TOPK, f, t->n, nullptr, 8.0 * t->n);
    DevBuf<int32_t> ax(ctx, cap), nv(ctx, cap), ay(ctx, cap * k), ac(ctx, cap * k);
    int grid = (int)imin64(ceil_div64(cap, 8), (int64_t)ctx->num_sms * 32);
    COV_LAUNCH(ctx, OTTOCOV_K_TOPK, 12.0 * t->n, topk_kernel, grid, 256, 0,
               t->keys, t->count, seg_start.p, (const u64*)ctx->scan_totals, t->n, k, ax.p, nv.p, ay.p, ac.p);
    u64 n_seg = 0;
    cov_readback(ctx, &n_seg, ctx->scan_totals, sizeof(u64));
    ctx->stats[OTTOCOV_K_TOPK].algo_bytes += 16.0 * (double)n_seg + 8.0 * k * (double)n_seg;
    ctx->topk_aid_x = ax.take(); ctx->topk_nvalid = nv.take();
    ctx->topk_aid_y = ay.take(); ctx->topk_cnt = ac.take();
    ctx->topk_n = (int64_t)n_seg;
}

// ---- candidate lookup: the consumer's join of session aids with the per-aid top-N rows ---------------------
// (model/retrieve.py:75-91, `df_aids[['aid']].unique().join(df_count[['aid','aid_next']], on='aid')`): for
// every query aid, its row of the top-K result (binary search over the ascending aid_x column) or nothing.
__global__ void __launch_bounds__(256) topk_lookup_kernel(const int32_t* __restrict__ q, int64_t n,
                                                          const int32_t* __restrict__ aid_x, int64_t n_seg,
                                                          const int32_t* __restrict__ nvalid,
                                                          const int32_t* __restrict__ aid_y,
                                                          const int32_t* __restrict__ cnt, int k,
                                                          int32_t* __restrict__ o_nv, int32_t* __restrict__ o_y,
                                                          int32_t* __restrict__ o_c) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per query
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int32_t a = q[w];
    int64_t lo = 0, hi = n_seg;                      // first row with aid_x >= a
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (aid_x[mid] < a) lo = mid + 1; else hi = mid;
    }
    const bool hit = lo < n_seg && aid_x[lo] == a;
    const int nv = hit ? nvalid[lo] : 0;
    if (lane == 0) o_nv[w] = nv;
    if (lane < k) {
        o_y[w * k + lane] = (lane < nv) ? aid_y[lo * k + lane] : -1;
        o_c[w * k + lane] = (lane < nv) ? cnt[lo * k + lane] : 0;
    }
}

void topk_lookup_impl(ottocov_ctx* ctx, const int32_t* aids, int64_t n, int where, int32_t* n_valid, int32_t* aid_y,
                      int32_t* cnt) {
    if (ctx->topk_k == 0) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_topk_lookup before ottocov_table_topk");
    if (n == 0) return;
    const int k = ctx->topk_k;
    DevBuf<int32_t> dq, dnv, dy, dc;
    const int32_t* q = aids;
    int32_t *pnv = n_valid, *py = aid_y, *pc = cnt;
    if (where == OTTOCOV_HOST) {
        dq.alloc(ctx, n); dnv.alloc(ctx, n); dy.alloc(ctx, n * k); dc.alloc(ctx, n * k);
        CUDA_CHECK(cudaMemcpyAsync(dq.p, aids, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        q = dq.p; pnv = dnv.p; py = dy.p; pc = dc.p;
    }
    COV_LAUNCH(ctx, OTTOCOV_K_TOPK, 4.0 * n + 8.0 * k * n, topk_lookup_kernel, (unsigned)ceil_div64(n * 32, 256), 256, 0, q, n,
               ctx->topk_aid_x, ctx->topk_n, ctx->topk_nvalid, ctx->topk_aid_y, ctx->topk_cnt, k, pnv, py, pc);
    if (where == OTTOCOV_HOST) {
        CUDA_CHECK(cudaMemcpyAsync(n_valid, pnv, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_y, py, n * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(cnt, pc, n * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
