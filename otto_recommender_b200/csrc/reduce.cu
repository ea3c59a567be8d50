// reduce.cu -- everything that happens to (aid, aid_next, count) tables after the sort.
//
//   reduce_sorted      sorted keys -> distinct keys + run lengths          groupby(...).count()  count_co_events.py:70-71
//                      sorted (key, count) -> distinct keys + summed count  groupby(...).sum()    count_co_events.py:168
//   merge_tables_impl  concat + sort + reduce                              concat_files_w_stats   count_co_events.py:112-115,168
//   filter_table_impl  count >= threshold                                   count_co_events.py:131-132,156,172
//   fetch_table_impl   unpack; optional global order by count desc + head   count_co_events.py:173-175
//   partition          stable split by hash(aid) % R for the multi-GPU exchange (no reference counterpart)
#include "internal.cuh"
#include "scan.cuh"

// ---- run heads ------------------------------------------------------------------------------------------
template <class IdxT>
struct RunHeads {
    static constexpr int NC = 1;
    const u64* keys;
    u64* ukeys;
    IdxT* ustart;
    __device__ u64 value(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        ukeys[pre[0]] = keys[i];
        ustart[pre[0]] = (IdxT)i;
    }
};

// count only: value() without outputs, used to size the result exactly
struct RunHeadsCount {
    static constexpr int NC = 1;
    const u64* keys;
    __device__ u64 value(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1ull : 0ull; }
    __device__ void apply(int64_t, u64, const u64*) const {}
};

template <class IdxT>
__global__ void __launch_bounds__(256) run_length_kernel(const IdxT* __restrict__ ustart, int64_t n_runs,
                                                         int64_t n, u32* __restrict__ count) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_runs) return;
    const u64 a = (u64)ustart[u];
    const u64 b = (u + 1 < n_runs) ? (u64)ustart[u + 1] : (u64)n;
    count[u] = (u32)(b - a);
}

template <class IdxT>
__global__ void __launch_bounds__(256) run_sum_kernel(const IdxT* __restrict__ ustart, int64_t n_runs,
                                                      int64_t n, const u32* __restrict__ vals,
                                                      u32* __restrict__ count) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_runs) return;
    const u64 a = (u64)ustart[u];
    const u64 b = (u + 1 < n_runs) ? (u64)ustart[u + 1] : (u64)n;
    u64 s = 0;
    for (u64 i = a; i < b; ++i) s += vals[i];
    count[u] = (u32)(s > 0xFFFFFFFFull ? 0xFFFFFFFFull : s);
}

// sums of tile totals only (pass 1 + 2 of the scan framework), to size outputs exactly
template <class F>
static u64 scan_count(ottocov_ctx* ctx, int family, const F& f, int64_t n, double bytes) {
    if (n <= 0) return 0;
    const int64_t n_tiles = ceil_div64(n, SCAN_TILE);
    DevBuf<u64> sums(ctx, (size_t)n_tiles + 1);
    COV_LAUNCH(ctx, family, bytes, (scan_reduce_kernel<F>), (unsigned)n_tiles, SCAN_THREADS, 0, f, n, n_tiles, sums.p);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, n_tiles * 16.0, scan_block_sums_kernel, 1, 1024, 0, sums.p, n_tiles, sums.p + n_tiles);
    u64 tot = 0;
    CUDA_CHECK(cudaMemcpyAsync(&tot, sums.p + n_tiles, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return tot;
}

template <class IdxT>
static void reduce_sorted_t(ottocov_ctx* ctx, const u64* keys, const u32* vals, int64_t n, int64_t n_runs,
                            u64* ukeys, u32* ucount) {
    DevBuf<IdxT> ustart(ctx, n_runs);
    RunHeads<IdxT> f;
    f.keys = keys; f.ukeys = ukeys; f.ustart = ustart.p;
    scan_apply(ctx, OTTOCOV_K_RLE, f, n, nullptr, 16.0 * n + (8.0 + sizeof(IdxT)) * n_runs);
    const unsigned grid = (unsigned)ceil_div64(n_runs, 256);
    if (vals)
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, sizeof(IdxT) * n_runs + 4.0 * n + 4.0 * n_runs, run_sum_kernel<IdxT>, grid, 256, 0,
                   ustart.p, n_runs, n, vals, ucount);
    else
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, sizeof(IdxT) * n_runs + 4.0 * n_runs, run_length_kernel<IdxT>, grid, 256, 0,
                   ustart.p, n_runs, n, ucount);
}

void reduce_sorted(ottocov_ctx* ctx, const u64* keys, const u32* vals, int64_t n, u64** out_keys,
                   u32** out_count, int64_t* n_out) {
    *out_keys = nullptr; *out_count = nullptr; *n_out = 0;
    if (n <= 0) return;
    RunHeadsCount fc; fc.keys = keys;
    const int64_t n_runs = (int64_t)scan_count(ctx, OTTOCOV_K_RLE, fc, n, 8.0 * n);
    DevBuf<u64> ukeys(ctx, n_runs);
    DevBuf<u32> ucount(ctx, n_runs);
    if (n < (int64_t)0xFFFFFFFFll) reduce_sorted_t<u32>(ctx, keys, vals, n, n_runs, ukeys.p, ucount.p);
    else reduce_sorted_t<u64>(ctx, keys, vals, n, n_runs, ukeys.p, ucount.p);
    *out_keys = ukeys.take();
    *out_count = ucount.take();
    *n_out = n_runs;
}

// ---- small reductions -------------------------------------------------------------------------------------
// out[0] |= OR of keys, out[1] += sum of count, out[2] = max(count)
__global__ void __launch_bounds__(256) table_stats_kernel(const u64* __restrict__ keys,
                                                          const u32* __restrict__ count, int64_t n,
                                                          u64* __restrict__ out) {
    u64 o = 0, s = 0, m = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (keys) o |= keys[i];
        if (count) { const u64 c = count[i]; s += c; m = c > m ? c : m; }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, k);
        s += __shfl_xor_sync(0xffffffffu, s, k);
        const u64 mm = __shfl_xor_sync(0xffffffffu, m, k);
        m = mm > m ? mm : m;
    }
    if ((threadIdx.x & 31) == 0) {
        if (o) atomicOr(&out[0], o);
        if (s) atomicAdd(&out[1], s);
        atomicMax(&out[2], m);
    }
}

static void table_stats(ottocov_ctx* ctx, const u64* keys, const u32* count, int64_t n, u64 out[3]) {
    out[0] = out[1] = out[2] = 0;
    if (n <= 0) return;
    DevBuf<u64> d(ctx, 3);
    CUDA_CHECK(cudaMemsetAsync(d.p, 0, 3 * sizeof(u64), ctx->stream));
    int grid = (int)imin64(ceil_div64(n, 256), (int64_t)ctx->num_sms * 16);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 12.0 * n, table_stats_kernel, grid, 256, 0, keys, count, n, d.p);
    CUDA_CHECK(cudaMemcpyAsync(out, d.p, 3 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

static int bits_of(u64 v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

static int aid_bits_from_or(u64 o) {
    int b = bits_of(o >> 32), c = bits_of(o & 0xFFFFFFFFull);
    int r = b > c ? b : c;
    return r > 0 ? r : 1;
}

int64_t table_total_impl(ottocov_ctx* ctx, const ottocov_table* t) {
    u64 st[3];
    table_stats(ctx, nullptr, t->count, t->n, st);
    return (int64_t)st[1];
}

// ---- sort + reduce of an arbitrary (key, count) multiset: owns and consumes keys/count buffers --------------
static ottocov_table* sort_reduce_pairs(ottocov_ctx* ctx, DevBuf<u64>& keys, DevBuf<u32>& count, int64_t n,
                                        int aid_bits) {
    ottocov_table* out = new ottocov_table();
    out->aid_bits = aid_bits;
    if (n <= 0) return out;
    try {
        DevBuf<u64> kalt(ctx, n);
        DevBuf<u32> valt(ctx, n);
        u64* k = keys.p; u64* ka = kalt.p; u32* v = count.p; u32* va = valt.p;
        BitField fields[2] = {{0, aid_bits}, {32, 32 + aid_bits}};
        radix_sort_pairs(ctx, k, ka, v, va, n, fields, 2);
        reduce_sorted(ctx, k, v, n, &out->keys, &out->count, &out->n);
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

ottocov_table* merge_tables_impl(ottocov_ctx* ctx, ottocov_table* const* tabs, int n_tabs) {
    int64_t n = 0;
    int aid_bits = 1;
    for (int i = 0; i < n_tabs; ++i) {
        if (!tabs[i]) COV_THROW(OTTOCOV_ERR_ARG, "NULL table in merge");
        n += tabs[i]->n;
        aid_bits = tabs[i]->aid_bits > aid_bits ? tabs[i]->aid_bits : aid_bits;
    }
    DevBuf<u64> keys(ctx, n);
    DevBuf<u32> count(ctx, n);
    int64_t o = 0;
    for (int i = 0; i < n_tabs; ++i) {
        if (tabs[i]->n == 0) continue;
        CUDA_CHECK(cudaMemcpyAsync(keys.p + o, tabs[i]->keys, tabs[i]->n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(count.p + o, tabs[i]->count, tabs[i]->n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        o += tabs[i]->n;
    }
    return sort_reduce_pairs(ctx, keys, count, n, aid_bits);
}

__global__ void __launch_bounds__(256) pack_keys_kernel(const int32_t* __restrict__ aid,
                                                        const int32_t* __restrict__ aid_next, int64_t n,
                                                        u64* __restrict__ keys, u32* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t a = aid[i], b = aid_next[i];
    if (a < 0 || b < 0) *bad = 1;
    keys[i] = ((u64)(u32)a << 32) | (u64)(u32)b;
}

ottocov_table* table_from_arrays_impl(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next,
                                      const u32* count, int64_t n, int where) {
    if (n == 0) return new ottocov_table();
    DevBuf<int32_t> da, db;
    DevBuf<u32> dc(ctx, n);
    const int32_t* pa = aid; const int32_t* pb = aid_next;
    if (where == OTTOCOV_HOST) {
        da.alloc(ctx, n); db.alloc(ctx, n);
        CUDA_CHECK(cudaMemcpyAsync(da.p, aid, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(db.p, aid_next, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(dc.p, count, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        pa = da.p; pb = db.p;
    } else {
        CUDA_CHECK(cudaMemcpyAsync(dc.p, count, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    DevBuf<u64> keys(ctx, n);
    DevBuf<u32> bad(ctx, 1);
    CUDA_CHECK(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 16.0 * n, pack_keys_kernel, (unsigned)ceil_div64(n, 256), 256, 0, pa, pb, n, keys.p, bad.p);
    u32 hbad = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hbad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    u64 st[3];
    table_stats(ctx, keys.p, nullptr, n, st);     // synchronises
    if (hbad) COV_THROW(OTTOCOV_ERR_DATA, "negative aid in table");
    return sort_reduce_pairs(ctx, keys, dc, n, aid_bits_from_or(st[0]));
}

ottocov_table* table_from_packed_impl(ottocov_ctx* ctx, const u64* keys_in, const u32* count, int64_t n, int where) {
    if (n == 0) return new ottocov_table();
    DevBuf<u64> keys(ctx, n);
    DevBuf<u32> dc(ctx, n);
    const cudaMemcpyKind kind = (where == OTTOCOV_HOST) ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CUDA_CHECK(cudaMemcpyAsync(keys.p, keys_in, n * 8, kind, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(dc.p, count, n * 4, kind, ctx->stream));
    u64 st[3];
    table_stats(ctx, keys.p, nullptr, n, st);
    if ((st[0] >> 63) || ((st[0] >> 31) & 1)) COV_THROW(OTTOCOV_ERR_DATA, "packed key with a negative aid");
    return sort_reduce_pairs(ctx, keys, dc, n, aid_bits_from_or(st[0]));
}

// ---- threshold filter -------------------------------------------------------------------------------------
struct KeepAtLeast {
    static constexpr int NC = 1;
    const u64* keys;
    const u32* count;
    u32 min_count;
    u64* okeys;
    u32* ocount;
    __device__ u64 value(int64_t i) const { return count[i] >= min_count ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        okeys[pre[0]] = keys[i];
        ocount[pre[0]] = count[i];
    }
};

ottocov_table* filter_table_impl(ottocov_ctx* ctx, const ottocov_table* t, u32 min_count) {
    ottocov_table* out = new ottocov_table();
    out->aid_bits = t->aid_bits;
    if (t->n == 0) return out;
    try {
        KeepAtLeast f;
        f.keys = t->keys; f.count = t->count; f.min_count = min_count; f.okeys = nullptr; f.ocount = nullptr;
        const int64_t m = (int64_t)scan_count(ctx, OTTOCOV_K_FILTER, f, t->n, 4.0 * t->n);
        if (m > 0) {
            DevBuf<u64> ok(ctx, m);
            DevBuf<u32> oc(ctx, m);
            f.okeys = ok.p; f.ocount = oc.p;
            scan_apply(ctx, OTTOCOV_K_FILTER, f, t->n, nullptr, 8.0 * t->n + 12.0 * t->n + 12.0 * m);
            out->keys = ok.take(); out->count = oc.take(); out->n = m;
        }
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

// ---- fetch --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_kernel(const u64* __restrict__ keys, const u32* __restrict__ count,
                                                     const u32* __restrict__ perm, int64_t n,
                                                     int32_t* __restrict__ aid, int32_t* __restrict__ aid_next,
                                                     int32_t* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = perm ? (int64_t)perm[i] : i;
    const u64 k = keys[j];
    const u32 c = count[j];
    aid[i] = (int32_t)(k >> 32);
    aid_next[i] = (int32_t)(k & 0xFFFFFFFFu);
    cnt[i] = (int32_t)(c > 0x7FFFFFFFu ? 0x7FFFFFFFu : c);   // cast(pl.Int32), count_co_events.py:175
}

__global__ void __launch_bounds__(256) order_keys_kernel(const u32* __restrict__ count, int64_t n, u32 maxc,
                                                         u64* __restrict__ okeys, u32* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    okeys[i] = (u64)(maxc - count[i]);
    idx[i] = (u32)i;
}

void fetch_table_impl(ottocov_ctx* ctx, const ottocov_table* t, int order, int64_t head, int32_t* aid,
                      int32_t* aid_next, int32_t* count, int64_t cap, int where, int64_t* n_out) {
    int64_t m = t->n;
    if (head >= 0 && head < m) m = head;
    if (n_out) *n_out = m;
    if (m > cap) COV_THROW(OTTOCOV_ERR_CAPACITY, "fetch needs room for %lld rows, caller gave %lld", (long long)m, (long long)cap);
    if (m == 0) return;
    DevBuf<u32> idx, idx_alt;
    DevBuf<u64> okeys, okeys_alt;
    const u32* perm = nullptr;
    if (order == OTTOCOV_ORDER_COUNT_DESC) {
        if (t->n >= (int64_t)0xFFFFFFFFll) COV_THROW(OTTOCOV_ERR_ARG, "count-ordered fetch limited to 2^32-2 rows");
        u64 st[3];
        table_stats(ctx, nullptr, t->count, t->n, st);
        const u32 maxc = (u32)st[2];
        idx.alloc(ctx, t->n); idx_alt.alloc(ctx, t->n); okeys.alloc(ctx, t->n); okeys_alt.alloc(ctx, t->n);
        COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 16.0 * t->n, order_keys_kernel, (unsigned)ceil_div64(t->n, 256), 256, 0,
                   t->count, t->n, maxc, okeys.p, idx.p);
        BitField f[1] = {{0, bits_of((u64)maxc)}};
        u64* k = okeys.p; u64* ka = okeys_alt.p; u32* v = idx.p; u32* va = idx_alt.p;
        radix_sort_pairs(ctx, k, ka, v, va, t->n, f, 1);    // stable: ties stay in (aid, aid_next) order
        perm = v;
    } else if (order != OTTOCOV_ORDER_KEY) {
        COV_THROW(OTTOCOV_ERR_ARG, "unknown order %d", order);
    }
    DevBuf<int32_t> ta, tb, tc;
    int32_t *pa = aid, *pb = aid_next, *pc = count;
    if (where == OTTOCOV_HOST) {
        ta.alloc(ctx, m); tb.alloc(ctx, m); tc.alloc(ctx, m);
        pa = ta.p; pb = tb.p; pc = tc.p;
    }
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 28.0 * m, unpack_kernel, (unsigned)ceil_div64(m, 256), 256, 0, t->keys, t->count, perm, m, pa, pb, pc);
    if (where == OTTOCOV_HOST) {
        CUDA_CHECK(cudaMemcpyAsync(aid, pa, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_next, pb, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(count, pc, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// ---- partition by destination rank ---------------------------------------------------------------------------
struct ToDest {
    static constexpr int NC = 1;
    const u64* keys;
    const u32* count;
    u32 n_ranks, dest;
    u64* okeys;
    u32* ocount;
    u64 base;
    __device__ u64 value(int64_t i) const { return hash_dest((u32)(keys[i] >> 32), n_ranks) == dest ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        okeys[base + pre[0]] = keys[i];
        ocount[base + pre[0]] = count[i];
    }
};

void partition_table_impl(ottocov_ctx* ctx, const ottocov_table* t, int n_ranks, u64* keys_out,
                          u32* count_out, int64_t* rows_per_dest) {
    u64 base = 0;
    for (int r = 0; r < n_ranks; ++r) {
        ToDest f;
        f.keys = t->keys; f.count = t->count; f.n_ranks = (u32)n_ranks; f.dest = (u32)r;
        f.okeys = keys_out; f.ocount = count_out; f.base = base;
        u64 tot[1] = {0};
        scan_apply(ctx, OTTOCOV_K_PARTITION, f, t->n, tot, 16.0 * t->n + 12.0 * t->n / n_ranks);
        rows_per_dest[r] = (int64_t)tot[0];
        base += tot[0];
    }
}
