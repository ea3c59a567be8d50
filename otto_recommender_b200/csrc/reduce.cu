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
#include <type_traits>

// ---- scan / look-back state ---------------------------------------------------------------------------
void scan_state_prepare(ottocov_ctx* ctx, size_t status_words, u32* epoch_out) {
    if (!ctx->scan_ticket) {
        CUDA_CHECK(cudaMalloc((void**)&ctx->scan_ticket, 64));
        CUDA_CHECK(cudaMalloc((void**)&ctx->scan_totals, 8 * sizeof(u64)));
    }
    if (status_words > ctx->scan_status_words) {
        if (ctx->scan_status) CUDA_CHECK(cudaFreeAsync(ctx->scan_status, ctx->stream));
        ctx->scan_status = nullptr;
        ctx->scan_status_words = 0;
        const size_t cap = status_words + status_words / 4 + 1024;
        CUDA_CHECK(cudaMallocAsync((void**)&ctx->scan_status, cap * sizeof(u64), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(ctx->scan_status, 0, cap * sizeof(u64), ctx->stream));
        ctx->scan_status_words = cap;
        ctx->scan_epoch = 0;
    }
    if (ctx->scan_epoch >= 63) {
        CUDA_CHECK(cudaMemsetAsync(ctx->scan_status, 0, ctx->scan_status_words * sizeof(u64), ctx->stream));
        ctx->scan_epoch = 0;
    }
    *epoch_out = ++ctx->scan_epoch;
    CUDA_CHECK(cudaMemsetAsync(ctx->scan_ticket, 0, 64, ctx->stream));
}

// ---- run-length / segmented-sum reduce of sorted keys, fused with the count threshold ---------------------
// One pass over the sorted keys (read 8 B/key, + 4 B/key payload when summing); writes 12 B per KEPT
// run only.  Per tile of 2048 keys:
//   1. heads (key != previous key) and ends (key != next key) per 32-key row by ballot; the running
//      sum since the last head needs no shuffles when counting (lane - head lane + 1);
//   2. the tile publishes (has_head, sum after its last head).  A run that began in an earlier tile
//      gets its carry-in by walking back to the nearest tile that has a head -- a pure function of the
//      data, so there is no prefix chain to wait on (almost always one step);
//   3. every run END now knows the run's total; keep = total >= min_count;
//   4. kept ends are compacted with the usual chained scan (aggregate / inclusive-prefix look-back).
constexpr int RLE_THREADS = 256;
constexpr int RLE_WARPS = RLE_THREADS / 32;
constexpr u64 RLE_HAS_HEAD = 1ull << 55;
constexpr u64 RLE_VALUE_MASK = (1ull << 55) - 1;

// ITEMS rows of 32 keys per warp: 16 when counting (tile = 4096 keys = 32 KB; 4 CTAs/SM keep 128 KB of
// reads in flight per SM, which is what hides the look-back round trips), 8 when summing a payload.
template <bool HAS_VALS, int ITEMS, int MINB>
__global__ void __launch_bounds__(RLE_THREADS, MINB)
rle_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals, int64_t n, int64_t n_tiles, u32 min_count,
           int sym, u64* __restrict__ out_keys, u32* __restrict__ out_count, u64* status_tail, u64* status_keep,
           u32* ticket, u32 epoch, u64* __restrict__ totals) {
    typedef typename std::conditional<HAS_VALS, u64, u32>::type sum_t;      // in-tile sums: counts fit 32 bits
    constexpr int TILE = RLE_THREADS * ITEMS;
    __shared__ u64 s_wsum[RLE_WARPS];      // sum since the warp chunk's last head (whole chunk if none)
    __shared__ u32 s_whead[RLE_WARPS];     // chunk contains a head
    __shared__ u64 s_wcarry[RLE_WARPS];    // carry-in for runs that began before the chunk
    __shared__ u32 s_wkeep[RLE_WARPS];
    __shared__ u64 s_outbase;
    __shared__ u32 s_tile;
    __shared__ u32 s_first_head;           // the tile's first key starts a run: no carry-in to fetch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t cb = tile * TILE + (int64_t)warp * 32 * ITEMS;    // chunk base
    const u32 lt = lanemask_lt();
    const u32 le = lt | (1u << lane);

    u64 key[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int64_t idx = cb + r * 32 + lane;
        key[r] = (idx < n) ? __ldcs(keys + idx) : 0ull;
    }
    // neighbours across the chunk edges
    u64 edge_prev = 0, edge_next = 0;
    if (lane == 0 && cb > 0 && cb <= n) edge_prev = keys[cb - 1];
    if (lane == 31 && cb + 32 * ITEMS < n) edge_next = keys[cb + 32 * ITEMS];

    sum_t s[ITEMS];                        // inclusive sum since the last head at or before this element
    u32 end_bits = 0, open_bits = 0;       // bit r: element is a run end / its run began before the chunk
    u64 carry = 0;                         // sum since the last head, at the end of the previous row
    bool seen_head = false;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int64_t idx = cb + r * 32 + lane;
        const bool valid = idx < n;
        u64 kprev = __shfl_up_sync(0xffffffffu, key[r], 1);
        u64 knext = __shfl_down_sync(0xffffffffu, key[r], 1);
        const u64 prow_last = __shfl_sync(0xffffffffu, key[r > 0 ? r - 1 : 0], 31);
        const u64 nrow_first = __shfl_sync(0xffffffffu, key[r < ITEMS - 1 ? r + 1 : r], 0);
        if (lane == 0) kprev = (r == 0) ? edge_prev : prow_last;
        if (lane == 31) knext = (r == ITEMS - 1) ? edge_next : nrow_first;
        const bool head = valid && (idx == 0 || key[r] != kprev);
        const bool end = valid && (idx == n - 1 || key[r] != knext);
        if (r == 0 && threadIdx.x == 0) s_first_head = head ? 1u : 0u;
        const u32 hm = __ballot_sync(0xffffffffu, head);
        const u32 m = hm & le;
        const bool has = m != 0;
        const int pl = 31 - __clz(m);                      // lane of the last head at or before this lane
        u64 si;
        if (!HAS_VALS) {
            si = has ? (u64)(lane - pl + 1) : carry + (u64)(lane + 1);
        } else {
            const u32 v = valid ? __ldcs(vals + idx) : 0u;
            u64 inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const u64 before_head = __shfl_sync(0xffffffffu, inc - v, has ? pl : 0);
            si = has ? inc - before_head : carry + inc;
        }
        s[r] = (sum_t)si;
        if (end) end_bits |= 1u << r;
        if (!has && !seen_head) open_bits |= 1u << r;
        carry = __shfl_sync(0xffffffffu, si, 31);
        seen_head = seen_head || (hm != 0);
    }
    if (lane == 0) { s_wsum[warp] = carry; s_whead[warp] = seen_head ? 1u : 0u; }
    __syncthreads();

    if (threadIdx.x == 0) {
        // tile aggregate: sum after the tile's last head (whole tile if it has none)
        u64 tail = 0; bool hh = false;
        for (int w = RLE_WARPS - 1; w >= 0; --w) {
            tail += s_wsum[w];
            if (s_whead[w]) { hh = true; break; }
        }
        const u64 tag = (u64)epoch << 56;
        st_volatile_u64(status_tail + tile, SC_FLAG_AGG | tag | (hh ? RLE_HAS_HEAD : 0ull) | (tail & RLE_VALUE_MASK));
        // carry-in: only needed when the tile does not open with a head; walk back to the nearest tile
        // that contains one (tile 0 always does)
        u64 c_in = 0;
        if (!s_first_head) {
            for (int64_t t = tile - 1; t >= 0; --t) {
                u64 x;
                while (true) {
                    x = ld_volatile_u64(status_tail + t);
                    if ((u32)((x >> 56) & 0x3F) == epoch && (x >> 62) != 0) break;
                    __nanosleep(40);
                }
                c_in += x & RLE_VALUE_MASK;
                if (x & RLE_HAS_HEAD) break;
            }
        }
        u64 cur = c_in;
        for (int w = 0; w < RLE_WARPS; ++w) {
            s_wcarry[w] = cur;
            cur = s_whead[w] ? s_wsum[w] : cur + s_wsum[w];
        }
    }
    __syncthreads();

    // totals at the run ends, threshold
    const u64 wc = s_wcarry[warp];
    u32 keep_bits = 0, wkeep = 0;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        u64 total = (u64)s[r] + (((open_bits >> r) & 1u) ? wc : 0ull);
        if (sym && (u32)(key[r] >> 32) == (u32)key[r]) total *= 2;      // (a, a): both orders of each event pair
        const bool keep = ((end_bits >> r) & 1u) && total >= (u64)min_count;
        wkeep += __popc(__ballot_sync(0xffffffffu, keep));
        if (keep) keep_bits |= 1u << r;
    }
    if (lane == 0) s_wkeep[warp] = wkeep;
    __syncthreads();
    if (warp == 0) {
        u64 k = (lane < RLE_WARPS) ? (u64)s_wkeep[lane] : 0ull;
        k = warp_sum_u64(k);
        const u64 pre = warp_lookback_sum(status_keep + tile, 1, tile, k, epoch);
        if (lane == 0) {
            s_outbase = pre;
            if (tile == n_tiles - 1) totals[0] = pre + k;
        }
    }
    __syncthreads();
    u64 ob = s_outbase;
    for (int w = 0; w < warp; ++w) ob += s_wkeep[w];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const bool keep = (keep_bits >> r) & 1u;
        const u32 km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            u64 total = (u64)s[r] + (((open_bits >> r) & 1u) ? wc : 0ull);
            if (sym && (u32)(key[r] >> 32) == (u32)key[r]) total *= 2;
            const u64 o = ob + __popc(km & lt);
            out_keys[o] = key[r];
            out_count[o] = (u32)(total > 0xFFFFFFFFull ? 0xFFFFFFFFull : total);
        }
        ob += __popc(km);
    }
}

// shrink an over-allocated result when most of it is unused (keeps long-lived tables small)
template <class T>
static T* shrink_to_fit(ottocov_ctx* ctx, DevBuf<T>& buf, int64_t used) {
    if ((int64_t)buf.n <= 2 * used + 1024) return buf.take();
    DevBuf<T> small(ctx, used);
    if (used > 0)
        CUDA_CHECK(cudaMemcpyAsync(small.p, buf.p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return small.take();
}

void reduce_sorted(ottocov_ctx* ctx, const u64* keys, const u32* vals, int64_t n, u32 min_count, bool sym,
                   u64** out_keys, u32** out_count, int64_t* n_out) {
    *out_keys = nullptr; *out_count = nullptr; *n_out = 0;
    if (n <= 0) return;
    const int items = vals ? 8 : 16;
    const int64_t n_tiles = ceil_div64(n, (int64_t)RLE_THREADS * items);
    DevBuf<u64> ukeys(ctx, n);             // upper bound: every key distinct (blocks come from the cache)
    DevBuf<u32> ucount(ctx, n);
    u32 epoch;
    scan_state_prepare(ctx, 2 * (size_t)n_tiles, &epoch);
    u64* st_tail = ctx->scan_status;
    u64* st_keep = ctx->scan_status + n_tiles;
    static int minb = 0;
    if (!minb) { const char* e = getenv("OTTOCOV_RLE_MINB"); minb = (e && atoi(e) == 3) ? 3 : 4; }
    if (vals)
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, 12.0 * n, (rle_kernel<true, 8, 2>), (unsigned)n_tiles, RLE_THREADS, 0, keys, vals, n,
                   n_tiles, min_count, sym ? 1 : 0, ukeys.p, ucount.p, st_tail, st_keep, ctx->scan_ticket, epoch, ctx->scan_totals);
    else if (minb == 3)
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n, (rle_kernel<false, 16, 3>), (unsigned)n_tiles, RLE_THREADS, 0, keys, vals, n,
                   n_tiles, min_count, sym ? 1 : 0, ukeys.p, ucount.p, st_tail, st_keep, ctx->scan_ticket, epoch, ctx->scan_totals);
    else
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n, (rle_kernel<false, 16, 4>), (unsigned)n_tiles, RLE_THREADS, 0, keys, vals, n,
                   n_tiles, min_count, sym ? 1 : 0, ukeys.p, ucount.p, st_tail, st_keep, ctx->scan_ticket, epoch, ctx->scan_totals);
    u64 rows = 0;
    cov_readback(ctx, &rows, ctx->scan_totals, sizeof(u64));
    ctx->stats[OTTOCOV_K_RLE].algo_bytes += 12.0 * (double)rows;
    *out_keys = shrink_to_fit(ctx, ukeys, (int64_t)rows);
    *out_count = shrink_to_fit(ctx, ucount, (int64_t)rows);
    *n_out = (int64_t)rows;
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
    cov_readback(ctx, out, d.p, 3 * sizeof(u64));
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
        reduce_sorted(ctx, k, v, n, 1, false, &out->keys, &out->count, &out->n);
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

// ---- mirror of a half table -------------------------------------------------------------------------------
struct MirrorRows {
    static constexpr int NC = 1;
    const u64* keys;
    const u32* count;
    int64_t n;          // rows copied through unchanged in front of the mirrored ones (0: transpose only)
    u64* okeys;
    u32* ocount;
    __device__ u64 value(int64_t i) const { return ((u32)(keys[i] >> 32) != (u32)keys[i]) ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        const u64 k = keys[i];
        const u32 c = count[i];
        if (n) { okeys[i] = k; ocount[i] = c; }
        if (v) { okeys[n + pre[0]] = (k << 32) | (k >> 32); ocount[n + pre[0]] = c; }
    }
};

ottocov_table* mirror_table_impl(ottocov_ctx* ctx, const ottocov_table* half, bool transpose_only) {
    ottocov_table* out = new ottocov_table();
    out->aid_bits = half->aid_bits;
    const int64_t n = half->n;
    if (n == 0) return out;
    try {
        DevBuf<u64> keys(ctx, 2 * n), kalt;
        DevBuf<u32> count(ctx, 2 * n), calt;
        MirrorRows f;
        f.keys = half->keys; f.count = half->count; f.n = transpose_only ? 0 : n; f.okeys = keys.p; f.ocount = count.p;
        u64 tot[1];
        scan_apply(ctx, OTTOCOV_K_ORDER, f, n, tot, 36.0 * n);
        const int64_t m = f.n + (int64_t)tot[0];
        if (m == 0) return out;
        kalt.alloc(ctx, m); calt.alloc(ctx, m);
        u64* k = keys.p; u64* ka = kalt.p; u32* v = count.p; u32* va = calt.p;
        BitField fields[2] = {{0, half->aid_bits}, {32, 32 + half->aid_bits}};
        radix_sort_pairs(ctx, k, ka, v, va, m, fields, 2);
        // the sorted rows may sit in either half of the double buffer: hand that half to the table
        if (k == keys.p) { out->keys = keys.take(); out->count = count.take(); }
        else { out->keys = kalt.take(); out->count = calt.take(); }
        out->n = m;
    } catch (...) {
        delete out;
        throw;
    }
    return out;
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
        DevBuf<u64> ok(ctx, t->n);
        DevBuf<u32> oc(ctx, t->n);
        KeepAtLeast f;
        f.keys = t->keys; f.count = t->count; f.min_count = min_count; f.okeys = ok.p; f.ocount = oc.p;
        u64 tot[1];
        scan_apply(ctx, OTTOCOV_K_FILTER, f, t->n, tot, 4.0 * t->n);
        const int64_t m = (int64_t)tot[0];
        ctx->stats[OTTOCOV_K_FILTER].algo_bytes += 20.0 * (double)m;
        out->keys = shrink_to_fit(ctx, ok, m);
        out->count = shrink_to_fit(ctx, oc, m);
        out->n = m;
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

// dest rank stamped into key bits 56..63 (valid while aids stay below 2^24), counted per destination
__global__ void __launch_bounds__(256) stamp_dest_kernel(const u64* __restrict__ keys, int64_t n, u32 n_ranks,
                                                         u64* __restrict__ out, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_c[256];
    s_c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 k = keys[i];
        const u32 d = hash_dest((u32)(k >> 32), n_ranks);
        out[i] = k | ((u64)d << 56);
        atomicAdd(&s_c[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

__global__ void __launch_bounds__(256) unstamp_copy_kernel(const u64* __restrict__ keys, const u32* __restrict__ cnt,
                                                           int64_t n, u64* __restrict__ okeys, u32* __restrict__ ocnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    okeys[i] = keys[i] & 0x00FFFFFFFFFFFFFFull;
    ocnt[i] = cnt[i];
}

void partition_table_impl(ottocov_ctx* ctx, const ottocov_table* t, int n_ranks, u64* keys_out,
                          u32* count_out, int64_t* rows_per_dest) {
    const int64_t n = t->n;
    for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = 0;
    if (n == 0) return;
    if (t->aid_bits <= 24 && n_ranks <= 256) {
        // one stamping pass + ONE stable radix pass on the destination bits (one host sync in total)
        DevBuf<u64> k0(ctx, n), k1(ctx, n);
        DevBuf<u32> c0(ctx, n), c1(ctx, n);
        DevBuf<unsigned long long> cnt(ctx, 256);
        CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, 256 * sizeof(unsigned long long), ctx->stream));
        COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 16.0 * n, stamp_dest_kernel, (int)imin64(ceil_div64(n, 1024), (int64_t)ctx->num_sms * 8),
                   256, 0, t->keys, n, (u32)n_ranks, k0.p, cnt.p);
        CUDA_CHECK(cudaMemcpyAsync(c0.p, t->count, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        unsigned long long h[256];
        cov_readback(ctx, h, cnt.p, 256 * sizeof(unsigned long long));
        int bits = 1;
        while ((1 << bits) < n_ranks) ++bits;
        BitField f[1] = {{56, 56 + bits}};
        u64* k = k0.p; u64* ka = k1.p; u32* v = c0.p; u32* va = c1.p;
        radix_sort_pairs(ctx, k, ka, v, va, n, f, 1);
        COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 24.0 * n, unstamp_copy_kernel, (unsigned)ceil_div64(n, 256), 256, 0, k, v, n,
                   keys_out, count_out);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = (int64_t)h[r];
        return;
    }
    u64 base = 0;                    // general fallback: one compaction per destination
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
