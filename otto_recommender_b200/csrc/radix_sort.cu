// radix_sort.cu -- single-sweep LSD radix sort of 64-bit keys (+ optional 32-bit payload).
//
// This is the engine behind the reference's groupby(['aid','aid_next']) (model/count_co_events.py
// :70-71, :168): co-event pairs are packed into u64 keys, sorted, then run-length reduced.
//
// Shape (per pass: read n keys once, write n keys once = 16 B/key, the HBM floor for a
// distribution pass):
//   * one up-front kernel builds the digit histograms of ALL passes in a single read;
//   * each pass is one kernel.  A CTA takes the next tile (atomic ticket, so a CTA only ever waits
//     for tiles that already started), ranks its 4096 keys by digit with warp match-any, publishes
//     its per-digit tile counts, resolves the per-digit exclusive prefix over earlier tiles by
//     decoupled look-back on epoch-tagged 64-bit status words, stages the keys in shared memory in
//     digit order and writes them out as coalesced per-digit runs.
//   * status word = flag(2) | epoch(6) | value(56).  The epoch changes every pass, so the status
//     array is never re-zeroed between passes.
// Only the significant bit fields are sorted: for co-event keys that is 2 x aid_bits (42 bits for
// 1.8 M aids -> 6 passes of 7 bits), not 64.
#include "internal.cuh"

constexpr int RS_THREADS = 256;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 4096 keys = 32 KB
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_MAX_BITS;
constexpr int RS_MAX_PASSES = 16;

constexpr u64 ST_VALUE_MASK = (1ull << 56) - 1;
constexpr u64 ST_FLAG_AGG = 1ull << 62;
constexpr u64 ST_FLAG_PREFIX = 2ull << 62;

struct PassList {
    int n;
    int shift[RS_MAX_PASSES];
    int bits[RS_MAX_PASSES];
};

// ---- histograms of every pass in one read ----------------------------------------------------------
__global__ void __launch_bounds__(256) rs_histogram_kernel(const u64* __restrict__ keys, int64_t n,
                                                           PassList pl, u64* __restrict__ ghist) {
    extern __shared__ u32 s_hist[];   // [pl.n][RS_RADIX]
    for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += blockDim.x) s_hist[j] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 k = ld_stream_u64(keys + i);
        for (int p = 0; p < pl.n; ++p) {
            const u32 d = (u32)(k >> pl.shift[p]) & ((1u << pl.bits[p]) - 1u);
            atomicAdd(&s_hist[p * RS_RADIX + d], 1u);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += blockDim.x) {
        const u32 c = s_hist[j];
        if (c) atomicAdd(&ghist[j], (u64)c);
    }
}

// in-place exclusive scan of each pass's 256 bins
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(u64* __restrict__ ghist) {
    __shared__ u64 s_warp[RS_RADIX / 32 + 1];
    u64* row = ghist + (size_t)blockIdx.x * RS_RADIX;
    u64 tot;
    const u64 v = row[threadIdx.x];
    const u64 ex = block_exclusive_scan<u64, RS_RADIX>(v, s_warp, &tot);
    row[threadIdx.x] = ex;
}

// ---- one distribution pass ---------------------------------------------------------------------------
template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS)
rs_onesweep_kernel(const u64* __restrict__ keys_in, u64* __restrict__ keys_out,
                   const u32* __restrict__ vals_in, u32* __restrict__ vals_out, int64_t n, int shift,
                   int bits, const u64* __restrict__ digit_base, u64* status, u32* ticket, u32 epoch) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_keys = reinterpret_cast<u64*>(s_raw);                              // [RS_TILE]
    u32* s_whist = reinterpret_cast<u32*>(s_keys + RS_TILE);                  // [RS_WARPS][RS_RADIX]
    u32* s_dstart = s_whist + RS_WARPS * RS_RADIX;                            // [RS_RADIX]
    u64* s_gbase = reinterpret_cast<u64*>(s_dstart + RS_RADIX);               // [RS_RADIX]
    u32* s_scan = reinterpret_cast<u32*>(s_gbase + RS_RADIX);                 // [RS_WARPS + 1] (+pad)
    u32* s_tile = s_scan + 16;                                                // [1] (+pad to 16)
    u32* s_vals = s_tile + 16;                                                // [RS_TILE] if HAS_VALS

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 radix = 1u << bits, mask = radix - 1u;

    if (tid == 0) *s_tile = atomicAdd(ticket, 1u);
    for (int j = tid; j < RS_WARPS * RS_RADIX; j += RS_THREADS) s_whist[j] = 0;
    __syncthreads();
    const int64_t tile = *s_tile;
    const int64_t tile_base = tile * RS_TILE;
    const int tile_n = (int)min((int64_t)RS_TILE, n - tile_base);

    // -- load (warp-striped: every access is a coalesced 256 B row) --------------------------------
    u64 key[RS_IPT];
    u32 val[RS_IPT];
    const int64_t lbase = tile_base + (int64_t)warp * 32 * RS_IPT + lane;
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const int64_t idx = lbase + i * 32;
        key[i] = (idx < n) ? ld_stream_u64(keys_in + idx) : ~0ull;
        if (HAS_VALS) val[i] = (idx < n) ? __ldcs(vals_in + idx) : 0u;
    }

    // -- rank inside the warp's 512-key segment by digit (stable) ----------------------------------
    u32 rnk[RS_IPT];
    u32* my_whist = s_whist + warp * RS_RADIX;
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const bool valid = (lbase + i * 32) < n;
        const u32 d = valid ? ((u32)(key[i] >> shift) & mask) : radix;   // `radix` never matches a digit
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader && valid) {
            old = my_whist[d];
            my_whist[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // -- per digit: exclusive offsets over warps, tile total ------------------------------------------
    u32 total = 0;
    if (tid < (int)radix) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const u32 c = s_whist[w * RS_RADIX + tid];
            s_whist[w * RS_RADIX + tid] = run;
            run += c;
        }
        total = run;
    }
    u32 blk_total;
    const u32 dstart = block_exclusive_scan<u32, RS_THREADS>(total, s_scan, &blk_total);

    // -- decoupled look-back: exclusive count of this digit over all earlier tiles --------------------
    if (tid < (int)radix) {
        s_dstart[tid] = dstart;
        const u64 tag = (u64)epoch << 56;
        u64* mine = status + (size_t)tile * radix + tid;
        u64 excl = 0;
        if (tile == 0) {
            st_volatile_u64(mine, ST_FLAG_PREFIX | tag | (u64)total);
        } else {
            st_volatile_u64(mine, ST_FLAG_AGG | tag | (u64)total);
            int64_t t = tile - 1;
            while (true) {
                const u64 v = ld_volatile_u64(status + (size_t)t * radix + tid);
                if ((u32)((v >> 56) & 0x3F) != epoch || (v >> 62) == 0) {
                    __nanosleep(20);
                    continue;
                }
                excl += v & ST_VALUE_MASK;
                if ((v >> 62) == 2) break;
                --t;
            }
            st_volatile_u64(mine, ST_FLAG_PREFIX | tag | (excl + (u64)total));
        }
        s_gbase[tid] = digit_base[tid] + excl - (u64)dstart;
    }
    __syncthreads();

    // -- stage keys in shared memory in digit order ------------------------------------------------------
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        if ((lbase + i * 32) < n) {
            const u32 d = (u32)(key[i] >> shift) & mask;
            const u32 pos = s_dstart[d] + my_whist[d] + rnk[i];
            s_keys[pos] = key[i];
            if (HAS_VALS) s_vals[pos] = val[i];
        }
    }
    __syncthreads();

    // -- coalesced per-digit runs out --------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const int j = tid + i * RS_THREADS;
        if (j < tile_n) {
            const u64 k = s_keys[j];
            const u32 d = (u32)(k >> shift) & mask;
            const u64 g = s_gbase[d] + (u64)j;
            keys_out[g] = k;
            if (HAS_VALS) vals_out[g] = s_vals[j];
        }
    }
}

static size_t rs_smem_bytes(bool has_vals) {
    size_t b = (size_t)RS_TILE * 8 + (size_t)RS_WARPS * RS_RADIX * 4 + RS_RADIX * 4 + RS_RADIX * 8 +
               16 * 4 + 16 * 4;
    if (has_vals) b += (size_t)RS_TILE * 4;
    return b;
}

static void ensure_sweep_state(ottocov_ctx* ctx, size_t words) {
    if (!ctx->sweep_ticket) {
        CUDA_CHECK(cudaMalloc((void**)&ctx->sweep_ticket, RS_MAX_PASSES * sizeof(u32)));
    }
    if (words > ctx->sweep_status_words) {
        // stream-ordered: earlier passes on this stream are done with the old array by the time
        // the free executes
        if (ctx->sweep_status) CUDA_CHECK(cudaFreeAsync(ctx->sweep_status, ctx->stream));
        ctx->sweep_status = nullptr;
        ctx->sweep_status_words = 0;
        size_t cap = words + words / 4;
        CUDA_CHECK(cudaMallocAsync((void**)&ctx->sweep_status, cap * sizeof(u64), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(ctx->sweep_status, 0, cap * sizeof(u64), ctx->stream));
        ctx->sweep_status_words = cap;
        ctx->sweep_epoch = 0;
    }
}

static u32 next_epoch(ottocov_ctx* ctx) {
    if (ctx->sweep_epoch >= 63) {
        CUDA_CHECK(cudaMemsetAsync(ctx->sweep_status, 0, ctx->sweep_status_words * sizeof(u64), ctx->stream));
        ctx->sweep_epoch = 0;
    }
    return ++ctx->sweep_epoch;
}

int radix_sort_pairs(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n,
                     const BitField* fields, int n_fields) {
    PassList pl;
    pl.n = 0;
    for (int f = 0; f < n_fields; ++f) {
        const int w = fields[f].hi - fields[f].lo;
        if (w <= 0) continue;
        const int np = (w + RS_MAX_BITS - 1) / RS_MAX_BITS;
        int lo = fields[f].lo;
        for (int p = 0; p < np; ++p) {
            const int b = w / np + (p < w % np ? 1 : 0);
            if (pl.n >= RS_MAX_PASSES) COV_THROW(OTTOCOV_ERR_ARG, "too many radix passes");
            pl.shift[pl.n] = lo;
            pl.bits[pl.n] = b;
            ++pl.n;
            lo += b;
        }
    }
    if (n <= 1 || pl.n == 0) return 0;
    const bool has_vals = vals != nullptr;

    DevBuf<u64> ghist(ctx, (size_t)pl.n * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ghist.p, 0, (size_t)pl.n * RS_RADIX * sizeof(u64), ctx->stream));
    {
        int grid = (int)imin64(ceil_div64(n, 256 * 8), (int64_t)ctx->num_sms * 8);
        COV_LAUNCH(ctx, OTTOCOV_K_HIST, 8.0 * n, rs_histogram_kernel, grid, 256,
                   (size_t)pl.n * RS_RADIX * sizeof(u32), keys, n, pl, ghist.p);
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_scan_hist_kernel, pl.n, RS_RADIX, 0, ghist.p);
    }

    const int64_t n_tiles = ceil_div64(n, RS_TILE);
    ensure_sweep_state(ctx, (size_t)n_tiles * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ctx->sweep_ticket, 0, RS_MAX_PASSES * sizeof(u32), ctx->stream));

    const size_t smem = rs_smem_bytes(has_vals);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false)));
        attr_set = true;
    }
    const double pass_bytes = (has_vals ? 24.0 : 16.0) * (double)n;
    for (int p = 0; p < pl.n; ++p) {
        const u32 epoch = next_epoch(ctx);
        if (has_vals) {
            COV_LAUNCH(ctx, OTTOCOV_K_SORT_PASS, pass_bytes, rs_onesweep_kernel<true>, (unsigned)n_tiles,
                       RS_THREADS, smem, keys, alt, vals, valt, n, pl.shift[p], pl.bits[p],
                       ghist.p + (size_t)p * RS_RADIX, ctx->sweep_status, ctx->sweep_ticket + p, epoch);
        } else {
            COV_LAUNCH(ctx, OTTOCOV_K_SORT_PASS, pass_bytes, rs_onesweep_kernel<false>, (unsigned)n_tiles,
                       RS_THREADS, smem, keys, alt, (const u32*)nullptr, (u32*)nullptr, n, pl.shift[p],
                       pl.bits[p], ghist.p + (size_t)p * RS_RADIX, ctx->sweep_status,
                       ctx->sweep_ticket + p, epoch);
        }
        u64* tk = keys; keys = alt; alt = tk;
        if (has_vals) { u32* tv = vals; vals = valt; valt = tv; }
    }
    return pl.n;
}
