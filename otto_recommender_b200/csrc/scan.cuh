// scan.cuh -- generic "reduce / scan block sums / apply" exclusive-scan framework.
//
// Every compaction in the engine (duplicate removal, split by type, window records, run heads,
// threshold filter, segment heads, destination partition) is the same shape: each element i has
// up to NC small counters; element i needs the exclusive prefix of every counter.  A functor F
// supplies
//     static constexpr int NC;                    number of counters (1..3)
//     __device__ u64  value(int64_t i) const;      the counters of element i, packed
//     __device__ void apply(int64_t i, u64 packed_value, const u64* prefix /*[NC]*/) const;
// Packing inside one tile: NC==1 -> the full 64 bits; NC==2 -> counter 0 in bits [0,52),
// counter 1 in [52,64) (must stay < 4096 per tile: it is a 0/1 flag everywhere it is used);
// NC==3 -> 21 bits each (0/1 flags).  Across tiles the counters are carried unpacked as u64.
//
// Element order inside a tile: warp w owns elements [w*32*ITEMS, (w+1)*32*ITEMS); in round r its
// lane l handles element w*32*ITEMS + r*32 + l, so every warp access is a coalesced 32-wide row.
// No inter-CTA waiting anywhere (three plain launches), so it cannot hang.
#pragma once
#include "internal.cuh"

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_MAX_NC = 3;

template <int NC>
__device__ __forceinline__ void scan_unpack(u64 p, u64* c) {
    if (NC == 1) {
        c[0] = p;
    } else if (NC == 2) {
        c[0] = p & ((1ull << 52) - 1);
        c[1] = p >> 52;
    } else {
        c[0] = p & 0x1FFFFF;
        c[1] = (p >> 21) & 0x1FFFFF;
        c[2] = (p >> 42) & 0x1FFFFF;
    }
}

// pass 1: per-tile totals, unpacked: block_sums[c * n_tiles + tile]
template <class F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(F f, int64_t n, int64_t n_tiles,
                                                                    u64* __restrict__ block_sums) {
    __shared__ u64 s_part[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)warp * 32 * SCAN_ITEMS + lane;
    u64 sum = 0;
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        int64_t i = base + r * 32;
        if (i < n) sum += f.value(i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) s_part[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_part[w];
        u64 c[SCAN_MAX_NC];
        scan_unpack<F::NC>(t, c);
        for (int k = 0; k < F::NC; ++k) block_sums[(int64_t)k * n_tiles + blockIdx.x] = c[k];
    }
}

// pass 2: one CTA per counter turns its row of tile totals into exclusive prefixes; totals[c] = sum
static __global__ void __launch_bounds__(1024) scan_block_sums_kernel(u64* __restrict__ block_sums,
                                                               int64_t n_tiles,
                                                               u64* __restrict__ totals) {
    __shared__ u64 s_warp[1024 / 32 + 1];
    u64* row = block_sums + (int64_t)blockIdx.x * n_tiles;
    u64 carry = 0;
    for (int64_t b = 0; b < n_tiles; b += 1024) {
        int64_t i = b + threadIdx.x;
        u64 v = (i < n_tiles) ? row[i] : 0;
        u64 tot;
        u64 ex = block_exclusive_scan<u64, 1024>(v, s_warp, &tot);
        if (i < n_tiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = carry;
}

// pass 3: recompute the values, scan inside the tile, add the tile prefix, hand to F::apply
template <class F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(F f, int64_t n, int64_t n_tiles,
                                                                   const u64* __restrict__ block_prefix) {
    __shared__ u64 s_warp_tot[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)warp * 32 * SCAN_ITEMS + lane;
    u64 v[SCAN_ITEMS], ex[SCAN_ITEMS];
    u64 carry = 0;      // packed running total of this warp's earlier rounds
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        int64_t i = base + r * 32;
        v[r] = (i < n) ? f.value(i) : 0;
        u64 inc = v[r];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        ex[r] = carry + inc - v[r];
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_warp_tot[warp] = carry;
    __syncthreads();
    u64 warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += s_warp_tot[w];
    u64 tile_pref[SCAN_MAX_NC];
    for (int k = 0; k < F::NC; ++k) tile_pref[k] = block_prefix[(int64_t)k * n_tiles + blockIdx.x];
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        int64_t i = base + r * 32;
        if (i < n) {
            u64 c[SCAN_MAX_NC], pre[SCAN_MAX_NC];
            scan_unpack<F::NC>(warp_off + ex[r], c);
            for (int k = 0; k < F::NC; ++k) pre[k] = tile_pref[k] + c[k];
            f.apply(i, v[r], pre);
        }
    }
}

// Host driver.  totals_host (may be nullptr) receives the NC grand totals and forces a stream sync.
template <class F>
static void scan_apply(ottocov_ctx* ctx, int family, const F& f, int64_t n, u64* totals_host,
                       double algo_bytes) {
    if (totals_host)
        for (int k = 0; k < F::NC; ++k) totals_host[k] = 0;
    if (n <= 0) return;
    const int64_t n_tiles = ceil_div64(n, SCAN_TILE);
    DevBuf<u64> sums(ctx, (size_t)n_tiles * F::NC + SCAN_MAX_NC);
    u64* totals_dev = sums.p + (size_t)n_tiles * F::NC;
    COV_LAUNCH(ctx, family, algo_bytes * 0.5, (scan_reduce_kernel<F>), (unsigned)n_tiles,
               SCAN_THREADS, 0, f, n, n_tiles, sums.p);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, n_tiles * 16.0 * F::NC, scan_block_sums_kernel, F::NC, 1024, 0,
               sums.p, n_tiles, totals_dev);
    COV_LAUNCH(ctx, family, algo_bytes * 0.5, (scan_apply_kernel<F>), (unsigned)n_tiles,
               SCAN_THREADS, 0, f, n, n_tiles, sums.p);
    if (totals_host) {
        CUDA_CHECK(cudaMemcpyAsync(totals_host, totals_dev, sizeof(u64) * F::NC,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
}
