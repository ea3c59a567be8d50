// scan.cuh -- single-pass ("chained scan", decoupled look-back) exclusive-scan framework.
//
// Every compaction in the engine (duplicate removal + split by type, window records, threshold
// filter, segment heads, destination partition) has the same shape: element i carries up to NC small
// counters and needs the exclusive prefix of each.  A functor F supplies
//     static constexpr int NC;                    number of counters (1..3)
//     __device__ u64  value(int64_t i) const;      the counters of element i, packed
//     __device__ void apply(int64_t i, u64 packed_value, const u64* prefix /*[NC]*/) const;
// Packing inside one tile: NC==1 -> the full 64 bits; NC==2 -> counter 0 in bits [0,50), counter 1
// in [50,64) (a 0/1 flag everywhere it is used: <= 4096 per tile); NC==3 -> 21 bits each (0/1 flags).  Across tiles the
// counters travel unpacked, one 64-bit status word per (tile, counter):
//     flag(2) | epoch(6) | value(56)        flag 1 = tile aggregate, 2 = inclusive prefix
// The epoch changes with every launch, so the status array is never cleared between launches.
//
// One kernel, one read of the input: a CTA takes the next tile (atomic ticket => it only ever waits
// for tiles that have already started), scans it, publishes its aggregate, warp 0 resolves the
// exclusive prefix over earlier tiles by a 32-wide look-back, then every element is handed to
// F::apply with its global prefix.
//
// 512 threads x 8 rows = 4096 elements per tile.  Element order inside a tile: warp w owns elements [w*32*ITEMS, (w+1)*32*ITEMS); in round r its lane
// l handles element w*32*ITEMS + r*32 + l, so every warp access is a coalesced 32-wide row.
#pragma once
#include "internal.cuh"

#ifndef OTTOCOV_SCAN_ITEMS
#define OTTOCOV_SCAN_ITEMS 8
#endif
#ifndef OTTOCOV_SCAN_MINB
#define OTTOCOV_SCAN_MINB 3      // 3 CTAs / SM (40 registers): measured 0.6 ms faster per step than 2 CTAs at 64
#endif
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = OTTOCOV_SCAN_ITEMS;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_MAX_NC = 3;
constexpr int SCAN_NC2_SHIFT = 50;             // NC == 2: counter 1 lives in bits [50, 64)
static_assert(SCAN_TILE < (1 << (64 - SCAN_NC2_SHIFT)) && SCAN_TILE < (1 << 21), "per-tile flag counters would overflow");

constexpr u64 SC_VALUE_MASK = (1ull << 56) - 1;
constexpr u64 SC_FLAG_AGG = 1ull << 62;
constexpr u64 SC_FLAG_PREFIX = 2ull << 62;

// Optional per-thread side accumulation while the elements are read (e.g. the loader's range / order statistics, which
// would otherwise cost a pass of their own): a functor with a member type `Acc` supplies
//     __device__ void acc_init(Acc&) const;   __device__ u64 value(int64_t i, Acc&) const;   __device__ void acc_flush(Acc&) const;
// acc_flush is called by every thread right after its last value() (warp-reduce + a few atomics).
template <class F, class = void> struct scan_has_acc { static constexpr bool value = false; };
template <class F> struct scan_has_acc<F, decltype((void)sizeof(typename F::Acc))> { static constexpr bool value = true; };

template <int NC>
__device__ __forceinline__ void scan_unpack(u64 p, u64* c) {
    if (NC == 1) {
        c[0] = p;
    } else if (NC == 2) {
        c[0] = p & ((1ull << SCAN_NC2_SHIFT) - 1);
        c[1] = p >> SCAN_NC2_SHIFT;
    } else {
        c[0] = p & 0x1FFFFF;
        c[1] = (p >> 21) & 0x1FFFFF;
        c[2] = (p >> 42) & 0x1FFFFF;
    }
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Called by all 32 lanes of ONE warp.  `word` points at status[tile] for one counter, consecutive tiles
// are `stride` words apart.  Publishes `aggregate`, returns the exclusive prefix over tiles < tile and
// publishes the inclusive prefix.
__device__ __forceinline__ u64 warp_lookback_sum(u64* word, int64_t stride, int64_t tile, u64 aggregate,
                                                 u32 epoch) {
    const int lane = threadIdx.x & 31;
    const u64 tag = (u64)epoch << 56;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(word, SC_FLAG_PREFIX | tag | aggregate);
        return 0;
    }
    if (lane == 0) st_volatile_u64(word, SC_FLAG_AGG | tag | aggregate);
    u64 excl = 0;
    int64_t base = tile - 1;                       // lane l looks at tile base - l
    while (true) {
        const int64_t idx = base - lane;
        const u64 v = (idx >= 0) ? ld_volatile_u64(word - (tile - idx) * stride) : (SC_FLAG_PREFIX | tag);
        const bool ready = ((u32)((v >> 56) & 0x3F) == epoch) && ((v >> 62) != 0);
        const u32 rmask = __ballot_sync(0xffffffffu, ready);
        const u32 pmask = __ballot_sync(0xffffffffu, ready && (v >> 62) == 2);
        const int first_nr = (~rmask) ? (__ffs(~rmask) - 1) : 32;    // lanes [0, first_nr) are ready
        const int first_p = pmask ? (__ffs(pmask) - 1) : 32;
        if (first_p < first_nr) {                  // a prefix inside the ready run: done
            excl += warp_sum_u64(lane <= first_p ? (v & SC_VALUE_MASK) : 0ull);
            break;
        }
        excl += warp_sum_u64(lane < first_nr ? (v & SC_VALUE_MASK) : 0ull);
        base -= first_nr;
        if (first_nr == 0) __nanosleep(40);
    }
    if (lane == 0) st_volatile_u64(word, SC_FLAG_PREFIX | tag | (excl + aggregate));
    return excl;
}

// CHAINED: one logical scan cut into several launches (carry_in = inclusive totals of the earlier launches); a
// separate instantiation so that the single-launch kernels compile exactly as before.
template <class F, bool CHAINED>
__global__ void __launch_bounds__(SCAN_THREADS, OTTOCOV_SCAN_MINB) scan_onepass_kernel(F f, int64_t n, int64_t n_tiles, u64* status,
                                                                     u32* ticket, u32 epoch,
                                                                     u64* __restrict__ totals,
                                                                     const u64* __restrict__ carry_in) {
    __shared__ u64 s_warp_tot[SCAN_THREADS / 32];
    __shared__ u64 s_tile_pref[SCAN_MAX_NC];
    __shared__ u32 s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SCAN_TILE + (int64_t)warp * 32 * SCAN_ITEMS + lane;
    u64 v[SCAN_ITEMS], ex[SCAN_ITEMS];
    u64 carry = 0;      // packed running total of this warp's earlier rounds
    if constexpr (scan_has_acc<F>::value) {
        typename F::Acc acc;
        f.acc_init(acc);
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r) {
            const int64_t i = base + r * 32;
            v[r] = (i < n) ? f.value(i, acc) : 0;
        }
        f.acc_flush(acc);
    } else {
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r) {
            const int64_t i = base + r * 32;
            v[r] = (i < n) ? f.value(i) : 0;
        }
    }
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        u64 inc = v[r];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        ex[r] = carry + inc - v[r];
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_warp_tot[warp] = carry;
    __syncthreads();
    if (warp < F::NC) {                          // warp k resolves counter k: the NC look-backs run side by side
        u64 t = (lane < SCAN_THREADS / 32) ? s_warp_tot[lane] : 0ull;
        t = warp_sum_u64(t);
        u64 c[SCAN_MAX_NC];
        scan_unpack<F::NC>(t, c);
        u64 mine = c[0];
#pragma unroll
        for (int k = 1; k < F::NC; ++k) mine = (warp == k) ? c[k] : mine;
        u64 pre = warp_lookback_sum(status + tile * F::NC + warp, F::NC, tile, mine, epoch);
        if (CHAINED) pre += carry_in[warp];
        if (lane == 0) {
            s_tile_pref[warp] = pre;
            if (tile == n_tiles - 1) totals[warp] = pre + mine;
        }
    }
    __syncthreads();
    u64 warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += s_warp_tot[w];
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        const int64_t i = base + r * 32;
        if (i < n) {
            u64 c[SCAN_MAX_NC], pre[SCAN_MAX_NC];
            scan_unpack<F::NC>(warp_off + ex[r], c);
#pragma unroll
            for (int k = 0; k < F::NC; ++k) pre[k] = s_tile_pref[k] + c[k];
            f.apply(i, v[r], pre);
        }
    }
}

// grow-only status words + ticket + totals; returns the epoch to launch with
void scan_state_prepare(ottocov_ctx* ctx, size_t status_words, u32* epoch_out);     // reduce.cu

// Host driver.  totals_host (may be nullptr) receives the NC grand totals and forces a stream sync.
// carry_in / totals_out (device, [NC]): chain one logical scan over several launches -- a launch adds carry_in to
// every prefix and writes its inclusive totals (carry included) to totals_out (default: ctx->scan_totals).
template <class F>
static void scan_apply(ottocov_ctx* ctx, int family, const F& f, int64_t n, u64* totals_host,
                       double algo_bytes, const u64* carry_in = nullptr, u64* totals_out = nullptr) {
    if (totals_host)
        for (int k = 0; k < F::NC; ++k) totals_host[k] = 0;
    if (n <= 0) return;
    const int64_t n_tiles = ceil_div64(n, SCAN_TILE);
    u32 epoch;
    scan_state_prepare(ctx, (size_t)n_tiles * F::NC, &epoch);
    u64* tout = totals_out ? totals_out : ctx->scan_totals;
    if (carry_in)
        COV_LAUNCH(ctx, family, algo_bytes, (scan_onepass_kernel<F, true>), (unsigned)n_tiles, SCAN_THREADS, 0, f, n, n_tiles,
                   ctx->scan_status, ctx->scan_ticket, epoch, tout, carry_in);
    else
        COV_LAUNCH(ctx, family, algo_bytes, (scan_onepass_kernel<F, false>), (unsigned)n_tiles, SCAN_THREADS, 0, f, n, n_tiles,
                   ctx->scan_status, ctx->scan_ticket, epoch, tout, (const u64*)nullptr);
    if (totals_host) {
        cov_readback(ctx, totals_host, tout, sizeof(u64) * F::NC);
    }
}
