// expand.cu -- pair expansion + reduce-by-key for ONE co-event kind.
//
// Replaces self_merge (model/count_co_events.py:17-38: join on session, drop the event joined with
// itself, |ts_next - ts| <= 24 h) and one iteration of count_co_events' loop (:64-72: type filter,
// |dt| <= W, groupby(aid, aid_next).count()).  The reference materialises all n^2 joined rows per
// session and filters them; here nothing is rejected:
//
//   window kernel   one thread per source event (type == type_this).  Target events of one type are
//                   sorted by (session, ts), so the in-window targets are ONE index range [lo, hi);
//                   found by galloping + binary search outwards from the source's own insertion rank
//                   (xrank).  Emits lo and cnt = hi - lo (minus 1 when source and target type agree:
//                   the event itself lies in its own window and is the only excluded row, :23-27).
//   scan            exclusive scan of cnt -> exact output offsets; zero-count sources are compacted
//                   away so every record owns >= 1 output.
//   tile search     one thread per 2048-output tile finds the first record of the tile.
//   expand kernel   output-balanced: each CTA writes exactly one tile of packed u64 keys
//                   (aid << 32 | aid_next), whatever the session lengths are (a 498-event session and
//                   a 2-event session cost the same per emitted pair).  Threads write 2 adjacent keys
//                   with one 128-bit store; a warp store covers 512 contiguous bytes.
//   sort + RLE      radix_sort.cu, reduce.cu.
// If the pair count exceeds the budget, the output space is cut into chunks (any cut point works,
// tiles are addressed by output offset) and the partial tables are merged at the end.
#include "internal.cuh"
#include "scan.cuh"
#include <atomic>
#include <thread>
#include <memory>

#ifndef OTTOCOV_EXS_MINB
#define OTTOCOV_EXS_MINB 6        // 40 registers without spills; measured 5.84 (4 CTAs/SM) -> 5.40 (5) -> 5.27 ms (6) on 742 M keys
#endif
constexpr int EX_THREADS = 256;
constexpr int EX_TILE = 2048;
constexpr int EX_PER = EX_TILE / EX_THREADS;          // consecutive outputs per thread

// first index in [0, start] with keys[idx] >= bound, given keys[i] >= bound for all i >= start
__device__ __forceinline__ u32 lower_bound_back(const u64* __restrict__ keys, u32 start, u64 bound) {
    u32 hi = start, lo, step = 1;
    while (true) {
        if (hi == 0) return 0;
        const u32 probe = (hi >= step) ? hi - step : 0;
        if (keys[probe] >= bound) { hi = probe; step <<= 1; }
        else { lo = probe; break; }
    }
    while (hi - lo > 1) {               // keys[lo] < bound <= keys[hi]
        const u32 mid = lo + ((hi - lo) >> 1);
        if (keys[mid] >= bound) hi = mid; else lo = mid;
    }
    return hi;
}

// first index in [start, n] with keys[idx] > bound, given keys[i] <= bound for all i < start
__device__ __forceinline__ u32 upper_bound_fwd(const u64* __restrict__ keys, u32 n, u32 start, u64 bound) {
    u32 lo = start, hi, step = 1;
    while (true) {
        if (lo >= n) return n;
        const u32 probe = (n - lo > step) ? lo + step - 1 : n - 1;
        if (keys[probe] <= bound) { lo = probe + 1; step <<= 1; }
        else { hi = probe; break; }
    }
    while (lo < hi) {                   // everything < lo is <= bound; keys[hi] > bound
        const u32 mid = lo + ((hi - lo) >> 1);
        if (keys[mid] <= bound) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// per source event: target range start and count
__global__ void __launch_bounds__(256) window_kernel(const u64* __restrict__ src_key,
                                                     const u32* __restrict__ src_xrank,   // nullptr => same array
                                                     int64_t n_src, const u64* __restrict__ tgt_key,
                                                     u32 n_tgt, u32 window, u32* __restrict__ lo_out,
                                                     u32* __restrict__ cnt_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_src) return;
    const u64 k = src_key[j];
    const u32 t = (u32)k;
    const u64 s = k & 0xFFFFFFFF00000000ull;
    const u64 lower = s | (u64)(t > window ? t - window : 0u);
    const u64 upper = s | (u64)(t > 0xFFFFFFFFu - window ? 0xFFFFFFFFu : t + window);
    const bool self = (src_xrank == nullptr);
    const u32 start = self ? (u32)j : src_xrank[j];
    const u32 lo = lower_bound_back(tgt_key, start, lower);
    const u32 hi = upper_bound_fwd(tgt_key, n_tgt, start, upper);
    lo_out[j] = lo;
    cnt_out[j] = hi - lo - (self ? 1u : 0u);
}

// Symmetric kinds (source type == target type): every unordered event pair {i, j} is emitted ONCE, from
// its earlier event, as the canonical key (min aid, max aid); count(a, b) == count(b, a) is restored by
// mirroring the reduced table.  Targets are the events after j inside the window.
__global__ void __launch_bounds__(256) window_fwd_kernel(const u64* __restrict__ key, int64_t n, u32 window,
                                                         u32* __restrict__ lo_out, u32* __restrict__ cnt_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u64 k = key[j];
    const u32 t = (u32)k;
    const u64 upper = (k & 0xFFFFFFFF00000000ull) | (u64)(t > 0xFFFFFFFFu - window ? 0xFFFFFFFFu : t + window);
    const u32 hi = upper_bound_fwd(key, (u32)n, (u32)j, upper);
    if (lo_out) lo_out[j] = (u32)j + 1u;          // the canonical expansion knows it: targets start right after the source
    cnt_out[j] = hi - (u32)j - 1u;
}

// General time range dt_lo <= ts_tgt - ts_src <= dt_hi (a config whose MIN/MAX_TIME_TO_NEXT pre-filter is not the
// reference's symmetric +-24 h, count_co_events.py:33-36): plain binary searches over the whole target array.
// self_in: source and target type agree and 0 lies in the range, so the event itself is inside and is excluded.
__global__ void __launch_bounds__(256) window_range_kernel(const u64* __restrict__ src_key, int64_t n_src,
                                                           const u64* __restrict__ tgt_key, u32 n_tgt, int64_t dt_lo,
                                                           int64_t dt_hi, int self_in, u32* __restrict__ lo_out,
                                                           u32* __restrict__ cnt_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_src) return;
    const u64 k = src_key[j];
    const int64_t t = (int64_t)(u32)k;
    const u64 s = k & 0xFFFFFFFF00000000ull;
    const int64_t a = t + dt_lo, b = t + dt_hi;
    u32 lo = 0, cnt = 0;
    if (b >= 0 && a <= 0xFFFFFFFFll && a <= b) {
        const u64 lower = s | (u64)(a < 0 ? 0 : a);
        const u64 upper = s | (u64)(b > 0xFFFFFFFFll ? 0xFFFFFFFFll : b);
        u32 l = 0, h = n_tgt;                       // first index with key >= lower
        while (l < h) { const u32 m = l + ((h - l) >> 1); if (tgt_key[m] >= lower) h = m; else l = m + 1; }
        lo = l;
        h = n_tgt;                                  // first index with key > upper
        while (l < h) { const u32 m = l + ((h - l) >> 1); if (tgt_key[m] > upper) h = m; else l = m + 1; }
        cnt = l - lo - (self_in ? 1u : 0u);
    }
    lo_out[j] = lo;
    cnt_out[j] = cnt;
}

// compaction of the non-empty sources into records (src, lo, output offset)
struct WindowRecords {
    static constexpr int NC = 2;
    const u32* lo;
    const u32* cnt;
    u32* rec_src;
    u32* rec_lo;
    u64* rec_off;
    __device__ u64 value(int64_t j) const {
        const u64 c = cnt[j];
        return c | ((u64)(c != 0) << SCAN_NC2_SHIFT);
    }
    __device__ void apply(int64_t j, u64 v, const u64* pre) const {
        if (!v) return;
        const u64 r = pre[1];
        rec_src[r] = (u32)j;
        if (rec_lo) rec_lo[r] = lo[j];            // symmetric kinds: lo == src + 1, neither stored nor re-read
        rec_off[r] = pre[0];
    }
};

// tile t of a launch starts at output offset out_begin + t * EX_TILE: last record with off <= that
__global__ void __launch_bounds__(256) tile_search_kernel(const u64* __restrict__ rec_off, int64_t n_rec,
                                                          u64 out_begin, u64 out_end, int64_t n_tiles,
                                                          u32* __restrict__ tile_rec) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    u64 o = out_begin + (u64)t * EX_TILE;
    if (o > out_end - 1) o = out_end - 1;        // entry n_tiles = record of the last output
    int64_t lo = 0, hi = n_rec;                  // first record with off > o
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (rec_off[mid] <= o) lo = mid + 1; else hi = mid;
    }
    tile_rec[t] = (u32)(lo - 1);
}

// MIX: the key is written through the bijective mix of internal.cuh (bucketed hash reduce); never together
// with a destination stamp (the multi-GPU path mixes on the receiving side).
template <bool CANON, bool MIX>
__device__ __forceinline__ u64 make_pair_key(u32 a, u32 b, u32 n_dest, const KeyMix& mix) {
    u32 x = a, y = b;
    if (CANON) { x = a < b ? a : b; y = a < b ? b : a; }
    if (MIX) return key_mix_fwd(mix, x, y);
    u64 k = ((u64)x << 32) | (u64)y;
    if (n_dest > 1) k |= (u64)hash_dest(x, n_dest) << 56;      // destination rank of the row (x, .)
    return k;
}

// ---- the two halves of a tile that both expansion kernels share ------------------------------------------------
constexpr int EX_PITCH = EX_TILE + 2;

// Stage the records that own outputs of tile `tile` in shared memory (s_off: first output of the record relative to
// the tile, s_lo: target index of that output, s_aid: source aid, s_src: source index).  Returns the record count;
// *n_out_p = outputs of the tile.  Ends with a barrier.
template <bool SELF>
__device__ __forceinline__ u32 stage_tile_records(const u32* __restrict__ rec_src, const u32* __restrict__ rec_lo,
                                                  const u64* __restrict__ rec_off, const u32* __restrict__ tile_rec,
                                                  const u32* __restrict__ aid_src, u64 out_begin, u64 out_end, int64_t tile,
                                                  u32* s_off, u32* s_lo, u32* s_aid, u32* s_src, u32* n_out_p) {
    const u64 o0 = out_begin + (u64)tile * EX_TILE;
    const u32 n_out = (u32)min((u64)EX_TILE, out_end - o0);
    const u32 r0 = tile_rec[tile];
    const u32 r1 = tile_rec[tile + 1];                // record of the tile's last output (clamped)
    // records that own outputs of this tile: r0 .. r_last, r_last = record of output o0 + n_out - 1
    u32 r_last = r1;
    if ((u64)tile * EX_TILE + EX_TILE + out_begin < out_end) {
        // tile_rec[t+1] is the record of the NEXT tile's first output; it owns outputs of this
        // tile only if it starts before that output
        if (rec_off[r1] >= o0 + n_out) r_last = r1 - 1;
    }
    const u32 n_rec = r_last - r0 + 1;                // <= EX_TILE (offsets strictly increase)
    for (u32 j = threadIdx.x; j < n_rec; j += EX_THREADS) {
        const u32 r = r0 + j;
        const u64 off = rec_off[r];
        const u32 src = rec_src[r];
        u32 lo = rec_lo ? rec_lo[r] : src + 1u;
        u32 rel;
        if (off <= o0) { rel = 0; lo += (u32)(o0 - off); }   // only j == 0: skip outputs of earlier tiles
        else rel = (u32)(off - o0);
        s_off[j] = rel;
        s_lo[j] = lo;
        s_aid[j] = aid_src[src];
        if (SELF) s_src[j] = src;
    }
    __syncthreads();
    *n_out_p = n_out;
    return n_rec;
}

// Thread t produces the EX_PER consecutive outputs [EX_PER t, EX_PER t + EX_PER) of the tile: one binary search
// for the first, a linear walk over the record offsets for the rest.  Only outputs k < n_out are defined.
// MODE: how a record (its owner event) and the events of its range make a key
//   EXM_CROSS  owner = source event, range = target events of another type:  key (owner aid, range aid)
//   EXM_SELF   the same with source type == target type: the owner itself lies inside its range and is skipped
//   EXM_CANON  symmetric kind, each unordered event pair once:                key (min aid, max aid)
//   EXM_SWAP   owner = TARGET event, range = source events (taken when the target type is much rarer than the source
//              type -- carts and orders against clicks -- so the window pass walks the few, not the many):
//                                                                             key (range aid, owner aid)
enum { EXM_CROSS = 0, EXM_SELF = 1, EXM_CANON = 2, EXM_SWAP = 3 };

template <int MODE, bool MIX>
__device__ __forceinline__ void make_tile_keys(const u32* s_off, const u32* s_lo, const u32* s_aid, const u32* s_src,
                                               u32 n_rec, u32 n_out, const u32* __restrict__ aid_tgt, u32 n_dest,
                                               const KeyMix& mix, u64 (&key)[EX_PER]) {
    constexpr bool SELF = MODE == EXM_SELF, CANON = MODE == EXM_CANON, SWAP = MODE == EXM_SWAP;
    const u32 k0 = (u32)EX_PER * threadIdx.x;
    if (k0 >= n_out) return;
    u32 lo = 0, hi = n_rec;                       // last record with s_off <= k0
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (s_off[mid] <= k0) lo = mid; else hi = mid;
    }
    u32 j = lo;
    u32 a = s_aid[j];
    u32 tb = s_lo[j] - s_off[j];                  // target index = tb + k (mod 2^32)
    u32 src = SELF ? s_src[j] : 0u;
    u32 next_off = (j + 1 < n_rec) ? s_off[j + 1] : 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < EX_PER; ++q) {
        const u32 k = k0 + q;
        if (k < n_out) {
            if (k >= next_off) {                  // offsets strictly increase: at most one step per output
                ++j;
                a = s_aid[j];
                tb = s_lo[j] - s_off[j];
                if (SELF) src = s_src[j];
                next_off = (j + 1 < n_rec) ? s_off[j + 1] : 0xFFFFFFFFu;
            }
            u32 tgt = tb + k;
            if (SELF) tgt += (tgt >= src);
            const u32 other = aid_tgt[tgt];
            key[q] = SWAP ? make_pair_key<false, MIX>(other, a, n_dest, mix) : make_pair_key<CANON, MIX>(a, other, n_dest, mix);
        }
    }
}

// Persistent: CTA b writes tiles b, b + gridDim.x, ...  Per tile the records that own its outputs are staged in
// shared memory; thread t then produces the EX_PER consecutive outputs [EX_PER t, EX_PER t + EX_PER).  The keys go
// through a swizzled shared-memory transpose and leave as 128-bit stores, 512 contiguous bytes per warp instruction.
// ghist != nullptr: the digit histograms of the distribution passes that will sort these keys (pl) are
// accumulated here, in shared memory while the keys are still in registers, and flushed once per CTA -- the
// sort's own histogram kernel (one more read of every key) is not needed.
template <int MODE, bool MIX>
__global__ void __launch_bounds__(EX_THREADS)
expand_kernel(const u32* __restrict__ rec_src, const u32* __restrict__ rec_lo,
              const u64* __restrict__ rec_off, const u32* __restrict__ tile_rec,
              const u32* __restrict__ aid_src, const u32* __restrict__ aid_tgt, u64 out_begin,
              u64 out_end, u64* __restrict__ dst, u32 n_dest, KeyMix mix, int64_t n_tiles, PassList pl,
              u64* __restrict__ ghist) {
    constexpr bool SELF = MODE == EXM_SELF;
    __shared__ __align__(16) u32 s_buf[3 * EX_PITCH];
    __shared__ u32 s_src[SELF ? EX_TILE + 1 : 1];
    extern __shared__ u32 s_hist[];                   // [pl.n][RS_RADIX] when ghist
    u32* s_off = s_buf;
    u32* s_lo = s_buf + EX_PITCH;
    u32* s_aid = s_buf + 2 * EX_PITCH;
    u64* s_out = reinterpret_cast<u64*>(s_buf);       // [EX_TILE]: re-uses s_off / s_lo once the keys sit in registers
    const bool hist = ghist != nullptr;
    if (hist)
        for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += EX_THREADS) s_hist[j] = 0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u32 n_out;
        const u32 n_rec = stage_tile_records<SELF>(rec_src, rec_lo, rec_off, tile_rec, aid_src, out_begin, out_end, tile,
                                                   s_off, s_lo, s_aid, s_src, &n_out);
        u64 key[EX_PER];
        const u32 k0 = (u32)EX_PER * threadIdx.x;
        make_tile_keys<MODE, MIX>(s_off, s_lo, s_aid, s_src, n_rec, n_out, aid_tgt, n_dest, mix, key);
        if (hist && k0 < n_out) {
            for (int p = 0; p < pl.n; ++p) {              // pass parameters are read once per pass, not per key
                const int sh = pl.shift[p];
                const u32 msk = (1u << pl.bits[p]) - 1u;
                u32* hrow = s_hist + p * RS_RADIX;
#pragma unroll
                for (int q = 0; q < EX_PER; ++q)
                    if (k0 + q < n_out) atomicAdd(&hrow[(u32)(key[q] >> sh) & msk], 1u);
            }
        }
        __syncthreads();                                  // every thread is done with the staged records
        if (k0 < n_out) {
#pragma unroll
            for (int q = 0; q < EX_PER; ++q)              // output o sits at o ^ ((o >> 3) & 7): conflict-free read-back
                if (k0 + q < n_out) s_out[k0 + (u32)(q ^ (int)(threadIdx.x & 7))] = key[q];
        }
        __syncthreads();
        u64* tile_dst = dst + (size_t)tile * EX_TILE;
#pragma unroll
        for (int it = 0; it < EX_TILE / (2 * EX_THREADS); ++it) {
            const u32 k = (u32)it * (2 * EX_THREADS) + 2 * threadIdx.x;
            if (k >= n_out) continue;
            const u32 m = (k >> 3) & 7u;
            const u64 key0 = s_out[k ^ m];
            if (k + 1 < n_out) {
                const u64 key1 = s_out[(k + 1) ^ m];
                if (vec_ok) {
                    ulonglong2 v; v.x = key0; v.y = key1;
                    __stcs(reinterpret_cast<ulonglong2*>(tile_dst + k), v);
                } else {
                    __stcs(tile_dst + k, key0);
                    __stcs(tile_dst + k + 1, key1);
                }
            } else {
                __stcs(tile_dst + k, key0);
            }
        }
        __syncthreads();                                  // s_buf is re-staged by the next tile
    }
    if (hist) {
        __syncthreads();
        for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += EX_THREADS) {
            const u32 c = s_hist[j];
            if (c) atomicAdd(&ghist[j], (u64)c);
        }
    }
}

// ---- expansion fused with the FIRST distribution pass of the bucketed hash reduce -------------------------------
// The first pass of an LSD sort needs no stability (there is no earlier order to preserve), and the digit is made of
// hash bits (uniform whatever the aid skew).  So the tile's keys never go to HBM in expansion order: they are ranked
// inside the tile with one shared-memory fetch-and-add per key (any order inside a digit will do), staged in digit
// order, and each digit's run is appended to that digit's REGION -- room is reserved with one global atomic per
// (tile, digit) on the region's fill counter.  Regions hold their expected share + slack; a reservation that
// does not fit raises HR_FLAG_FUSED_OVERFLOW (the host re-runs the unfused way) and the run is dropped.
// Saves one write and one read of every key (16 of 64 B/key) and the pass's look-back chain.
//   reg_base[d]  byte address of region d (own HBM, or a peer's receive stripe mapped over NVLink: the multi-GPU
//                exchange IS this pass when the digit's high bits are the destination rank)
//   d = ((owner rank of the key) << sub_bits) | ((mixed key >> sh1) & (2^sub_bits - 1));  owner = 0 when n_dest <= 1
//   keep range   only digits in [d_lo, d_hi) are kept (chunked runs take the hash space a digit range at a time, so
//                every chunk holds complete sums and thresholds stay fused)
//   pl / ghist   digit histograms of the REMAINING passes, over the kept keys, per destination rank:
//                ghist[(dest * pl.n + p) * RS_RADIX + digit]
struct ScatterArgs {
    const u64* reg_base;     // [n_digits] device array of byte addresses
    const u64* reg_cap;      // [n_digits] region capacities in keys
    unsigned long long* cursor;   // [n_digits] keys appended to each region so far
    u32* flags;
    int sh1, sub_bits;
    u32 n_digits, d_lo, d_hi;
    u32 n_dest;
};

template <int MODE, bool DIST>
__global__ void __launch_bounds__(EX_THREADS, OTTOCOV_EXS_MINB)
expand_scatter_kernel(const u32* __restrict__ rec_src, const u32* __restrict__ rec_lo,
                      const u64* __restrict__ rec_off, const u32* __restrict__ tile_rec,
                      const u32* __restrict__ aid_src, const u32* __restrict__ aid_tgt, u64 out_begin,
                      u64 out_end, KeyMix mix, int64_t n_tiles, PassList pl, u64* __restrict__ ghist, ScatterArgs sa) {
    constexpr bool SELF = MODE == EXM_SELF;
    __shared__ __align__(16) u32 s_buf[3 * EX_PITCH];
    __shared__ u32 s_src[SELF ? EX_TILE + 1 : 1];
    __shared__ u32 s_cnt[RS_RADIX];                   // keys of the tile per digit, then their first staging slot
    __shared__ u64 s_gptr[RS_RADIX];                  // byte address of the digit's run - 8 * first staging slot; 0 = dropped
    __shared__ u32 s_scan[EX_THREADS / 32 + 1];
    __shared__ unsigned char s_dig[DIST ? EX_TILE : 4];   // DIST: digit of every staged key (the owner is not part of the mixed key)
    extern __shared__ u32 s_hist[];                   // [n_dest][pl.n][RS_RADIX]
    u32* s_off = s_buf;
    u32* s_lo = s_buf + EX_PITCH;
    u32* s_aid = s_buf + 2 * EX_PITCH;
    u64* s_out = reinterpret_cast<u64*>(s_buf);       // [EX_TILE]
    const int tid = threadIdx.x;
    const u32 nd = DIST ? sa.n_dest : 1u;
    const int n_hist = (int)nd * pl.n * RS_RADIX;
    for (int j = tid; j < n_hist; j += EX_THREADS) s_hist[j] = 0;
    for (int j = tid; j < RS_RADIX; j += EX_THREADS) s_cnt[j] = 0;
    const u32 sub_mask = (1u << sa.sub_bits) - 1u;
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u32 n_out;
        const u32 n_rec = stage_tile_records<SELF>(rec_src, rec_lo, rec_off, tile_rec, aid_src, out_begin, out_end, tile,
                                                   s_off, s_lo, s_aid, s_src, &n_out);
        u64 key[EX_PER];
        const u32 k0 = (u32)EX_PER * tid;
        // DIST: plain keys first -- the owner rank is a function of the plain aid -- mixed right below
        if (DIST) make_tile_keys<MODE, false>(s_off, s_lo, s_aid, s_src, n_rec, n_out, aid_tgt, 0u, mix, key);
        else make_tile_keys<MODE, true>(s_off, s_lo, s_aid, s_src, n_rec, n_out, aid_tgt, 0u, mix, key);
        u32 dig[(EX_PER + 3) / 4];                    // digits, 8 bits each
        u32 rnk[EX_PER / 2];                          // ranks inside (tile, digit), 16 bits each
#pragma unroll
        for (int q = 0; q < (EX_PER + 3) / 4; ++q) dig[q] = 0;
#pragma unroll
        for (int q = 0; q < EX_PER / 2; ++q) rnk[q] = 0;
        u32 keep = 0;
        if (k0 < n_out) {
#pragma unroll
            for (int q = 0; q < EX_PER; ++q) {
                if (k0 + q < n_out) {
                    u32 dest = 0;
                    if (DIST) {
                        const u64 plain = key[q];
                        dest = hash_dest((u32)(plain >> 32), sa.n_dest);          // owner of the row (aid, .)
                        key[q] = key_mix_fwd(mix, (u32)(plain >> 32), (u32)plain);
                    }
                    const u32 d = (dest << sa.sub_bits) | ((u32)(key[q] >> sa.sh1) & sub_mask);
                    if (d >= sa.d_lo && d < sa.d_hi) {
                        keep |= 1u << q;
                        const u32 r = atomicAdd(&s_cnt[d], 1u);
                        dig[q >> 2] |= d << (8 * (q & 3));
                        rnk[q >> 1] |= r << (16 * (q & 1));
                    }
                }
            }
            for (int p = 0; p < pl.n; ++p) {              // histograms of the remaining passes, kept keys only
                const int sh = pl.shift[p];
                const u32 msk = (1u << pl.bits[p]) - 1u;
#pragma unroll
                for (int q = 0; q < EX_PER; ++q)
                    if ((keep >> q) & 1u) {
                        const u32 dest = DIST ? ((dig[q >> 2] >> (8 * (q & 3))) & 0xFFu) >> sa.sub_bits : 0u;
                        atomicAdd(&s_hist[(dest * pl.n + p) * RS_RADIX + ((u32)(key[q] >> sh) & msk)], 1u);
                    }
            }
        }
        __syncthreads();                                  // ranks are final; the staged records are no longer needed
        // per digit: first staging slot, and room in the digit's region.  The reservation is a global fetch-and-add
        // (~1 us round trip): it is issued here and its result is first used after the keys have been staged, so the
        // round trip hides behind the barrier and the shared-memory scatter.
        u32 c = 0;
        if (tid < (int)sa.n_digits) c = s_cnt[tid];
        u32 kept_total;
        const u32 dstart = block_exclusive_scan<u32, EX_THREADS>(c, s_scan, &kept_total);
        unsigned long long g = 0;
        if (tid < (int)sa.n_digits) {
            s_cnt[tid] = dstart;
            if (c) g = atomicAdd(&sa.cursor[tid], (unsigned long long)c);
        }
        __syncthreads();
        if (keep) {
#pragma unroll
            for (int q = 0; q < EX_PER; ++q)
                if ((keep >> q) & 1u) {
                    const u32 d = (dig[q >> 2] >> (8 * (q & 3))) & 0xFFu;
                    const u32 r = (rnk[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
                    s_out[s_cnt[d] + r] = key[q];
                    if (DIST) s_dig[s_cnt[d] + r] = (unsigned char)d;
                }
        }
        if (tid < (int)sa.n_digits) {
            u64 gp = 0;
            if (c) {
                if (g + c <= sa.reg_cap[tid]) gp = sa.reg_base[tid] + (g - (u64)dstart) * 8ull;
                else atomicOr(sa.flags, HR_FLAG_FUSED_OVERFLOW);
            }
            s_gptr[tid] = gp;
        }
        __syncthreads();
        // digit-ordered runs out: the digit of slot j is recomputed from the key itself (single GPU) or read back from
        // the byte staged next to it (the owner rank is not part of the mixed key)
#pragma unroll
        for (int it = 0; it < EX_TILE / EX_THREADS; ++it) {
            const u32 j = (u32)it * EX_THREADS + tid;
            if (j < kept_total) {
                const u64 k = s_out[j];
                const u32 d = DIST ? (u32)s_dig[j] : ((u32)(k >> sa.sh1) & sub_mask);
                const u64 gp = s_gptr[d];
                if (gp) __stcs(reinterpret_cast<u64*>(gp + 8ull * j), k);
            }
        }
        __syncthreads();                                  // s_buf is re-staged by the next tile
        if (tid < RS_RADIX) s_cnt[tid] = 0;               // (the barrier at the end of stage_tile_records orders this)
    }
    __syncthreads();
    for (int j = tid; j < n_hist; j += EX_THREADS) {
        const u32 c = s_hist[j];
        if (c) atomicAdd(&ghist[j], (u64)c);
    }
}

struct Segment {
    int tgt_type;            // type of the events in a record's range
    int own_type;            // type of the record owners (the kind's source type, or its target type when swapped)
    bool self;
    bool swap = false;       // EXM_SWAP: owners are the kind's TARGET events
    u64 n_pairs = 0;
    u64 n_rec = 0;
    DevBuf<u32> rec_src, rec_lo;
    DevBuf<u64> rec_off;
};

// Everything the window pass learns about one co-event kind on the loaded events: which sources emit,
// where their targets start, and the exact output offset of every source.  ottocov_count runs plan ->
// expand -> sort -> reduce in one go; the multi-GPU path stops after expand to exchange the raw keys.
struct ExpandPlan {
    ottocov_spec spec;
    int A = 0;
    u32 W = 0;                       // symmetric range |dt| <= W ...
    bool general = false;            // ... or the general range dt_lo <= dt <= dt_hi
    int64_t dt_lo = 0, dt_hi = 0;
    bool sym = false;
    u32 user_min = 1;
    int aid_bits = 1;
    u64 P = 0;                       // keys to emit (half pairs when sym)
    std::vector<Segment*> segs;
    ~ExpandPlan() { for (auto* s : segs) delete s; }
};

void free_plan(ottocov_ctx* ctx) {
    delete static_cast<ExpandPlan*>(ctx->plan);
    ctx->plan = nullptr;
}

static ottocov_table* make_empty_table(int aid_bits) {
    ottocov_table* t = new ottocov_table();
    t->aid_bits = aid_bits;
    return t;
}

static ExpandPlan* make_plan(ottocov_ctx* ctx, const ottocov_spec* spec, bool distributed) {
    if (!ctx->loaded) COV_THROW(OTTOCOV_ERR_STATE, "co-event counting before ottocov_load_events");
    if (spec->type_this < 0 || spec->type_this > 2) COV_THROW(OTTOCOV_ERR_ARG, "type_this must be 0..2");
    if (spec->next_mask == 0 || spec->next_mask > 7) COV_THROW(OTTOCOV_ERR_ARG, "next_mask must be 1..7");
    if (spec->window < 0) COV_THROW(OTTOCOV_ERR_ARG, "window must be >= 0");
    ExpandPlan* pl = new ExpandPlan();
    try {
        pl->spec = *spec;
        pl->A = spec->type_this;
        // |dt| <= window (count_co_events.py:69) inside the pre-filter dt_min <= dt <= dt_max (:33-36, config.py:41-42)
        const int64_t pre_lo = (spec->flags & OTTOCOV_DT_RANGE) ? spec->dt_min : -86400;
        const int64_t pre_hi = (spec->flags & OTTOCOV_DT_RANGE) ? spec->dt_max : 86400;
        pl->dt_lo = pre_lo > -spec->window ? pre_lo : -spec->window;
        pl->dt_hi = pre_hi < spec->window ? pre_hi : spec->window;
        pl->general = pl->dt_lo != -pl->dt_hi || pl->dt_hi > 0x7FFFFFFFll || pl->dt_hi < 0;
        pl->W = pl->general ? 0u : (u32)pl->dt_hi;
        pl->aid_bits = ctx->info.aid_bits > 0 ? ctx->info.aid_bits : 1;
        pl->user_min = spec->min_count > 1 ? spec->min_count : 1;
        // symmetric shortcut: one canonical key per unordered event pair, mirrored after the reduce.  It
        // pays when few rows are left to mirror (a threshold) or when the keys are about to cross NVLink
        // anyway; OTTOCOV_SYM_OFF / OTTOCOV_SYM_ON force the choice.
        const bool sym_kind = spec->next_mask == (1u << pl->A) && !pl->general;
        pl->sym = sym_kind && !(spec->flags & OTTOCOV_SYM_OFF) &&
                  ((spec->flags & OTTOCOV_SYM_ON) || pl->user_min > 1 || distributed);
        const TypeArray& src = ctx->ta[pl->A];
        for (int B = 0; B < 3; ++B) {
            if (!((spec->next_mask >> B) & 1)) continue;
            const TypeArray& tgt = ctx->ta[B];
            if (src.n == 0 || tgt.n == 0) continue;
            Segment* sg = new Segment();
            pl->segs.push_back(sg);
            const bool self_in = (pl->A == B) && pl->dt_lo <= 0 && pl->dt_hi >= 0;
            sg->self = pl->general ? self_in : (pl->A == B);
            // walk the rarer side: carts / orders are ~10x / ~40x rarer than clicks, so for click -> cart-or-buy the
            // window pass runs over the target events and looks for their sources
            sg->swap = (pl->A != B) && !pl->sym && tgt.n * 2 < src.n;
            const TypeArray& own = sg->swap ? tgt : src;                 // record owners
            const TypeArray& rng = sg->swap ? src : tgt;                 // events of their ranges
            sg->own_type = sg->swap ? B : pl->A;
            sg->tgt_type = sg->swap ? pl->A : B;
            const int ot = sg->own_type, rt = sg->tgt_type;
            const u32* xr = nullptr;
            if (!sg->self) xr = (rt == (ot + 1) % 3) ? own.xrank[0] : own.xrank[1];
            // owner at t, range event at t': the kind asks dt_lo <= t_tgt - t_src <= dt_hi; seen from a target owner the
            // range is -dt_hi <= t_src - t_tgt <= -dt_lo
            const int64_t r_lo = sg->swap ? -pl->dt_hi : pl->dt_lo, r_hi = sg->swap ? -pl->dt_lo : pl->dt_hi;
            DevBuf<u32> lo, cnt(ctx, own.n);
            if (!pl->sym || pl->general) lo.alloc(ctx, own.n);
            if (pl->general)
                COV_LAUNCH(ctx, OTTOCOV_K_WINDOW, 20.0 * own.n, window_range_kernel, (unsigned)ceil_div64(own.n, 256), 256, 0,
                           own.skey, own.n, rng.skey, (u32)rng.n, r_lo, r_hi, self_in ? 1 : 0, lo.p, cnt.p);
            else if (pl->sym)
                COV_LAUNCH(ctx, OTTOCOV_K_WINDOW, 12.0 * own.n, window_fwd_kernel, (unsigned)ceil_div64(own.n, 256), 256, 0,
                           own.skey, own.n, pl->W, (u32*)nullptr, cnt.p);
            else
                COV_LAUNCH(ctx, OTTOCOV_K_WINDOW, 20.0 * own.n, window_kernel, (unsigned)ceil_div64(own.n, 256), 256, 0,
                           own.skey, xr, own.n, rng.skey, (u32)rng.n, pl->W, lo.p, cnt.p);
            const bool implicit_lo = pl->sym && !pl->general;
            sg->rec_src.alloc(ctx, own.n); sg->rec_off.alloc(ctx, own.n);
            if (!implicit_lo) sg->rec_lo.alloc(ctx, own.n);
            WindowRecords f;
            f.lo = implicit_lo ? nullptr : lo.p; f.cnt = cnt.p;
            f.rec_src = sg->rec_src.p; f.rec_lo = implicit_lo ? nullptr : sg->rec_lo.p; f.rec_off = sg->rec_off.p;
            u64 tot[2];
            scan_apply(ctx, OTTOCOV_K_WINDOW, f, own.n, tot, (implicit_lo ? 4.0 : 8.0) * own.n + 8.0 * own.n + (implicit_lo ? 12.0 : 16.0) * own.n);
            sg->n_pairs = tot[0];
            sg->n_rec = tot[1];
            pl->P += tot[0];
        }
    } catch (...) {
        delete pl;
        throw;
    }
    return pl;
}

// keys of plan outputs [c0, c1) -> dst[0 .. c1-c0).  n_dest > 1 also stamps hash(aid of the key) % n_dest
// into key bits [56, 64) so one radix pass on those bits groups the keys by destination rank.
// mix != nullptr: keys are written mixed (bucketed hash reduce), n_dest must be 0.
// hist_passes + ghist (device, [passes][RS_RADIX], zeroed by the caller): also accumulate the digit histograms
// of those distribution passes over the emitted keys.
static void expand_range(ottocov_ctx* ctx, const ExpandPlan* pl, u64 c0, u64 c1, u64* dst_base, u32 n_dest,
                         const KeyMix* mix = nullptr, const PassList* hist_passes = nullptr, u64* ghist = nullptr) {
    const KeyMix mx = mix ? *mix : KeyMix();
    PassList hp;
    hp.n = 0;
    if (hist_passes && ghist) hp = *hist_passes; else ghist = nullptr;
    const size_t hist_smem = (size_t)hp.n * RS_RADIX * sizeof(u32);
    u64 seg_start = 0;
    for (Segment* sg : pl->segs) {
        const u64 seg_end = seg_start + sg->n_pairs;
        const u64 a = c0 > seg_start ? c0 : seg_start;
        const u64 b = c1 < seg_end ? c1 : seg_end;
        if (a < b) {
            const u64 ob = a - seg_start, oe = b - seg_start;       // segment-local output range
            const int64_t n_tiles = ceil_div64((int64_t)(oe - ob), EX_TILE);
            DevBuf<u32> tile_rec(ctx, n_tiles + 1);
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, 0, tile_search_kernel, (unsigned)ceil_div64(n_tiles + 1, 256), 256, 0,
                       sg->rec_off.p, (int64_t)sg->n_rec, ob, oe, n_tiles, tile_rec.p);
            const TypeArray& own = ctx->ta[sg->own_type];
            const TypeArray& tgt = ctx->ta[sg->tgt_type];
            u64* dst = dst_base + (a - c0);
            const double bytes = 8.0 * (double)(oe - ob);
            const unsigned grid = (unsigned)imin64(n_tiles, (int64_t)ctx->num_sms * 5);     // 5 CTAs / SM at 48 registers
#define EX_LAUNCH(MODE_, MIX_)                                                                                      \
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, bytes, (expand_kernel<MODE_, MIX_>), grid, EX_THREADS, hist_smem,          \
                       sg->rec_src.p, sg->rec_lo.p, sg->rec_off.p, tile_rec.p, own.aid, tgt.aid, ob, oe, dst, n_dest, mx, \
                       n_tiles, hp, ghist)
            if (mix) {
                if (pl->sym) EX_LAUNCH(EXM_CANON, true);
                else if (sg->self) EX_LAUNCH(EXM_SELF, true);
                else if (sg->swap) EX_LAUNCH(EXM_SWAP, true);
                else EX_LAUNCH(EXM_CROSS, true);
            } else {
                if (pl->sym) EX_LAUNCH(EXM_CANON, false);
                else if (sg->self) EX_LAUNCH(EXM_SELF, false);
                else if (sg->swap) EX_LAUNCH(EXM_SWAP, false);
                else EX_LAUNCH(EXM_CROSS, false);
            }
#undef EX_LAUNCH
        }
        seg_start = seg_end;
    }
}

// The whole plan expanded by expand_scatter_kernel: every segment appends to the same digit regions.
// ghist: device [n_dest][rest.n][RS_RADIX], zeroed by the caller.
static void expand_scatter_all(ottocov_ctx* ctx, const ExpandPlan* pl, const KeyMix& mix, const PassList& rest, u64* ghist,
                               const ScatterArgs& sa) {
    const u32 nd = sa.n_dest > 1 ? sa.n_dest : 1u;
    const size_t hist_smem = (size_t)nd * rest.n * RS_RADIX * sizeof(u32);
    for (Segment* sg : pl->segs) {
        if (sg->n_pairs == 0) continue;
        const int64_t n_tiles = ceil_div64((int64_t)sg->n_pairs, EX_TILE);
        DevBuf<u32> tile_rec(ctx, n_tiles + 1);
        COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, 0, tile_search_kernel, (unsigned)ceil_div64(n_tiles + 1, 256), 256, 0,
                   sg->rec_off.p, (int64_t)sg->n_rec, (u64)0, (u64)sg->n_pairs, n_tiles, tile_rec.p);
        const TypeArray& own = ctx->ta[sg->own_type];
        const TypeArray& tgt = ctx->ta[sg->tgt_type];
        const double bytes = 8.0 * (double)sg->n_pairs;
#define EXS_LAUNCH(MODE_)                                                                                       \
        do {                                                                                                            \
            auto kern = sa.n_dest > 1 ? expand_scatter_kernel<MODE_, true> : expand_scatter_kernel<MODE_, false>;       \
            if (hist_smem > 8192) cov_func_smem(ctx, (const void*)kern, hist_smem);                                     \
            int per_sm = 0;                                                                                             \
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EX_THREADS, hist_smem));            \
            const unsigned grid = (unsigned)imin64(n_tiles, (int64_t)ctx->num_sms * (per_sm > 0 ? per_sm : 1));          \
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, bytes, kern, grid, EX_THREADS, hist_smem, sg->rec_src.p, sg->rec_lo.p,    \
                       sg->rec_off.p, tile_rec.p, own.aid, tgt.aid, (u64)0, (u64)sg->n_pairs, mix, n_tiles, rest, ghist, sa); \
        } while (0)
        if (pl->sym) EXS_LAUNCH(EXM_CANON);
        else if (sg->self) EXS_LAUNCH(EXM_SELF);
        else if (sg->swap) EXS_LAUNCH(EXM_SWAP);
        else EXS_LAUNCH(EXM_CROSS);
#undef EXS_LAUNCH
    }
}

// Region table of one chunk (digits [d_lo, d_hi)): uniform regions of `cap` keys, or -- after an overflow -- exact
// regions sized by the fill counters of the failed attempt (`fill`, which kept counting past the capacity).
__global__ void __launch_bounds__(RS_RADIX) region_table_kernel(u64* reg_base, u64* reg_cap, u64* reg_off, u64 base_addr, u64 cap,
                                                                const unsigned long long* fill, u64 pad, u32 d_lo, u32 d_hi) {
    __shared__ u64 s_warp[RS_RADIX / 32 + 1];
    const u32 d = threadIdx.x;
    const bool in = d >= d_lo && d < d_hi;
    u64 c = 0;
    if (in) c = fill ? (((u64)fill[d] + pad + 1) & ~1ull) : cap;
    u64 tot;
    const u64 off = block_exclusive_scan<u64, RS_RADIX>(c, s_warp, &tot);
    reg_cap[d] = c;
    reg_off[d] = off;
    reg_base[d] = in ? base_addr + off * 8ull : 0ull;
}

// Slack of a digit region over its expected share, in percent (test knob: a negative value forces the overflow
// retry).  Hashed digits are uniform; what the slack absorbs is the multiplicity of hot pairs.
static int fuse_slack_pct() {
    static int v = -1000;
    if (v == -1000) { const char* e = getenv("OTTOCOV_FUSE_SLACK_PCT"); v = e ? atoi(e) : 25; if (v < -90) v = -90; }
    return v;
}
static bool fuse_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OTTOCOV_NO_FUSED_PASS"); v = (e && atoi(e)) ? 0 : 1; }
    return v != 0;
}

// Bucketed hash reduce with its first distribution pass fused into the expansion (hash_reduce.cu, HashPre).
// The hash space is taken `digits per chunk` digit regions at a time when the key buffers would not fit the pair
// budget; every chunk holds complete sums, so the threshold and the symmetric mirror stay fused whatever the chunking.
// Returns the partial tables (disjoint key sets) in `partials`.
static void count_fused(ottocov_ctx* ctx, const ExpandPlan* pl, const KeyMix& mix, u64 budget, u32 min_count,
                        std::vector<ottocov_table*>& partials, ottocov_count_info& ci) {
    const u64 P = pl->P;
    const int bb_big = hashed_big_bucket_bits((int64_t)P, mix.kb);      // whole-bucket reduce when it saves a pass
    const bool big = bb_big > 0;
    const int bb = big ? bb_big : hashed_bucket_bits((int64_t)P, mix.kb);
    BitField bucket_field[1] = {{mix.kb - bb, mix.kb}};
    const PassList full = make_pass_list(bucket_field, 1);
    BitField rest_field[1] = {{mix.kb - bb + full.bits[0], mix.kb}};
    const PassList rest = make_pass_list(rest_field, 1);
    const u32 n_digits = 1u << full.bits[0];
    DevBuf<u64> reg(ctx, 3 * RS_RADIX);                                  // byte addresses | capacities | key offsets
    u64* reg_base = reg.p; u64* reg_cap = reg.p + RS_RADIX; u64* reg_off = reg.p + 2 * RS_RADIX;
    DevBuf<unsigned long long> cursor(ctx, 2 * RS_RADIX);                // fill counters | copy kept for an exact retry
    DevBuf<unsigned long long> ctr(ctx, 2);
    DevBuf<u64> ghist(ctx, (size_t)rest.n * RS_RADIX);
    const u64 cap = (((u64)((double)P / n_digits * (1.0 + fuse_slack_pct() / 100.0)) + 4096) + 1) & ~1ull;
    const u64 pad = 64;
    unsigned long long fill[RS_RADIX];
    bool exact = false;                  // regions of this chunk sized by `fill` (second attempt after an overflow)
    u32 d0 = 0;
    while (d0 < n_digits) {
        u32 dc = 0;                      // digit regions of this chunk
        u64 total = 0;                   // keys its buffer holds
        if (!exact) {
            dc = (u32)(budget / cap);
            if (dc < 1) dc = 1;
            if (dc > n_digits - d0) dc = n_digits - d0;
            total = (u64)dc * cap;
        } else {
            while (d0 + dc < n_digits) {
                const u64 need = ((u64)fill[d0 + dc] + pad + 1) & ~1ull;
                if (dc > 0 && total + need > budget) break;
                total += need; ++dc;
            }
        }
        const bool single = dc == n_digits;
        DevBuf<u64> keys(ctx, (size_t)total);
        if (exact) CUDA_CHECK(cudaMemcpyAsync(cursor.p + RS_RADIX, cursor.p, RS_RADIX * sizeof(unsigned long long),
                                              cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(cursor.p, 0, RS_RADIX * sizeof(unsigned long long), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(ctr.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(ghist.p, 0, (size_t)rest.n * RS_RADIX * sizeof(u64), ctx->stream));
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, region_table_kernel, 1, RS_RADIX, 0, reg_base, reg_cap, reg_off,
                   reinterpret_cast<u64>(keys.p), cap, exact ? cursor.p + RS_RADIX : (const unsigned long long*)nullptr, pad,
                   d0, d0 + dc);
        ScatterArgs sa;
        sa.reg_base = reg_base; sa.reg_cap = reg_cap; sa.cursor = cursor.p; sa.flags = reinterpret_cast<u32*>(ctr.p + 1);
        sa.sh1 = full.shift[0]; sa.sub_bits = full.bits[0]; sa.n_digits = n_digits; sa.d_lo = d0; sa.d_hi = d0 + dc;
        sa.n_dest = 0;
        expand_scatter_all(ctx, pl, mix, rest, ghist.p, sa);
        cov_trace(ctx, "count: expand + first pass (fused)");
        // keys of this chunk: all of them when the chunk covers every digit (no host round trip), else the sum of
        // the fill counters
        u64 n_c = P;
        bool overflow = false;
        if (!single || exact) {
            cov_readback(ctx, fill, cursor.p, RS_RADIX * sizeof(unsigned long long));
            unsigned long long fl[2];
            cov_readback(ctx, fl, ctr.p, sizeof(fl));
            overflow = (fl[1] & HR_FLAG_FUSED_OVERFLOW) != 0;
            n_c = 0;
            for (u32 d = d0; d < d0 + dc; ++d) n_c += fill[d];
        }
        ottocov_table* part = nullptr;
        if (!overflow && n_c > 0) {
            DevBuf<u64> alt(ctx, (size_t)n_c);
            HashPre pre;
            pre.bb = bb; pre.first_bits = full.bits[0]; pre.big = big;
            pre.seg_cnt = reinterpret_cast<const u64*>(cursor.p + d0); pre.n_a = 1; pre.n_b = (int)dc;
            pre.seg_off = reg_off + d0; pre.ctr = ctr.p;
            int passes = 0;
            try {
                part = hashed_reduce(ctx, keys.p, alt.p, (int64_t)n_c, mix, min_count, pl->sym, pl->sym, &passes, ghist.p, &pre);
                ci.sort_passes = passes;
            } catch (const FusedOverflow&) {
                overflow = true;
                cov_readback(ctx, fill, cursor.p, RS_RADIX * sizeof(unsigned long long));
            }
        }
        if (overflow) {              // the counters kept counting: they now hold what each region really needs
            if (exact) COV_THROW(OTTOCOV_ERR_CUDA, "fused first pass overflowed regions sized by its own counters");
            exact = true;
            cov_trace(ctx, "count: region overflow, retrying with exact regions");
            continue;                // same d0
        }
        exact = false;
        if (part) partials.push_back(part);
        ci.n_chunks += 1;
        ci.fused = big ? 2 : 1;
        d0 += dc;
        cov_trace(ctx, "count: bucket passes + hash reduce");
    }
}

// cudaMemGetInfo is a driver round trip (it serialises with whatever else talks to the driver), and free + parked
// bytes barely move between steps: the figure is cached per context and refreshed when a count would not fit it.
static u64 auto_budget(ottocov_ctx* ctx, bool refresh) {
    if (ctx->budget_cache && !refresh) return ctx->budget_cache;
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    // blocks parked in our cache count as free; 16 B per pair for the double buffer plus <= 12 B per
    // pair for the reduced table
    u64 budget = (u64)((double)(free_b + ctx->cached_bytes) * 0.6 / 28.0);
    ctx->budget_cache = budget < (1u << 20) ? (1u << 20) : budget;
    return ctx->budget_cache;
}

ottocov_table* count_impl(ottocov_ctx* ctx, const ottocov_spec* spec) {
    ottocov_count_info& ci = ctx->last_count;
    memset(&ci, 0, sizeof(ci));
    cov_trace(ctx, "count: enter");
    ExpandPlan* pl = make_plan(ctx, spec, false);
    cov_trace(ctx, "count: plan (window+scan)");
    struct PlanGuard { ExpandPlan* p; ~PlanGuard() { delete p; } } plan_guard{pl};
    const u64 P = pl->P;
    const bool sym = pl->sym;
    const int aid_bits = pl->aid_bits;
    ci.n_pairs = (int64_t)(sym ? 2 * P : P);       // ordered co-event pairs, as the reference counts them
    if (P == 0) { ci.n_chunks = 0; return make_empty_table(aid_bits); }

    // ---- chunking by pair budget ----------------------------------------------------------------------
    u64 budget = spec->pair_budget > 0 ? (u64)spec->pair_budget : auto_budget(ctx, false);
    if (spec->pair_budget <= 0 && P > budget) budget = auto_budget(ctx, true);      // before chunking, look again
    cov_trace(ctx, "count: budget");
    budget = (budget / EX_TILE) * EX_TILE;
    if (budget == 0) budget = EX_TILE;

    const u32 fused_min = (P <= budget) ? pl->user_min : 1;     // thresholds apply to complete sums only
    // bucketed hash reduce (hash_reduce.cu) instead of full sort + run-length reduce: pays when a threshold
    // leaves few rows to bring back into key order; OTTOCOV_HASH_ON / OTTOCOV_HASH_OFF force the choice
    const bool hashed = hashed_reduce_supported(aid_bits) && !(spec->flags & OTTOCOV_HASH_OFF) &&
                        ((spec->flags & OTTOCOV_HASH_ON) || fused_min > 1);
    const bool single = P <= budget;
    const KeyMix mix = make_key_mix(aid_bits);
    bool mirrored = false;
    std::vector<ottocov_table*> partials;
    struct PartGuard {
        ottocov_ctx* c; std::vector<ottocov_table*>& v;
        ~PartGuard() { for (auto* t : v) { dev_free(c, t->keys); dev_free(c, t->count); delete t; } }
    } pguard{ctx, partials};

    // Fused path: bucketed hash reduce whose first pass runs inside the expansion.  Chunks are hash-digit ranges, so
    // sums are complete per chunk and the user's threshold stays fused whatever the budget.
    const int bb_all = hashed_bucket_bits((int64_t)P, mix.kb);
    const bool hashed_any = hashed_reduce_supported(aid_bits) && !(spec->flags & OTTOCOV_HASH_OFF) &&
                            ((spec->flags & OTTOCOV_HASH_ON) || pl->user_min > 1);
    const bool fused = hashed_any && fuse_enabled() && bb_all > RS_MAX_BITS;     // at least one pass after the fused one
    if (fused) {
        count_fused(ctx, pl, mix, budget, pl->user_min, partials, ci);
        mirrored = sym;
        ottocov_table* result;
        if (partials.empty()) result = make_empty_table(aid_bits);
        else if (partials.size() == 1) { result = partials[0]; partials.clear(); }
        else result = merge_tables_impl(ctx, partials.data(), (int)partials.size());     // disjoint key sets: a sort
        ci.n_unique = result->n;
        return result;
    }

    BitField fields[2] = {{0, aid_bits}, {32, 32 + aid_bits}};
    for (u64 c0 = 0; c0 < P; c0 += budget) {
        const u64 c1 = (c0 + budget < P) ? c0 + budget : P;
        const u64 cn = c1 - c0;
        DevBuf<u64> keys(ctx, cn), alt(ctx, cn);
        cov_trace(ctx, "count: alloc keys");
        if (hashed) {
            // the expansion also builds the digit histograms of the bucket passes (no extra read of the keys)
            const int bb = hashed_bucket_bits((int64_t)cn, mix.kb);
            BitField bucket_field[1] = {{mix.kb - bb, mix.kb}};
            const PassList hp = make_pass_list(bucket_field, 1);
            DevBuf<u64> ghist;
            if (hp.n > 0 && cn > 1) {
                ghist.alloc(ctx, (size_t)hp.n * RS_RADIX);
                CUDA_CHECK(cudaMemsetAsync(ghist.p, 0, (size_t)hp.n * RS_RADIX * sizeof(u64), ctx->stream));
            }
            expand_range(ctx, pl, c0, c1, keys.p, 0, &mix, &hp, ghist.p);
            cov_trace(ctx, "count: expand (mixed keys)");
            int passes = 0;
            mirrored = sym && single;                  // the mirrored rows come out of the same table scan
            ottocov_table* part = hashed_reduce(ctx, keys.p, alt.p, (int64_t)cn, mix, fused_min, sym, mirrored, &passes,
                                                ghist.p);
            partials.push_back(part);
            ci.sort_passes = passes;
            cov_trace(ctx, "count: bucket passes + hash reduce");
            ci.n_chunks += 1;
            continue;
        }
        expand_range(ctx, pl, c0, c1, keys.p, 0);
        cov_trace(ctx, "count: expand");
        u64* k = keys.p; u64* ka = alt.p; u32* v = nullptr; u32* va = nullptr;
        ci.sort_passes = radix_sort_pairs(ctx, k, ka, v, va, (int64_t)cn, fields, 2);
        cov_trace(ctx, "count: sort");
        ottocov_table* part = new ottocov_table();
        part->aid_bits = aid_bits;
        partials.push_back(part);
        reduce_sorted(ctx, k, nullptr, (int64_t)cn, fused_min, sym, &part->keys, &part->count, &part->n);
        cov_trace(ctx, "count: reduce");
        ci.n_chunks += 1;
    }

    ottocov_table* result;
    if (partials.size() == 1) {
        result = partials[0];
        partials.clear();
    } else {
        result = merge_tables_impl(ctx, partials.data(), (int)partials.size());
        if (pl->user_min > 1) {
            ottocov_table* f = filter_table_impl(ctx, result, pl->user_min);
            dev_free(ctx, result->keys); dev_free(ctx, result->count); delete result;
            result = f;
        }
    }
    if (sym && !mirrored) {                         // (a, b, c) -> also (b, a, c)
        ottocov_table* full = mirror_table_impl(ctx, result, false);
        dev_free(ctx, result->keys); dev_free(ctx, result->count); delete result;
        result = full;
    }
    cov_trace(ctx, "count: mirror/merge");
    ci.n_unique = result->n;
    return result;
}

// ---- streamed ingest + count (include/ottocov.h, ottocov_count_parts) ----------------------------------------------
// ---- the session column crosses PCIe run-length encoded ---------------------------------------------------------------
// The ETL writes each part ordered by session (etl/jsonl_to_parquet.py:59-84), so the 4-byte session column -- 31 % of
// the 13 B/event the loader needs -- is ~17 equal values in a row.  Host threads turn it into (session id, first row)
// runs while the DMA engine is busy with the aid and type columns; the runs (8 B per SESSION instead of 4 B per EVENT)
// are copied instead and a kernel writes the column back out on the device.  A part whose session column does not
// compress (rows in arbitrary order) is copied raw.
struct RleTask {
    const int32_t* session;
    int64_t rows;
    u32* out;                 // [2 * cap] (session id, first row) pairs, in page-locked memory
    int64_t cap;              // runs that fit; beyond that the part is copied raw
    int64_t n_runs;           // result: -1 = does not compress
    std::atomic<int> done{0};
};

// Branch-free in the common case (1.2 ns per value on one core against 2.6 for the obvious loop): the candidate run
// is always written at slot k and k only advances when the value changed.
static void rle_encode(RleTask* t) {
    const int32_t* s = t->session;
    const int64_t n = t->rows, cap = t->cap;
    u32* o = t->out;
    int64_t k = 0;
    if (n > 0) {
        if (cap < 2) k = -1;
        else {
            o[0] = (u32)s[0]; o[1] = 0; k = 1;
            int64_t i = 1;
            while (i < n && k >= 0) {
                const int64_t e = i + 1024 < n ? i + 1024 : n;
                if (k + (e - i) >= cap) {                      // close to the capacity: the careful loop
                    for (; i < e; ++i)
                        if (s[i] != s[i - 1]) {
                            if (k >= cap) { k = -1; break; }
                            o[2 * k] = (u32)s[i]; o[2 * k + 1] = (u32)i; ++k;
                        }
                } else {
                    for (; i < e; ++i) { o[2 * k] = (u32)s[i]; o[2 * k + 1] = (u32)i; k += (s[i] != s[i - 1]); }
                }
            }
        }
    }
    t->n_runs = k;
    t->done.store(1, std::memory_order_release);
}

// one thread per run: rows [first row of the run, first row of the next run) get the run's session id
__global__ void __launch_bounds__(256) rle_expand_kernel(const u32* __restrict__ runs, int64_t n_runs, int64_t rows,
                                                         int32_t* __restrict__ session) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const int32_t id = (int32_t)runs[2 * r];
    const int64_t a = runs[2 * r + 1];
    const int64_t b = (r + 1 < n_runs) ? (int64_t)runs[2 * r + 3] : rows;
    for (int64_t i = a; i < b; ++i) session[i] = id;
}

// per-kind state carried over the groups
struct PartsAccum {
    ottocov_spec spec;
    bool sym = false;
    bool started = false;
    u32 user_min = 1;
    int bb = 0;
    PassList full, rest;
    u32 n_digits = 0;
    u64 P_total = 0, n_pairs = 0;
    std::vector<u64*> bufs;                    // one region buffer per group (cov_alloc); after the group's passes its
    std::vector<u64*> sorted;                  //   first P_g slots hold the group's keys sorted by bucket (`sorted`)
    std::vector<u32*> bounds;                  // per group: first key of every bucket range (hash_reduce.cu)
    std::vector<u64> n_g;                      // keys of each group
    u32 rb = 1;                                // buckets per reduce CTA
    int64_t n_ranges = 0;
    DevBuf<unsigned long long> cursor;         // [groups][RS_RADIX] fill counters
    DevBuf<u64> reg;                           // [groups][3][RS_RADIX] byte addresses | capacities | key offsets
    DevBuf<u64> ghist;                         // [groups][passes][RS_RADIX]
    DevBuf<unsigned long long> ctr;
};

void count_parts_impl(ottocov_ctx* ctx, int n_parts, const int32_t* const* session, const int32_t* const* aid,
                      const int32_t* const* ts, const int8_t* const* type, const int64_t* rows, const ottocov_spec* specs,
                      int n_specs, ottocov_table** tables_out) {
    free_events(ctx);
    memset(&ctx->last_count, 0, sizeof(ctx->last_count));
    for (int k = 0; k < n_specs; ++k) tables_out[k] = nullptr;
    int64_t N = 0;
    for (int p = 0; p < n_parts; ++p) {
        if (rows[p] < 0) COV_THROW(OTTOCOV_ERR_ARG, "negative row count");
        if (rows[p] > 0 && (!session[p] || !aid[p] || !ts[p] || !type[p])) COV_THROW(OTTOCOV_ERR_ARG, "NULL column");
        N += rows[p];
    }
    ottocov_events_info total;
    memset(&total, 0, sizeof(total));
    total.was_sorted = 1;
    total.session_min = total.ts_min = 2147483647; total.session_max = total.ts_max = -2147483647 - 1;
    if (N == 0) {
        for (int k = 0; k < n_specs; ++k) tables_out[k] = make_empty_table(1);
        ctx->info = total; ctx->info.session_min = ctx->info.session_max = ctx->info.ts_min = ctx->info.ts_max = 0;
        return;
    }
    // ---- groups of consecutive parts: big enough to amortise the per-group host round trips, small enough that the
    //      work on a group hides behind the copy of the next one -- and SHRINKING towards the end, because the work on
    //      the last group is the one part of the per-group work that nothing hides (cumulative row shares below)
    static const double kShare[] = {0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.88, 0.94, 0.975, 1.0};
    const int max_groups = (int)(sizeof(kShare) / sizeof(kShare[0]));
    int n_groups = n_parts < max_groups ? n_parts : max_groups;
    std::vector<int> g_first(n_groups + 1, 0);
    {
        int64_t acc = 0; int g = 1;
        for (int p = 0; p < n_parts && g < n_groups; ++p) {
            acc += rows[p];
            const double share = n_parts < max_groups ? (double)g / n_groups : kShare[g - 1];
            if ((double)acc >= share * (double)N) { g_first[g++] = p + 1; }
        }
        for (; g <= n_groups; ++g) g_first[g] = n_parts;
        g_first[n_groups] = n_parts;
    }
    std::vector<int64_t> off(n_parts + 1, 0);
    for (int p = 0; p < n_parts; ++p) off[p + 1] = off[p] + rows[p];

    DevBuf<int32_t> d_session(ctx, N), d_aid(ctx, N), d_ts(ctx, N);
    DevBuf<int8_t> d_type(ctx, N);
    if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream;
    std::vector<cudaEvent_t> used;
    struct EventReturn {
        ottocov_ctx* c; std::vector<cudaEvent_t>& v;
        ~EventReturn() { for (cudaEvent_t e : v) c->sync_events.push_back(e); }
    } event_return{ctx, used};
    auto new_event = [&]() {
        cudaEvent_t e;
        if (!ctx->sync_events.empty()) { e = ctx->sync_events.back(); ctx->sync_events.pop_back(); }
        else CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        used.push_back(e);
        return e;
    };
    cudaEvent_t e0 = new_event();                       // the device blocks may still be in use by earlier work
    CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(cs, e0, 0));
    // ---- host threads run-length encode the session columns while the first copies are in flight ----------------------
    // Tasks are row ranges of <= 1 M rows of one part, encoded independently (a run that continues across two ranges is
    // simply two runs with the same id), handed to the threads in order so that the first groups are ready first.
    static int rle_on = -1;
    if (rle_on < 0) { const char* e = getenv("OTTOCOV_NO_SESSION_RLE"); rle_on = (e && atoi(e)) ? 0 : 1; }
    constexpr int64_t RLE_TASK_ROWS = 1 << 20;          // ~1.5 ms of one core: the first group is ready almost at once
    std::vector<std::unique_ptr<RleTask>> rle;
    std::vector<int64_t> task_row0, task_slot;              // first row (global) and staging slot (in runs) of each task
    std::vector<int> part_task0(n_parts + 1, 0);
    int64_t slot = 0;
    if (rle_on)
        for (int p = 0; p < n_parts; ++p) {
            part_task0[p] = (int)rle.size();
            for (int64_t a = 0; a < rows[p]; a += RLE_TASK_ROWS) {
                const int64_t m = imin64(RLE_TASK_ROWS, rows[p] - a);
                rle.emplace_back(new RleTask());
                rle.back()->session = session[p] + a; rle.back()->rows = m; rle.back()->cap = m / 8 + 16; rle.back()->n_runs = -1;
                task_row0.push_back(off[p] + a);
                task_slot.push_back(slot);
                slot += m / 8 + 16;
            }
        }
    part_task0[n_parts] = (int)rle.size();
    const size_t stage_bytes = (size_t)slot * 8;
    if (rle_on && stage_bytes > ctx->host_stage_bytes) {
        if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
        ctx->host_stage = nullptr; ctx->host_stage_bytes = 0;
        CUDA_CHECK(cudaHostAlloc(&ctx->host_stage, stage_bytes + stage_bytes / 8, cudaHostAllocMapped | cudaHostAllocPortable));
        ctx->host_stage_bytes = stage_bytes + stage_bytes / 8;
    }
    std::atomic<int> next_task{0};
    std::vector<std::thread> workers;
    struct JoinWorkers {                                    // never leave with a thread still reading the caller's buffers
        std::vector<std::thread>& w;
        ~JoinWorkers() { for (auto& t : w) if (t.joinable()) t.join(); }
    } join_workers{workers};
    if (rle_on && !rle.empty()) {
        const int n_tasks = (int)rle.size();
        for (int k = 0; k < n_tasks; ++k) rle[k]->out = reinterpret_cast<u32*>(ctx->host_stage) + task_slot[k] * 2;
        unsigned hw = std::thread::hardware_concurrency();
        int n_thr = (int)(hw ? hw : 4);
        if (n_thr > 32) n_thr = 32;
        if (n_thr > n_tasks) n_thr = n_tasks;
        for (int t = 0; t < n_thr; ++t)
            workers.emplace_back([&, n_tasks]() {
                for (;;) {                                  // in order: the first groups' runs are needed first
                    const int k = next_task.fetch_add(1);
                    if (k >= n_tasks) return;
                    rle_encode(rle[k].get());
                }
            });
    }
    int64_t h2d = 0;
    static int trace_on = -1;
    if (trace_on < 0) { const char* e = getenv("OTTOCOV_TRACE"); trace_on = (e && atoi(e)) ? 1 : 0; }
    cudaEvent_t tr_a = nullptr, tr_b = nullptr;
    if (trace_on) { cudaEventCreate(&tr_a); cudaEventCreate(&tr_b); cudaEventRecord(tr_a, cs); }
    ctx->begin(OTTOCOV_K_LOAD);
    // raw columns (type, aid, ts -- and session where the runs are switched off) of every group, enqueued at once: the DMA
    // engine starts at t = 0 and never waits for the host.  The session runs (small) follow on a second stream as the
    // encoder threads deliver them, group by group, while the main stream already works on the earlier groups.
    std::vector<cudaEvent_t> e_grp(n_groups), e_run(n_groups, nullptr);      // e_run[g] != null: group g's session column is on its way
    for (int g = 0; g < n_groups; ++g) {
        for (int p = g_first[g]; p < g_first[g + 1]; ++p)
            if (rows[p]) {
                CUDA_CHECK(cudaMemcpyAsync(d_type.p + off[p], type[p], rows[p], cudaMemcpyHostToDevice, cs));
                CUDA_CHECK(cudaMemcpyAsync(d_aid.p + off[p], aid[p], rows[p] * 4, cudaMemcpyHostToDevice, cs));
                CUDA_CHECK(cudaMemcpyAsync(d_ts.p + off[p], ts[p], rows[p] * 4, cudaMemcpyHostToDevice, cs));
                h2d += rows[p] * 9;
                if (!rle_on) {
                    CUDA_CHECK(cudaMemcpyAsync(d_session.p + off[p], session[p], rows[p] * 4, cudaMemcpyHostToDevice, cs));
                    h2d += rows[p] * 4;
                }
            }
        e_grp[g] = new_event();
        CUDA_CHECK(cudaEventRecord(e_grp[g], cs));
    }
    if (trace_on) cudaEventRecord(tr_b, cs);
    // The runs are NOT copied: a second stream's copies queue behind the big ones on the same DMA engine (measured: the
    // first group then waited for the whole 38 ms copy).  The staging buffer is page-locked memory, which the device can
    // read directly: rle_expand_kernel pulls the runs over PCIe itself, next to the DMA traffic.  Only a range that
    // does not compress needs a copy of its raw session column (main stream; rare: rows in arbitrary order).
    auto enqueue_runs = [&](int g) {
        if (!rle_on) return;
        for (int p = g_first[g]; p < g_first[g + 1]; ++p)
            for (int k = part_task0[p]; k < part_task0[p + 1]; ++k) {
                while (!rle[k]->done.load(std::memory_order_acquire)) std::this_thread::yield();
                if (rle[k]->n_runs >= 0) {
                    h2d += rle[k]->n_runs * 8;
                } else {
                    CUDA_CHECK(cudaMemcpyAsync(d_session.p + task_row0[k], rle[k]->session, rle[k]->rows * 4,
                                               cudaMemcpyHostToDevice, ctx->stream));
                    h2d += rle[k]->rows * 4;
                }
            }
        e_run[g] = e_grp[g];            // marks the group as handled
    };
    ctx->end(OTTOCOV_K_LOAD, (double)h2d);
    ctx->stats[OTTOCOV_K_LOAD].launches -= 1;           // copies, not kernels
    // no exit path may return while a copy still reads the caller's host buffers, or leave the main stream ahead of them
    struct JoinCopies {
        ottocov_ctx* c; cudaEvent_t last; bool done = false;
        ~JoinCopies() {
            cudaStreamWaitEvent(c->stream, last, 0);
            if (!done) cudaStreamSynchronize(c->copy_stream);
        }
    } join_copies{ctx, e_grp[n_groups - 1]};

    // The key width (bits of the largest aid) must be fixed before the first key is mixed.  It is taken from the FIRST
    // group (the loader finds it anyway) and checked on every later one: an aid that needs more bits sends the call to
    // the plain path below.  (Waiting for every aid column to land first would hold the first group back by a third of
    // the copy; scanning them on the host costs more than it saves.)
    int aid_bits = 0, aid_max = 0;
    KeyMix mix = make_key_mix(1);

    std::vector<PartsAccum> acc(n_specs);
    struct AccGuard {
        ottocov_ctx* c; std::vector<PartsAccum>& v;
        ~AccGuard() { for (auto& a : v) { for (u64* b : a.bufs) dev_free(c, b); for (u32* b : a.bounds) dev_free(c, b); } }
    } acc_guard{ctx, acc};
    bool fallback = false;
    std::vector<int64_t> g_rows(n_groups);

    cov_trace(ctx, "parts: copies enqueued, key width known");
    for (int g = 0; g < n_groups && !fallback; ++g) {
        const int64_t r0 = off[g_first[g]], r1 = off[g_first[g + 1]];
        g_rows[g] = r1 - r0;
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, e_grp[g], 0));
        enqueue_runs(g);
        if (r1 == r0) continue;
        if (rle_on)
            for (int k = part_task0[g_first[g]]; k < part_task0[g_first[g + 1]]; ++k)
                if (rle[k]->n_runs > 0)
                    COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 4.0 * rle[k]->rows + 8.0 * rle[k]->n_runs, rle_expand_kernel,
                               (unsigned)ceil_div64(rle[k]->n_runs, 256), 256, 0, (const u32*)rle[k]->out, rle[k]->n_runs,
                               rle[k]->rows, d_session.p + task_row0[k]);
        cov_trace(ctx, "parts:   group's columns have landed");
        load_events_impl(ctx, d_session.p + r0, d_aid.p + r0, d_ts.p + r0, d_type.p + r0, r1 - r0, OTTOCOV_DEVICE);
        cov_trace(ctx, "parts:   loader");
        const ottocov_events_info gi = ctx->info;
        total.n_rows_in += gi.n_rows_in; total.n_events += gi.n_events;
        for (int t = 0; t < 3; ++t) total.n_by_type[t] += gi.n_by_type[t];
        total.session_min = gi.session_min < total.session_min ? gi.session_min : total.session_min;
        total.session_max = gi.session_max > total.session_max ? gi.session_max : total.session_max;
        total.ts_min = gi.ts_min < total.ts_min ? gi.ts_min : total.ts_min;
        total.ts_max = gi.ts_max > total.ts_max ? gi.ts_max : total.ts_max;
        total.was_sorted = total.was_sorted && gi.was_sorted;
        aid_max = gi.aid_max > aid_max ? gi.aid_max : aid_max;
        if (aid_bits == 0) {                            // first group with rows: it fixes the key width
            aid_bits = gi.aid_bits;
            if (!hashed_reduce_supported(aid_bits)) { fallback = true; break; }      // keys too wide for the hash path
            mix = make_key_mix(aid_bits);
        } else if (gi.aid_bits > aid_bits) { fallback = true; break; }
        ctx->info.aid_bits = aid_bits;                  // keys of every group are made with that width
        for (int k = 0; k < n_specs; ++k) {
            PartsAccum& a = acc[k];
            ExpandPlan* pl = make_plan(ctx, &specs[k], false);
            cov_trace(ctx, "parts:   window");
            struct PlanGuard { ExpandPlan* p; ~PlanGuard() { delete p; } } plan_guard{pl};
            if (!a.started) {
                a.started = true;
                a.spec = specs[k];
                a.sym = pl->sym;
                a.user_min = pl->user_min;
                // bucket bits from the key count this kind will have if the other groups look like this one
                const u64 P_est = (u64)((double)pl->P * (double)N / (double)(r1 - r0)) + 1;
                a.bb = hashed_bucket_bits((int64_t)P_est, mix.kb);
                if (a.bb <= RS_MAX_BITS) a.bb = (RS_MAX_BITS + 1 < mix.kb) ? RS_MAX_BITS + 1 : mix.kb;   // >= 1 pass after the fused one
                BitField f[1] = {{mix.kb - a.bb, mix.kb}};
                a.full = make_pass_list(f, 1);
                if (a.full.n < 2) { fallback = true; break; }         // keys of < 9 bits: not worth a fused pass
                BitField rf[1] = {{mix.kb - a.bb + a.full.bits[0], mix.kb}};
                a.rest = make_pass_list(rf, 1);
                a.n_digits = 1u << a.full.bits[0];
                a.rb = hashed_range_buckets((int64_t)P_est, a.bb);
                a.n_ranges = hashed_n_ranges(a.bb, a.rb);
                a.cursor.alloc(ctx, (size_t)n_groups * RS_RADIX);
                a.reg.alloc(ctx, (size_t)n_groups * 3 * RS_RADIX);
                a.ghist.alloc(ctx, (size_t)n_groups * a.rest.n * RS_RADIX);
                a.ctr.alloc(ctx, 2);
                CUDA_CHECK(cudaMemsetAsync(a.cursor.p, 0, (size_t)n_groups * RS_RADIX * sizeof(unsigned long long), ctx->stream));
                CUDA_CHECK(cudaMemsetAsync(a.ghist.p, 0, (size_t)n_groups * a.rest.n * RS_RADIX * sizeof(u64), ctx->stream));
                CUDA_CHECK(cudaMemsetAsync(a.ctr.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
                a.bufs.assign(n_groups, nullptr); a.sorted.assign(n_groups, nullptr); a.bounds.assign(n_groups, nullptr);
                a.n_g.assign(n_groups, 0);
            }
            if (pl->sym != a.sym) COV_THROW(OTTOCOV_ERR_CUDA, "inconsistent symmetric choice across groups");
            a.n_pairs += pl->sym ? 2 * pl->P : pl->P;
            const u64 capg = ((((u64)((double)pl->P / a.n_digits * (1.0 + fuse_slack_pct() / 100.0)) + 4096) + 1) & ~1ull);
            if (pl->P == 0) continue;
            a.bufs[g] = (u64*)cov_alloc(ctx, (size_t)a.n_digits * capg * 8);
            a.n_g[g] = pl->P;
            a.P_total += pl->P;
            u64* reg_base = a.reg.p + (size_t)g * 3 * RS_RADIX;
            u64* reg_off = reg_base + 2 * RS_RADIX;
            COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, region_table_kernel, 1, RS_RADIX, 0, reg_base, reg_base + RS_RADIX, reg_off,
                       reinterpret_cast<u64>(a.bufs[g]), capg, (const unsigned long long*)nullptr, (u64)0, 0u, a.n_digits);
            u64* ghist_g = a.ghist.p + (size_t)g * a.rest.n * RS_RADIX;
            ScatterArgs sa;
            sa.reg_base = reg_base; sa.reg_cap = reg_base + RS_RADIX; sa.cursor = a.cursor.p + (size_t)g * RS_RADIX;
            sa.flags = reinterpret_cast<u32*>(a.ctr.p + 1);
            sa.sh1 = a.full.shift[0]; sa.sub_bits = a.full.bits[0]; sa.n_digits = a.n_digits; sa.d_lo = 0; sa.d_hi = a.n_digits;
            sa.n_dest = 0;
            expand_scatter_all(ctx, pl, mix, a.rest, ghist_g, sa);
            cov_trace(ctx, "parts:   expansion + first pass");
            // the group's remaining passes right away, behind the copy of the next groups: afterwards its keys sit sorted
            // by bucket at the start of its region buffer, with the table of bucket-range boundaries next to them
            const u32* abort_flag = reinterpret_cast<const u32*>(a.ctr.p + 1);
            u64* sk = a.bufs[g];
            u64* ska = (u64*)cov_alloc(ctx, (size_t)pl->P * 8);
            u32* v = nullptr; u32* va = nullptr;
            try {
                radix_sort_passes(ctx, sk, ska, v, va, (int64_t)pl->P, a.rest, ghist_g,
                                  reinterpret_cast<const u64*>(a.cursor.p + (size_t)g * RS_RADIX), reg_off, 1, (int)a.n_digits, abort_flag);
            } catch (...) { dev_free(ctx, sk == a.bufs[g] ? ska : sk); throw; }
            if (sk == a.bufs[g]) dev_free(ctx, ska);
            else { dev_free(ctx, a.bufs[g]); a.bufs[g] = sk; }            // result landed in the second buffer: keep that one
            a.sorted[g] = sk;
            a.bounds[g] = (u32*)cov_alloc(ctx, (size_t)(a.n_ranges + 1) * 4);
            hashed_range_bounds(ctx, sk, (int64_t)pl->P, mix.kb - a.bb, a.rb, a.n_ranges, a.bounds[g], abort_flag);
        }
        cov_trace(ctx, "parts: group done (load, window, expansion, passes, bounds)");
    }
    for (int g = 0; g < n_groups; ++g)                  // a group skipped by a fallback: its runs must still be on their way
        if (!e_run[g]) enqueue_runs(g);
    ctx->h2d_bytes_last = h2d;
    total.aid_max = aid_max;
    total.aid_bits = aid_bits ? aid_bits : 1;

    // ---- remaining passes + hash reduce per kind, over the regions of all groups -------------------------------------------
    if (!fallback) {
        for (int k = 0; k < n_specs && !fallback; ++k) {
            PartsAccum& a = acc[k];
            memset(&ctx->last_count, 0, sizeof(ctx->last_count));
            ctx->last_count.n_pairs = (int64_t)a.n_pairs;
            if (!a.started || a.P_total == 0) { tables_out[k] = make_empty_table(aid_bits); continue; }
            // every group is bucket-sorted already: count the bucket ranges over all of them
            std::vector<const u64*> kp;
            std::vector<const u32*> bp;
            for (int g = 0; g < n_groups; ++g)
                if (a.sorted[g]) { kp.push_back(a.sorted[g]); bp.push_back(a.bounds[g]); }
            const int passes = 1 + a.rest.n;
            try {
                tables_out[k] = hashed_reduce_groups(ctx, kp.data(), bp.data(), (int)kp.size(), (int64_t)a.P_total, a.bb, a.rb, mix,
                                                     a.user_min, a.sym, a.sym, a.ctr.p);
            } catch (const FusedOverflow&) {
                fallback = true;                                      // a hot pair outgrew a region or a table: plain path below
                break;
            }
            cov_trace(ctx, "parts: reduce over all groups + survivors' sort");
            ctx->last_count.sort_passes = passes;
            ctx->last_count.n_chunks = n_groups;
            ctx->last_count.fused = 1;
            ctx->last_count.h2d_bytes = ctx->h2d_bytes_last;
            ctx->last_count.n_unique = tables_out[k]->n;
            for (u64*& b : a.bufs) { dev_free(ctx, b); b = nullptr; }
            for (u32*& b : a.bounds) { dev_free(ctx, b); b = nullptr; }
        }
    }
    if (fallback) {
        // the columns are all on the device by now (or will be: the main stream waits for the last copy)
        for (int k = 0; k < n_specs; ++k)
            if (tables_out[k]) { dev_free(ctx, tables_out[k]->keys); dev_free(ctx, tables_out[k]->count); delete tables_out[k]; tables_out[k] = nullptr; }
        for (auto& a : acc) { for (u64*& b : a.bufs) { dev_free(ctx, b); b = nullptr; } for (u32*& b : a.bounds) { dev_free(ctx, b); b = nullptr; } }
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, e_grp[n_groups - 1], 0));
        if (rle_on)                                     // (re-)materialise every run-length encoded session column
            for (size_t k = 0; k < rle.size(); ++k)
                if (rle[k]->n_runs > 0)
                    COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 4.0 * rle[k]->rows + 8.0 * rle[k]->n_runs, rle_expand_kernel,
                               (unsigned)ceil_div64(rle[k]->n_runs, 256), 256, 0, (const u32*)rle[k]->out, rle[k]->n_runs,
                               rle[k]->rows, d_session.p + task_row0[k]);
        load_events_impl(ctx, d_session.p, d_aid.p, d_ts.p, d_type.p, N, OTTOCOV_DEVICE);
        total = ctx->info;
        for (int k = 0; k < n_specs; ++k) tables_out[k] = count_impl(ctx, &specs[k]);
    }
    free_events(ctx);
    if (trace_on) {
        float ms = 0.f;
        cudaEventSynchronize(tr_b);
        cudaEventElapsedTime(&ms, tr_a, tr_b);
        fprintf(stderr, "[trace] parts: the copy stream was busy for %.3f ms (%.1f GB/s)\n", ms, (double)h2d / ms / 1e6);
        cudaEventDestroy(tr_a); cudaEventDestroy(tr_b);
    }
    ctx->last_count.h2d_bytes = ctx->h2d_bytes_last;
    ctx->info = total;
    ctx->info_only = true;
}

// ---- fused expansion + exchange (include/ottocov.h, "fused expansion + exchange") -----------------------------------
static inline int64_t align16(int64_t v) { return (v + 15) & ~(int64_t)15; }

static PassList xplan_rest_passes(const ottocov_xplan* plan) {
    const int kb = 2 * plan->aid_bits;
    BitField f[1] = {{kb - plan->bucket_bits + plan->sub_bits, kb}};
    return make_pass_list(f, 1);
}

void xplan_make_impl(int n_ranks, int aid_bits, int64_t max_local_keys, int64_t total_keys, int64_t stripe_cap,
                     int64_t mirror_cap, ottocov_xplan* out) {
    if (n_ranks < 2 || n_ranks > XCH_MAX_RANKS) COV_THROW(OTTOCOV_ERR_ARG, "fused exchange needs 2..%d ranks", XCH_MAX_RANKS);
    if (!hashed_reduce_supported(aid_bits) || 2 * aid_bits < 8) COV_THROW(OTTOCOV_ERR_ARG, "fused exchange needs 4 <= aid_bits <= 28");
    if (max_local_keys < 0 || total_keys < 0) COV_THROW(OTTOCOV_ERR_ARG, "negative key count");
    memset(out, 0, sizeof(*out));
    const int kb = 2 * aid_bits;
    // (owner, sub-digit) digits: n_ranks << sub <= max_digits.  Fewer digits = longer per-digit runs per tile (the
    // stores that cross NVLink), more digits = fewer bits left for the owner's remaining passes.
    static int max_digits = 0;
    if (!max_digits) {
        const char* e = getenv("OTTOCOV_XCH_DIGITS");       // tuning knob
        max_digits = e ? atoi(e) : 32;                      // measured at N = 2 (step ms): 128 -> 16.55, 64 -> 16.20, 32 -> 16.11, 16 -> 17.08
        if (max_digits < 2 || max_digits > 256) max_digits = 32;
    }
    int sub = 0;
    while ((n_ranks << (sub + 1)) <= max_digits) ++sub;
    int bb = hashed_bucket_bits(total_keys / n_ranks, kb);  // buckets of the keys one rank receives
    if (bb < sub + 1) bb = sub + 1;                          // at least one pass after the fused one
    if (bb > kb) { bb = kb; if (sub > bb - 1) sub = bb - 1; }
    out->n_ranks = n_ranks; out->aid_bits = aid_bits; out->bucket_bits = bb; out->sub_bits = sub;
    out->rest_passes = xplan_rest_passes(out).n;
    const int64_t n_sub = (int64_t)1 << sub;
    if (stripe_cap <= 0) stripe_cap = (int64_t)((double)max_local_keys / (double)(n_ranks * n_sub) * 1.25) + 2048;
    if (mirror_cap <= 0) { mirror_cap = max_local_keys / (8 * (int64_t)n_ranks); if (mirror_cap < (1 << 16)) mirror_cap = 1 << 16; }
    out->stripe_cap = (stripe_cap + 1) & ~(int64_t)1;
    out->mirror_cap = (mirror_cap + 3) & ~(int64_t)3;
    int64_t o = 0;
    out->off_counts = o;  o = align16(o + (int64_t)n_ranks * n_sub * 8);
    out->off_status = o;  o = align16(o + (int64_t)n_ranks * 4 * 8);
    out->off_hist = o;    o = align16(o + (int64_t)n_ranks * out->rest_passes * RS_RADIX * 8);
    out->off_keys = o;    o = align16(o + (int64_t)n_ranks * n_sub * out->stripe_cap * 8);
    out->off_mstatus = o; o = align16(o + (int64_t)n_ranks * 4 * 8);
    out->off_mkeys = o;   o = align16(o + (int64_t)n_ranks * out->mirror_cap * 8);
    out->off_mcnt = o;    o = align16(o + (int64_t)n_ranks * out->mirror_cap * 4);
    out->total_bytes = o;
}

// one block per destination rank: this source's stripe counts, pass histograms and status go to every rank
__global__ void __launch_bounds__(256) publish_scatter_kernel(PeerBases pb, ottocov_xplan plan, int rank,
                                                             const unsigned long long* __restrict__ cursor,
                                                             const u64* __restrict__ ghist, const u32* __restrict__ flags) {
    const int dest = blockIdx.x;
    const int n_sub = 1 << plan.sub_bits;
    const u64 base = pb.p[dest];
    u64* counts = reinterpret_cast<u64*>(base + plan.off_counts) + (size_t)rank * n_sub;
    for (int j = threadIdx.x; j < n_sub; j += blockDim.x) counts[j] = cursor[dest * n_sub + j];
    u64* hist = reinterpret_cast<u64*>(base + plan.off_hist) + (size_t)rank * plan.rest_passes * RS_RADIX;
    const u64* mine = ghist + (size_t)dest * plan.rest_passes * RS_RADIX;
    for (int j = threadIdx.x; j < plan.rest_passes * RS_RADIX; j += blockDim.x) hist[j] = mine[j];
    if (threadIdx.x == 0) {
        u64 need = 0;
        for (int d = 0; d < plan.n_ranks * n_sub; ++d) need = cursor[d] > need ? cursor[d] : need;
        u64* st = reinterpret_cast<u64*>(base + plan.off_status) + (size_t)rank * 4;
        st[0] = (u64)(*flags);
        st[1] = need;
        st[2] = 0; st[3] = 0;
    }
}

void expand_scatter_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const u64* peer_base_host) {
    ExpandPlan* pl = static_cast<ExpandPlan*>(ctx->plan);
    if (!pl) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_expand_scatter before ottocov_expand_prepare");
    struct PlanFree { ottocov_ctx* c; ~PlanFree() { free_plan(c); } } pf{ctx};
    const int R = plan->n_ranks;
    if (R < 2 || R > XCH_MAX_RANKS || rank < 0 || rank >= R) COV_THROW(OTTOCOV_ERR_ARG, "bad rank / n_ranks");
    if (pl->aid_bits > plan->aid_bits) COV_THROW(OTTOCOV_ERR_ARG, "local aids need %d bits, the plan has %d", pl->aid_bits, plan->aid_bits);
    const KeyMix mix = make_key_mix(plan->aid_bits);
    const PassList rest = xplan_rest_passes(plan);
    if (rest.n != plan->rest_passes || rest.n < 1) COV_THROW(OTTOCOV_ERR_ARG, "inconsistent exchange plan");
    const int n_sub = 1 << plan->sub_bits;
    const u32 n_digits = (u32)R * n_sub;
    DevBuf<u64> reg(ctx, 2 * RS_RADIX);
    DevBuf<unsigned long long> cursor(ctx, RS_RADIX + 2);               // + [flags word]
    DevBuf<u64> ghist(ctx, (size_t)R * rest.n * RS_RADIX);
    u64 h[2 * RS_RADIX];
    memset(h, 0, sizeof(h));
    PeerBases pb;
    memset(&pb, 0, sizeof(pb));
    for (int d = 0; d < R; ++d) {
        pb.p[d] = peer_base_host[d];
        for (int sb = 0; sb < n_sub; ++sb) {
            h[d * n_sub + sb] = peer_base_host[d] + (u64)plan->off_keys + ((u64)(rank * n_sub + sb) * (u64)plan->stripe_cap) * 8ull;
            h[RS_RADIX + d * n_sub + sb] = (u64)plan->stripe_cap;
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(reg.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));    // pageable: staged before return
    CUDA_CHECK(cudaMemsetAsync(cursor.p, 0, (RS_RADIX + 2) * sizeof(unsigned long long), ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(ghist.p, 0, (size_t)R * rest.n * RS_RADIX * sizeof(u64), ctx->stream));
    ScatterArgs sa;
    sa.reg_base = reg.p; sa.reg_cap = reg.p + RS_RADIX; sa.cursor = cursor.p;
    sa.flags = reinterpret_cast<u32*>(cursor.p + RS_RADIX);
    sa.sh1 = mix.kb - plan->bucket_bits; sa.sub_bits = plan->sub_bits; sa.n_digits = n_digits; sa.d_lo = 0; sa.d_hi = n_digits;
    sa.n_dest = (u32)R;
    if (pl->P > 0) expand_scatter_all(ctx, pl, mix, rest, ghist.p, sa);
    COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 0, publish_scatter_kernel, R, 256, 0, pb, *plan, rank, cursor.p, ghist.p, sa.flags);
}

__global__ void __launch_bounds__(256) sum_hist_kernel(const u64* __restrict__ per_src, int n_src, int n, u64* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    u64 s = 0;
    for (int r = 0; r < n_src; ++r) s += per_src[(size_t)r * n + j];
    out[j] = s;
}

__global__ void __launch_bounds__(256) stripe_off_kernel(u64* off, int n, u64 cap) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) off[j] = (u64)j * cap;
}

ottocov_table* reduce_received_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, u64 recv_area, u32 min_count, int sym,
                                    int64_t* need_cap) {
    const int R = plan->n_ranks;
    const int n_sub = 1 << plan->sub_bits;
    *need_cap = 0;
    // what every source published: [R][n_sub] stripe counts, then [R][4] status words (contiguous in the layout)
    const size_t words = (size_t)(plan->off_status - plan->off_counts) / 8 + (size_t)R * 4;
    if (words * 8 > 4096) COV_THROW(OTTOCOV_ERR_ARG, "exchange header larger than the read-back pad");
    u64 hdr[512];
    cov_readback(ctx, hdr, reinterpret_cast<const void*>(recv_area + plan->off_counts), words * 8);
    const u64* st = hdr + (size_t)(plan->off_status - plan->off_counts) / 8;
    u64 need = 0; bool over = false;
    for (int r = 0; r < R; ++r) {
        if (st[r * 4 + 0] & HR_FLAG_FUSED_OVERFLOW) over = true;
        if (st[r * 4 + 1] > need) need = st[r * 4 + 1];
    }
    if (over || need > (u64)plan->stripe_cap) { *need_cap = (int64_t)need; return nullptr; }
    int64_t n = 0;
    for (int j = 0; j < R * n_sub; ++j) n += (int64_t)hdr[j];
    const KeyMix mix = make_key_mix(plan->aid_bits);
    ottocov_table* out = nullptr;
    memset(&ctx->last_count, 0, sizeof(ctx->last_count));
    if (n == 0) { out = new ottocov_table(); out->aid_bits = plan->aid_bits; return out; }
    const PassList rest = xplan_rest_passes(plan);
    DevBuf<u64> ghist(ctx, (size_t)rest.n * RS_RADIX);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, sum_hist_kernel, (unsigned)ceil_div64(rest.n * RS_RADIX, 256), 256, 0,
               reinterpret_cast<const u64*>(recv_area + plan->off_hist), R, rest.n * RS_RADIX, ghist.p);
    DevBuf<u64> seg_off(ctx, (size_t)R * n_sub);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, stripe_off_kernel, (unsigned)ceil_div64(R * n_sub, 256), 256, 0, seg_off.p, R * n_sub,
               (u64)plan->stripe_cap);
    DevBuf<u64> alt(ctx, (size_t)n);
    HashPre pre;
    pre.bb = plan->bucket_bits; pre.first_bits = plan->sub_bits;
    pre.seg_cnt = reinterpret_cast<const u64*>(recv_area + plan->off_counts); pre.seg_off = seg_off.p;
    pre.n_a = R; pre.n_b = n_sub; pre.ctr = nullptr;
    pre.skip_sort = sym != 0;          // a symmetric kind's half table goes through ottocov_mirror_collect, which sorts
    int passes = 0;
    out = hashed_reduce(ctx, reinterpret_cast<u64*>(recv_area + plan->off_keys), alt.p, n, mix, min_count > 1 ? min_count : 1,
                        sym != 0, false, &passes, ghist.p, &pre);
    ctx->last_count.sort_passes = passes;
    ctx->last_count.n_chunks = 1;
    ctx->last_count.fused = 1;
    ctx->last_count.n_unique = out->n;
    return out;
}

// ---- multi-GPU building blocks: expand raw keys grouped by destination rank; reduce received keys --------
void expand_prepare_impl(ottocov_ctx* ctx, const ottocov_spec* spec, int64_t* n_keys, int* symmetric) {
    free_plan(ctx);
    ExpandPlan* pl = make_plan(ctx, spec, true);
    ctx->plan = pl;
    memset(&ctx->last_count, 0, sizeof(ctx->last_count));
    ctx->last_count.n_pairs = (int64_t)(pl->sym ? 2 * pl->P : pl->P);
    *n_keys = (int64_t)pl->P;
    *symmetric = pl->sym ? 1 : 0;
}

void expand_run_impl(ottocov_ctx* ctx, int n_ranks, u64* buf_a, u64* buf_b, int* result_in_b, int64_t* rows_per_dest) {
    ExpandPlan* pl = static_cast<ExpandPlan*>(ctx->plan);
    if (!pl) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_expand_run before ottocov_expand_prepare");
    if (n_ranks < 1 || n_ranks > 256) COV_THROW(OTTOCOV_ERR_ARG, "n_ranks must be 1..256");
    if (n_ranks > 1 && pl->aid_bits > 24) COV_THROW(OTTOCOV_ERR_ARG, "distributed expand needs aid < 2^24 (key bits 56..63 carry the destination)");
    struct PlanFree { ottocov_ctx* c; ~PlanFree() { free_plan(c); } } pf{ctx};
    *result_in_b = 0;
    for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = 0;
    const int64_t n = (int64_t)pl->P;
    if (n == 0) return;
    if (n_ranks == 1) {
        expand_range(ctx, pl, 0, pl->P, buf_a, 0u);
        rows_per_dest[0] = n;
        return;
    }
    // rows per destination = histogram of the stamp bits, accumulated by the expansion itself
    int bits = 1;
    while ((1 << bits) < n_ranks) ++bits;
    BitField dest_field[1] = {{56, 56 + bits}};
    const PassList dp = make_pass_list(dest_field, 1);
    DevBuf<u64> cnt(ctx, RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, RS_RADIX * sizeof(u64), ctx->stream));
    expand_range(ctx, pl, 0, pl->P, buf_a, (u32)n_ranks, nullptr, &dp, cnt.p);
    unsigned long long h[RS_RADIX];
    cov_readback(ctx, h, cnt.p, RS_RADIX * sizeof(u64));
    if (!buf_b) {                    // caller will push the keys itself (ottocov_push_keys): no local grouping
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = (int64_t)h[r];
        return;
    }
    u64* k = buf_a; u64* ka = buf_b; u32* v = nullptr; u32* va = nullptr;
    radix_sort_pairs(ctx, k, ka, v, va, n, dest_field, 1, cnt.p);      // the counts double as the pass histogram
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = (int64_t)h[r];
    *result_in_b = (k == buf_b) ? 1 : 0;
}

void push_keys_impl(ottocov_ctx* ctx, const u64* keys, int64_t n, int n_ranks, const u64* dest_ptrs_host) {
    if (n_ranks < 2 || n_ranks > 256) COV_THROW(OTTOCOV_ERR_ARG, "n_ranks must be 2..256");
    int bits = 1;
    while ((1 << bits) < n_ranks) ++bits;
    radix_partition_push(ctx, keys, n, 56, bits, dest_ptrs_host, n_ranks);
}

__global__ void __launch_bounds__(256) strip_dest_kernel(u64* __restrict__ keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] &= 0x00FFFFFFFFFFFFFFull;
}

ottocov_table* reduce_pairs_impl(ottocov_ctx* ctx, u64* keys, int64_t n, int aid_bits, u32 min_count, int sym,
                                 int strip_dest) {
    if (aid_bits < 1 || aid_bits > 32) COV_THROW(OTTOCOV_ERR_ARG, "aid_bits must be 1..32");
    ottocov_table* out = make_empty_table(aid_bits);
    if (n == 0) return out;
    const bool strip = (strip_dest & 1) != 0;
    const bool hashed = hashed_reduce_supported(aid_bits) && !(strip_dest & OTTOCOV_REDUCE_HASH_OFF) &&
                        ((strip_dest & OTTOCOV_REDUCE_HASH_ON) || min_count > 1);
    if (hashed) {
        delete out;
        const KeyMix mix = make_key_mix(aid_bits);
        const int bb = hashed_bucket_bits(n, mix.kb);
        BitField bucket_field[1] = {{mix.kb - bb, mix.kb}};
        const PassList hp = make_pass_list(bucket_field, 1);
        DevBuf<u64> ghist(ctx, (size_t)(hp.n > 0 ? hp.n : 1) * RS_RADIX);
        mix_keys_inplace(ctx, keys, n, mix, strip, hp, ghist.p);     // the strip pass, now also mixing + histograms
        DevBuf<u64> alt(ctx, n);
        int passes = 0;
        out = hashed_reduce(ctx, keys, alt.p, n, mix, min_count > 1 ? min_count : 1, sym != 0, false, &passes,
                            (hp.n > 0 && n > 1) ? ghist.p : nullptr);
        ctx->last_count.sort_passes = passes;
        ctx->last_count.n_chunks = 1;
        ctx->last_count.n_unique = out->n;
        return out;
    }
    try {
        if (strip)
            COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 16.0 * n, strip_dest_kernel, (unsigned)ceil_div64(n, 256), 256, 0, keys, n);
        DevBuf<u64> alt(ctx, n);
        BitField fields[2] = {{0, aid_bits}, {32, 32 + aid_bits}};
        u64* k = keys; u64* ka = alt.p; u32* v = nullptr; u32* va = nullptr;
        ctx->last_count.sort_passes = radix_sort_pairs(ctx, k, ka, v, va, n, fields, 2);
        reduce_sorted(ctx, k, nullptr, n, min_count > 1 ? min_count : 1, sym != 0, &out->keys, &out->count, &out->n);
        ctx->last_count.n_chunks = 1;
        ctx->last_count.n_unique = out->n;
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

// =====================================================================================================================
// EXTENSION: time-decay weighted co-event scores (include/ottocov.h, "EXTENSION"; no reference counterpart)
// =====================================================================================================================
// Built from the engine's existing pieces, favouring exactness over speed: the window records of the integer path, a
// plain expansion that also evaluates the weight of every pair, the (key, index) radix sort, three run-length reduces
// (count, low and high half of the 24-bit fixed-point weights) and one compaction.
constexpr int WQ_FRAC = 24;
constexpr u32 WQ_FLOOR = 1677722u;                 // round(0.10 * 2^24)
constexpr u32 WQ_MAX_PAIRS_PER_KEY = 1u << 20;     // 2^20 * 4096 < 2^32: the 12-bit halves are summed in 32 bits

__device__ __forceinline__ u32 decay_weight_q(u32 adt, u32 W) {
    if (W == 0) return 1u << WQ_FRAC;
    const u64 one = 1ull << WQ_FRAC;
    const u64 r = (((u64)adt << WQ_FRAC) + (W >> 1)) / W;          // round(|dt| / W * 2^24)
    const u32 q = r >= one ? 0u : (u32)(one - r);
    return q > WQ_FLOOR ? q : WQ_FLOOR;
}

// one thread per output pair of the tile; MODE as in make_tile_keys (never EXM_CANON: weights are kept per ordered pair)
template <int MODE>
__global__ void __launch_bounds__(256) expand_weighted_kernel(const u32* __restrict__ rec_src, const u32* __restrict__ rec_lo,
                                                              const u64* __restrict__ rec_off, const u32* __restrict__ tile_rec,
                                                              const u32* __restrict__ aid_own, const u32* __restrict__ aid_rng,
                                                              const u64* __restrict__ skey_own, const u64* __restrict__ skey_rng,
                                                              u64 n_out_total, u32 W, u64* __restrict__ keys,
                                                              u32* __restrict__ wq, u32* __restrict__ idx, u64 out_base) {
    const u64 tile = blockIdx.x;
    const u64 o0 = tile * EX_TILE;
    const u32 r0 = tile_rec[tile];
    u32 r1 = tile_rec[tile + 1];
    for (int q = 0; q < EX_TILE / 256; ++q) {
        const u64 o = o0 + (u64)q * 256 + threadIdx.x;
        if (o >= n_out_total) return;
        u32 lo = r0, hi = r1 + 1;                            // last record with rec_off <= o (records r0 .. r1 may own it)
        while (hi - lo > 1) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (rec_off[mid] <= o) lo = mid; else hi = mid;
        }
        const u32 src = rec_src[lo];
        u32 tgt = rec_lo[lo] + (u32)(o - rec_off[lo]);
        if (MODE == EXM_SELF) tgt += (tgt >= src);
        const u32 a = aid_own[src], b = aid_rng[tgt];
        const u32 t0 = (u32)skey_own[src], t1 = (u32)skey_rng[tgt];
        const u32 adt = t1 > t0 ? t1 - t0 : t0 - t1;
        keys[out_base + o] = (MODE == EXM_SWAP) ? (((u64)b << 32) | a) : (((u64)a << 32) | b);
        wq[out_base + o] = decay_weight_q(adt, W);
        idx[out_base + o] = (u32)(out_base + o);
    }
}

__global__ void __launch_bounds__(256) wq_gather_split_kernel(const u32* __restrict__ idx, const u32* __restrict__ wq, int64_t n,
                                                              u32* __restrict__ lo, u32* __restrict__ hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 w = wq[idx[i]];
    lo[i] = w & 0xFFFu;
    hi[i] = w >> 12;
}

// count / low sums / high sums of the same distinct keys -> rows with count >= min_count
struct WeightedRows {
    static constexpr int NC = 1;
    const u64* keys; const u32* cnt; const u32* slo; const u32* shi;
    u32 min_count;
    u64* o_keys; u32* o_cnt; u64* o_score; u32* flag;
    __device__ u64 value(int64_t i) const { return cnt[i] >= min_count ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (cnt[i] > WQ_MAX_PAIRS_PER_KEY) atomicOr(flag, 1u);
        if (!v) return;
        o_keys[pre[0]] = keys[i];
        o_cnt[pre[0]] = cnt[i];
        o_score[pre[0]] = ((u64)shi[i] << 12) + (u64)slo[i];
    }
};

static void wtable_release(ottocov_ctx* ctx, ottocov_wtable* t) {
    if (!t) return;
    dev_free(ctx, t->keys); dev_free(ctx, t->count); dev_free(ctx, t->score_fx);
    delete t;
}

ottocov_wtable* count_weighted_impl(ottocov_ctx* ctx, const ottocov_spec* spec) {
    ottocov_spec sp = *spec;
    sp.flags = (sp.flags & ~(u32)OTTOCOV_SYM_ON) | OTTOCOV_SYM_OFF;          // every ordered pair carries its own weight
    sp.min_count = 1;
    ExpandPlan* pl = make_plan(ctx, &sp, false);
    struct PlanGuard { ExpandPlan* p; ~PlanGuard() { delete p; } } plan_guard{pl};
    memset(&ctx->last_count, 0, sizeof(ctx->last_count));
    ctx->last_count.n_pairs = (int64_t)pl->P;
    ottocov_wtable* out = new ottocov_wtable();
    out->aid_bits = pl->aid_bits;
    const int64_t P = (int64_t)pl->P;
    if (P == 0) return out;
    try {
        if (P >= (int64_t)0xFFFFFFFFll) COV_THROW(OTTOCOV_ERR_CAPACITY, "weighted mode: at most 2^32-2 pairs per call (got %lld)", (long long)P);
        const u32 W = (u32)(spec->window > 0x7FFFFFFFll ? 0x7FFFFFFFll : (spec->window < 0 ? 0 : spec->window));
        DevBuf<u64> keys(ctx, P), kalt(ctx, P);
        DevBuf<u32> idx(ctx, P), ialt(ctx, P), wq(ctx, P);
        u64 base = 0;
        for (Segment* sg : pl->segs) {
            if (sg->n_pairs == 0) continue;
            const int64_t n_tiles = ceil_div64((int64_t)sg->n_pairs, EX_TILE);
            DevBuf<u32> tile_rec(ctx, n_tiles + 1);
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, 0, tile_search_kernel, (unsigned)ceil_div64(n_tiles + 1, 256), 256, 0,
                       sg->rec_off.p, (int64_t)sg->n_rec, (u64)0, (u64)sg->n_pairs, n_tiles, tile_rec.p);
            const TypeArray& own = ctx->ta[sg->own_type];
            const TypeArray& rng = ctx->ta[sg->tgt_type];
#define EXW_LAUNCH(MODE_)                                                                                               \
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, 16.0 * (double)sg->n_pairs, expand_weighted_kernel<MODE_>, (unsigned)n_tiles, 256, 0, \
                       sg->rec_src.p, sg->rec_lo.p, sg->rec_off.p, tile_rec.p, own.aid, rng.aid, own.skey, rng.skey,         \
                       (u64)sg->n_pairs, W, keys.p, wq.p, idx.p, base)
            if (sg->self) EXW_LAUNCH(EXM_SELF);
            else if (sg->swap) EXW_LAUNCH(EXM_SWAP);
            else EXW_LAUNCH(EXM_CROSS);
#undef EXW_LAUNCH
            base += sg->n_pairs;
        }
        BitField fields[2] = {{0, pl->aid_bits}, {32, 32 + pl->aid_bits}};
        u64* k = keys.p; u64* ka = kalt.p; u32* v = idx.p; u32* va = ialt.p;
        ctx->last_count.sort_passes = radix_sort_pairs(ctx, k, ka, v, va, P, fields, 2);
        DevBuf<u32> lo(ctx, P), hi(ctx, P);
        COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 16.0 * P, wq_gather_split_kernel, (unsigned)ceil_div64(P, 256), 256, 0, v, wq.p, P, lo.p, hi.p);
        u64* uk[3] = {nullptr, nullptr, nullptr};
        u32* uc[3] = {nullptr, nullptr, nullptr};
        int64_t un[3] = {0, 0, 0};
        struct Free3 { ottocov_ctx* c; u64** k; u32** v; ~Free3() { for (int i = 0; i < 3; ++i) { dev_free(c, k[i]); dev_free(c, v[i]); } } } f3{ctx, uk, uc};
        reduce_sorted(ctx, k, nullptr, P, 1, false, &uk[0], &uc[0], &un[0]);     // pair counts
        reduce_sorted(ctx, k, lo.p, P, 0, false, &uk[1], &uc[1], &un[1]);        // sums of the low 12 weight bits (a sum may be 0:
        reduce_sorted(ctx, k, hi.p, P, 0, false, &uk[2], &uc[2], &un[2]);        // threshold 0 keeps every key); sums of the high bits
        if (un[1] != un[0] || un[2] != un[0]) COV_THROW(OTTOCOV_ERR_CUDA, "weighted reduce: row counts disagree");
        const int64_t U = un[0];
        DevBuf<u64> okeys(ctx, U), oscore(ctx, U);
        DevBuf<u32> ocnt(ctx, U), flag(ctx, 1);
        CUDA_CHECK(cudaMemsetAsync(flag.p, 0, 4, ctx->stream));
        WeightedRows f;
        f.keys = uk[0]; f.cnt = uc[0]; f.slo = uc[1]; f.shi = uc[2];
        f.min_count = spec->min_count > 1 ? spec->min_count : 1;
        f.o_keys = okeys.p; f.o_cnt = ocnt.p; f.o_score = oscore.p; f.flag = flag.p;
        u64 tot[1];
        scan_apply(ctx, OTTOCOV_K_FILTER, f, U, tot, 28.0 * U);
        u32 hflag = 0;
        cov_readback(ctx, &hflag, flag.p, 4);
        if (hflag) COV_THROW(OTTOCOV_ERR_CAPACITY, "weighted mode: a pair occurs more than 2^20 times");
        out->n = (int64_t)tot[0];
        out->keys = okeys.take(); out->count = ocnt.take(); out->score_fx = oscore.take();
        ctx->last_count.n_chunks = 1;
        ctx->last_count.n_unique = out->n;
    } catch (...) {
        wtable_release(ctx, out);
        throw;
    }
    return out;
}

__global__ void __launch_bounds__(256) wtable_unpack_kernel(const u64* __restrict__ keys, const u32* __restrict__ cnt,
                                                            const u64* __restrict__ score_fx, const u32* __restrict__ order,
                                                            int64_t n, int32_t* __restrict__ aid, int32_t* __restrict__ aid_next,
                                                            double* __restrict__ score, int32_t* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = order ? (int64_t)order[i] : i;
    const u64 k = keys[r];
    aid[i] = (int32_t)(k >> 32); aid_next[i] = (int32_t)(u32)k;
    score[i] = (double)score_fx[r] / (double)(1u << WQ_FRAC);
    if (count) count[i] = (int32_t)cnt[r];
}

void wtable_fetch_impl(ottocov_ctx* ctx, const ottocov_wtable* t, int32_t* aid, int32_t* aid_next, double* score,
                       int32_t* count, int64_t cap, int where, int64_t* n_out) {
    const int64_t n = t->n;
    if (n_out) *n_out = n;
    if (cap < n) COV_THROW(OTTOCOV_ERR_CAPACITY, "weighted fetch needs room for %lld rows", (long long)n);
    if (n == 0) return;
    DevBuf<int32_t> da, db, dc;
    DevBuf<double> ds;
    int32_t *pa = aid, *pb = aid_next, *pc = count;
    double* ps = score;
    if (where == OTTOCOV_HOST) { da.alloc(ctx, n); db.alloc(ctx, n); dc.alloc(ctx, n); ds.alloc(ctx, n); pa = da.p; pb = db.p; pc = dc.p; ps = ds.p; }
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 40.0 * n, wtable_unpack_kernel, (unsigned)ceil_div64(n, 256), 256, 0, t->keys, t->count, t->score_fx,
               (const u32*)nullptr, n, pa, pb, ps, pc);
    if (where == OTTOCOV_HOST) {
        CUDA_CHECK(cudaMemcpyAsync(aid, pa, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_next, pb, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(score, ps, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (count) CUDA_CHECK(cudaMemcpyAsync(count, pc, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

__global__ void __launch_bounds__(256) wt_inv_score_kernel(const u64* __restrict__ score_fx, int64_t n, u64 mask, u64* __restrict__ sk,
                                                           u32* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sk[i] = (~score_fx[i]) & mask;            // ascending order of this = descending score
    idx[i] = (u32)i;
}

__global__ void __launch_bounds__(256) wt_aid_of_kernel(const u64* __restrict__ keys, const u32* __restrict__ idx, int64_t n,
                                                        u64* __restrict__ ak) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ak[i] = keys[idx[i]] >> 32;
}

// rows in (aid asc, score desc, aid_next asc) order: keep the first k of every aid
struct WeightedTopRows {
    static constexpr int NC = 1;
    const u64* ak;            // aid of row i in that order
    const u32* order;         // row of the table
    int64_t n;
    int k;
    u32* o_order; int32_t* o_rank;
    __device__ int64_t seg_start(int64_t i) const {
        const u64 a = ak[i];
        int64_t lo = 0, hi = i;                              // first row with ak >= a
        while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (ak[mid] < a) lo = mid + 1; else hi = mid; }
        return lo;
    }
    __device__ u64 value(int64_t i) const { return (i - seg_start(i)) < k ? 1ull : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        o_order[pre[0]] = order[i];
        o_rank[pre[0]] = (int32_t)(i - seg_start(i)) + 1;
    }
};

void wtable_topk_impl(ottocov_ctx* ctx, const ottocov_wtable* t, int k, int32_t* aid, int32_t* aid_next, double* score,
                      int32_t* rank, int64_t cap, int where, int64_t* n_out) {
    if (k < 1) COV_THROW(OTTOCOV_ERR_ARG, "k must be >= 1");
    const int64_t n = t->n;
    if (n_out) *n_out = 0;
    if (n == 0) return;
    // stable sort by descending score, then stable sort by aid: (aid asc, score desc, aid_next asc)
    DevBuf<u64> sk(ctx, n), ska(ctx, n);
    DevBuf<u32> idx(ctx, n), ia(ctx, n);
    const int score_bits = WQ_FRAC + 21;                     // <= 2^20 pairs of weight <= 2^24 per key
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 20.0 * n, wt_inv_score_kernel, (unsigned)ceil_div64(n, 256), 256, 0, t->score_fx, n,
               (1ull << score_bits) - 1ull, sk.p, idx.p);
    u64* k1 = sk.p; u64* k1a = ska.p; u32* v1 = idx.p; u32* v1a = ia.p;
    BitField fs[1] = {{0, score_bits}};
    radix_sort_pairs(ctx, k1, k1a, v1, v1a, n, fs, 1);
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 20.0 * n, wt_aid_of_kernel, (unsigned)ceil_div64(n, 256), 256, 0, t->keys, v1, n, k1a);
    u64* k2 = k1a; u64* k2a = k1;                            // the aid keys live in the spare buffer of the first sort
    u32* v2 = v1; u32* v2a = v1a;
    BitField fa[1] = {{0, t->aid_bits}};
    radix_sort_pairs(ctx, k2, k2a, v2, v2a, n, fa, 1);
    DevBuf<u32> o_order(ctx, n);
    DevBuf<int32_t> o_rank(ctx, n);
    WeightedTopRows f;
    f.ak = k2; f.order = v2; f.n = n; f.k = k; f.o_order = o_order.p; f.o_rank = o_rank.p;
    u64 tot[1];
    scan_apply(ctx, OTTOCOV_K_TOPK, f, n, tot, 24.0 * n);
    const int64_t m = (int64_t)tot[0];
    if (n_out) *n_out = m;
    if (cap == 0) return;                                    // size query
    if (cap < m) COV_THROW(OTTOCOV_ERR_CAPACITY, "weighted top-k needs room for %lld rows", (long long)m);
    DevBuf<int32_t> da, db;
    DevBuf<double> ds;
    int32_t *pa = aid, *pb = aid_next;
    double* ps = score;
    if (where == OTTOCOV_HOST) { da.alloc(ctx, m); db.alloc(ctx, m); ds.alloc(ctx, m); pa = da.p; pb = db.p; ps = ds.p; }
    COV_LAUNCH(ctx, OTTOCOV_K_ORDER, 40.0 * m, wtable_unpack_kernel, (unsigned)ceil_div64(m, 256), 256, 0, t->keys, t->count, t->score_fx,
               (const u32*)o_order.p, m, pa, pb, ps, (int32_t*)nullptr);
    const cudaMemcpyKind kind = (where == OTTOCOV_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (where == OTTOCOV_HOST) {
        CUDA_CHECK(cudaMemcpyAsync(aid, pa, m * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_next, pb, m * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(score, ps, m * 8, kind, ctx->stream));
    }
    CUDA_CHECK(cudaMemcpyAsync(rank, o_rank.p, m * 4, kind, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
