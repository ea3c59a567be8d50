// radix_sort.cu -- single-sweep LSD radix sort of 64-bit keys (+ optional 32-bit payload).
//
// This is the engine behind the reference's groupby(['aid','aid_next']) (model/count_co_events.py
// :70-71, :168): co-event pairs are packed into u64 keys, sorted, then run-length reduced.
//
// Shape (per pass: read n keys once, write n keys once = 16 B/key, the HBM floor for a
// distribution pass):
//   * one up-front kernel builds the digit histograms of ALL passes in a single read;
//   * each pass is one kernel.  A CTA takes the next tile (atomic ticket, so a CTA only ever waits
//     for tiles that already started), ranks its 4096 keys by digit (one ballot per digit bit finds
//     the lanes with the same digit; running per-warp counters make the rank stable), publishes the
//     per-digit tile counts, resolves the per-digit exclusive prefix over earlier tiles by decoupled
//     look-back on epoch-tagged 64-bit status words, stages the keys in digit order in shared memory
//     and writes them out as coalesced per-digit runs.  With `ptr_base` the per-digit runs start at
//     caller-given byte addresses instead (peer memory: the fused partition + exchange of dist.py).
//   * status word = flag(2) | epoch(6) | value(56).  The epoch changes every pass, so the status
//     array is never re-zeroed between passes.
// Only the significant bit fields are sorted: for co-event keys that is 2 x aid_bits (42 bits for
// 1.8 M aids -> 6 passes of 7 bits), not 64.
#include "internal.cuh"

#ifndef OTTOCOV_RS_THREADS
#define OTTOCOV_RS_THREADS 256
#endif
constexpr int RS_THREADS = OTTOCOV_RS_THREADS;
// Keys per thread = tile size / 256.  Measured on 742 M-key passes (experiments/README.md): 2.0 TB/s at 8, 2.66 at
// 16, 2.92 at 32, 2.97 at 40 -- the per-tile costs (look-back round trip, scans over the warp histograms, five CTA
// barriers) are amortised over more keys, and that outweighs the lower occupancy (128 registers, 2 CTAs/SM).
// A (key, value) tile carries half again as many registers per item, so it stays smaller.
#ifndef OTTOCOV_RS_IPT
#define OTTOCOV_RS_IPT 40
#endif
#ifndef OTTOCOV_RS_IPT_PAIRS
#define OTTOCOV_RS_IPT_PAIRS 16
#endif
constexpr int RS_IPT_KEYS = OTTOCOV_RS_IPT;
constexpr int RS_IPT_PAIRS = OTTOCOV_RS_IPT_PAIRS;
constexpr int rs_ipt(bool has_vals) { return has_vals ? RS_IPT_PAIRS : RS_IPT_KEYS; }
constexpr int rs_tile(bool has_vals) { return RS_THREADS * rs_ipt(has_vals); }
constexpr int RS_WARPS = RS_THREADS / 32;

constexpr u64 ST_VALUE_MASK = (1ull << 56) - 1;
constexpr u64 ST_FLAG_AGG = 1ull << 62;
constexpr u64 ST_FLAG_PREFIX = 2ull << 62;


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
// Order of work inside a CTA (one tile of 4096 keys):
//   A  load keys into registers (16 independent 8-byte loads per thread in flight)
//   B  stable rank of each key inside its warp's segment (ballot matching + running per-warp digit
//      counters); the per-warp digit histograms fall out of the same pass
//   C  per digit: offsets over warps, tile total, position of the digit inside the tile; the tile
//      total is published at once (flag AGG)
//   E  look-back RIGHT AWAY (before the long ranking phase, so the gap between a tile's AGG and its
//      PREFIX is one short walk and later tiles almost always find a PREFIX next door): thread d
//      walks status[tile-1][d], status[tile-2][d], ... RS_WINDOW entries per round trip, sums AGGs
//      until the first PREFIX, publishes its own PREFIX
//   D  keys go to their slot of the digit-ordered staging buffer (the warps without look-back duty do
//      this while the first `radix` threads are still walking)
//   F  staged keys leave as coalesced per-digit runs
#ifndef OTTOCOV_RS_WINDOW
#define OTTOCOV_RS_WINDOW 4
#endif
constexpr int RS_WINDOW = OTTOCOV_RS_WINDOW;

// lanes of the warp whose digit equals this lane's: one ballot per digit bit, 4 SASS instructions per bit
// (test bit -> predicate, VOTE, predicated NOT, AND).  Bits at and above NB are zero in every lane.
// Not `asm volatile`: the ballots of different rows are independent, and the scheduler may interleave them.
template <int NB>
__device__ __forceinline__ u32 match_digit_ballot(u32 d) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        u32 m;
        asm(
            "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
            "@!p not.b32 %0, %0;\n\t}"
            : "=r"(m)
            : "r"(d), "r"(1u << b));
        peers &= m;
    }
    return peers;
}

// Stable rank of each of the lane's IPT keys inside the warp's 32 x IPT-key segment (rows of 32 keys);
// bumps the running digit counters.  FULL = every row is valid (all tiles but the last).
// The running counter of a digit is a read-modify-write in shared memory, so consecutive rows of one counter
// array form a serial latency chain (ballots -> leader's load -> store -> shuffle), and with 2 CTAs of 8 warps
// per SM nothing else hides it (ncu, round 1: 0.57 eligible warps per scheduler).  SPLIT cuts the warp's rows
// into SPLIT groups of consecutive rows, each with its own counter array ("virtual warps" warp * SPLIT + g):
// the SPLIT chains are independent and are issued interleaved, phase by phase.
template <int NB, bool FULL, int IPT, int SPLIT>
__device__ __forceinline__ void rank_rows(const u64 (&key)[IPT], int shift, u32 mask, int my_n, int lane,
                                          u32 lt, u32* my_whist, u32 (&rnk2)[IPT / 2]) {
    constexpr int G = IPT / SPLIT;               // rows per group
    constexpr int STRIDE = 1 << NB;              // counters per group
    static_assert(IPT % SPLIT == 0 && G % 2 == 0, "row groups must pair up for the 16-bit rank packing");
#pragma unroll
    for (int j = 0; j < G; ++j) {
        u32 d[SPLIT], peers[SPLIT], old[SPLIT];
        int leader[SPLIT];
        bool valid[SPLIT];
#pragma unroll
        for (int g = 0; g < SPLIT; ++g) {
            const int i = g * G + j;
            valid[g] = FULL || i < my_n;
            d[g] = (u32)(key[i] >> shift) & mask;
            peers[g] = match_digit_ballot<NB>(d[g]);
            if (!FULL) {
                const u32 vm = __ballot_sync(0xffffffffu, valid[g]);
                peers[g] &= valid[g] ? vm : ~vm;
            }
            leader[g] = __ffs(peers[g]) - 1;
        }
#pragma unroll
        for (int g = 0; g < SPLIT; ++g) {
            old[g] = 0;
            if (lane == leader[g] && valid[g]) old[g] = my_whist[g * STRIDE + d[g]];
        }
#pragma unroll
        for (int g = 0; g < SPLIT; ++g)
            if (lane == leader[g] && valid[g]) my_whist[g * STRIDE + d[g]] = old[g] + __popc(peers[g]);
#pragma unroll
        for (int g = 0; g < SPLIT; ++g) {
            const int i = g * G + j;
            const u32 r = __shfl_sync(0xffffffffu, old[g], leader[g]) + __popc(peers[g] & lt);
            if (i & 1) rnk2[i >> 1] |= r << 16; else rnk2[i >> 1] = r;
        }
        __syncwarp();
    }
}

template <bool HAS_VALS, bool FULL, int IPT>
__device__ __forceinline__ void load_rows(const u64* __restrict__ keys_in, const u32* __restrict__ vals_in, int64_t lbase,
                                          int my_n, u64 (&key)[IPT], u32 (&val)[IPT]) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const bool valid = FULL || i < my_n;
        key[i] = valid ? ld_stream_u64(keys_in + lbase + i * 32) : ~0ull;
        if (HAS_VALS) val[i] = valid ? __ldcs(vals_in + lbase + i * 32) : 0u;
    }
}

template <bool HAS_VALS, bool FULL, int IPT, int SPLIT, int NB>
__device__ __forceinline__ void stage_rows(const u64 (&key)[IPT], const u32 (&val)[IPT], const u32 (&rnk2)[IPT / 2],
                                           int shift, u32 mask, int my_n, const u32* my_whist, u64* s_keys, u32* s_vals) {
    constexpr int G = IPT / SPLIT;
    constexpr int STRIDE = 1 << NB;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        if (FULL || i < my_n) {
            const u32 d = (u32)(key[i] >> shift) & mask;
            const u32 r = (i & 1) ? (rnk2[i >> 1] >> 16) : (rnk2[i >> 1] & 0xFFFFu);
            const u32 pos = my_whist[(i / G) * STRIDE + d] + r;
            s_keys[pos] = key[i];
            if (HAS_VALS) s_vals[pos] = val[i];
        }
    }
}

// s_gptr[d] = byte address of keys_out[global slot of the tile's first key of digit d] - 8 * (slot of that key in
// the staging buffer), so the key staged at slot j goes to s_gptr[d] + 8 j: one 64-bit add per key.
template <bool HAS_VALS, bool FULL, int IPT>
__device__ __forceinline__ void write_rows(const u64* s_keys, const u32* s_vals, const u64* s_gptr, int shift, u32 mask,
                                           int tid, int tile_n, const u64* keys_out, u32* vals_out) {
    const u64 toff = (u64)tid * 8u;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int j = tid + i * RS_THREADS;
        if (FULL || j < tile_n) {
            const u64 k = s_keys[j];
            const u32 d = (u32)(k >> shift) & mask;
            const u64 a = s_gptr[d] + toff + (u64)(i * RS_THREADS * 8);
            __stcs(reinterpret_cast<u64*>(a), k);
            if (HAS_VALS) {
                const u64 g = (a - reinterpret_cast<u64>(keys_out)) >> 3;
                vals_out[g] = s_vals[j];
            }
        }
    }
}

// ---- segmented input ------------------------------------------------------------------------------------
// The first distribution pass of the bucketed hash reduce is fused into the pair expansion (expand.cu): it needs
// no stability, so the expansion scatters its keys straight into one region per digit, reserving room with one
// atomic per (tile, digit).  The regions have slack, so the input of the NEXT pass is a list of segments
// (offset, count) instead of one array.  The logical input is their concatenation in table order.
// Table (device): n_segs, total tiles, then tile_start[n_segs + 1], in_off[n_segs], cnt[n_segs].
struct RsSegs {
    u32 n_segs;
    u32 total_tiles;
    u64 total_keys;
    const u32* tile_start;     // [n_segs + 1] first tile of each segment (tiles never straddle segments)
    const u64* in_off;         // [n_segs] offset of the segment in keys_in
    const u64* cnt;            // [n_segs] keys in the segment
    const u64* tile_in;        // [launch grid] per tile: offset of its first key in keys_in ...
    const u32* tile_n;         // [launch grid] ... and its key count (0 for tickets beyond the last tile)
    const u32* tile_first;     // [launch grid] segment index if the tile is the first of its segment, else ~0
    // Bucket bounds (optional; only when this pass is the LAST one).  Segment s = b * n_a + a holds the keys whose
    // earlier digit is b, so after this pass the keys with (this digit, earlier digit) = (d, b) are contiguous, and
    // they start where digit d starts + what the tiles of the segments before s hold of digit d -- exactly the
    // exclusive prefix the first tile of segment s resolves by look-back.  It records
    // bounds[d * n_b + b] = min(.., that position); entries of empty buckets stay ~0 (hash_reduce.cu fills them in).
    u64* bounds;
    u32 n_a, n_b;
};

// Per-tile table of a segmented launch (one thread per ticket of the launch grid): the pass kernel then needs one
// load per tile instead of a search over the segment table on its critical path.
__global__ void __launch_bounds__(256) rs_seg_tiles_kernel(const RsSegs* __restrict__ segs, u32 tile_keys, u32 grid_tiles,
                                                           u64* __restrict__ tile_in, u32* __restrict__ tile_n,
                                                           u32* __restrict__ tile_first) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= grid_tiles) return;
    if (t >= segs->total_tiles) { tile_in[t] = 0; tile_n[t] = 0; tile_first[t] = ~0u; return; }
    const u32* ts = segs->tile_start;
    u32 lo = 0, hi = segs->n_segs;                 // last segment whose first tile is <= t
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (ts[mid] <= t) lo = mid; else hi = mid;
    }
    const u64 within = (u64)(t - ts[lo]) * tile_keys;
    const u64 left = segs->cnt[lo] - within;
    tile_in[t] = segs->in_off[lo] + within;
    tile_n[t] = (u32)(left < tile_keys ? left : tile_keys);
    tile_first[t] = (t == ts[lo]) ? lo : ~0u;
}

// One block.  Segment s = b * n_a + a (b-major: b is the digit the keys were partitioned on, a the writer
// that filled the region) lives at slot a * n_b + b: key offset off_src[slot], count cnt_src[slot].
__global__ void __launch_bounds__(256) rs_seg_build_kernel(const u64* __restrict__ cnt_src, const u64* __restrict__ off_src,
                                                           int n_a, int n_b,
                                                           u32 tile_keys, RsSegs* hdr, u32* tile_start, u64* in_off, u64* cnt) {
    __shared__ u64 s_warp[256 / 32 + 1];
    __shared__ u64 s_carry[2];
    const int S = n_a * n_b;
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (int base = 0; base < S; base += 256) {
        const int s = base + threadIdx.x;
        u64 c = 0, slot = 0;
        if (s < S) {
            const int b = s / n_a, a = s % n_a;
            slot = (u64)a * n_b + b;
            c = cnt_src[slot];
            in_off[s] = off_src[slot];
            cnt[s] = c;
        }
        const u64 t = (c + tile_keys - 1) / tile_keys;
        u64 tot;
        const u64 ext = block_exclusive_scan<u64, 256>(t, s_warp, &tot);
        const u64 c0 = s_carry[0];
        if (s < S) tile_start[s] = (u32)(c0 + ext);
        u64 ktot;
        block_exclusive_scan<u64, 256>(c, s_warp, &ktot);
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] = c0 + tot; s_carry[1] += ktot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_start[S] = (u32)s_carry[0];
        hdr->n_segs = (u32)S;
        hdr->total_tiles = (u32)s_carry[0];
        hdr->total_keys = s_carry[1];
        hdr->tile_start = tile_start;
        hdr->in_off = in_off;
        hdr->cnt = cnt;
    }
}

__global__ void rs_seg_attach_kernel(RsSegs* hdr, const u64* tile_in, const u32* tile_n, const u32* tile_first, u64* bounds,
                                     u32 n_a, u32 n_b) {
    hdr->tile_in = tile_in;
    hdr->tile_n = tile_n;
    hdr->tile_first = tile_first;
    hdr->bounds = bounds;
    hdr->n_a = n_a;
    hdr->n_b = n_b;
}

#ifndef OTTOCOV_RS_MINB
#define OTTOCOV_RS_MINB 2          // keys only: 40 keys per thread need the 128-register budget of 2 CTAs/SM
#endif
#ifndef OTTOCOV_RS_MINB_PAIRS
#define OTTOCOV_RS_MINB_PAIRS 4    // (key, value) tiles of 4096: 64 registers, 4 CTAs/SM
#endif
#ifndef OTTOCOV_RS_SPLIT
#define OTTOCOV_RS_SPLIT 2         // independent ranking chains per warp (rank_rows)
#endif
constexpr int RS_SPLIT = OTTOCOV_RS_SPLIT;

template <bool HAS_VALS, int NB>
__global__ void __launch_bounds__(RS_THREADS, HAS_VALS ? OTTOCOV_RS_MINB_PAIRS : OTTOCOV_RS_MINB)
rs_onesweep_kernel(const u64* __restrict__ keys_in, u64* __restrict__ keys_out,
                   const u32* __restrict__ vals_in, u32* __restrict__ vals_out, int64_t n, int shift,
                   int bits, const u64* __restrict__ digit_base, const u64* __restrict__ ptr_base, u64* status,
                   u32* ticket, u32 epoch, const RsSegs* __restrict__ segs, const u32* __restrict__ abort_flag) {
    // The writer of a segmented input (the fused expansion) ran out of room in a region: the input is incomplete and
    // the host is about to repeat the step.  Every CTA of every later pass leaves at once (uniform branch), so nothing
    // is ranked against histograms that do not match the data.
    if (abort_flag != nullptr && (*abort_flag & HR_FLAG_FUSED_OVERFLOW)) return;
    constexpr int IPT = HAS_VALS ? RS_IPT_PAIRS : RS_IPT_KEYS;
    constexpr int RS_TILE = RS_THREADS * IPT;
    constexpr int SPLIT = RS_SPLIT;
    constexpr int STRIDE = 1 << NB;
    constexpr int VW = RS_WARPS * SPLIT;                                      // virtual warps = counter arrays
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_keys = reinterpret_cast<u64*>(s_raw);                              // [RS_TILE]
    u32* s_whist = reinterpret_cast<u32*>(s_keys + RS_TILE);                  // [VW][STRIDE]
    u32* s_dstart = s_whist + VW * STRIDE;                                    // [RS_RADIX]
    u64* s_gptr = reinterpret_cast<u64*>(s_dstart + RS_RADIX);                // [RS_RADIX]
    u32* s_scan = reinterpret_cast<u32*>(s_gptr + RS_RADIX);                  // [RS_WARPS + 1] (+pad to 48)
    u32* s_tile = s_scan + 48;                                                // [8]: tile, tile_n, input offset (u64)
    u32* s_vals = s_tile + 16;                                                // [RS_TILE] if HAS_VALS

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 radix = 1u << bits, mask = radix - 1u;

    if (warp == 0) {
        u32 t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        u64 in_base = (u64)t * RS_TILE;
        int64_t left = n - (int64_t)in_base;
        u32 first_of = ~0u;
        if (segs) {
            in_base = segs->tile_in[t];
            left = (int64_t)segs->tile_n[t];
            if (segs->bounds) first_of = segs->tile_first[t];
        }
        if (lane == 0) {
            s_tile[0] = t;
            s_tile[1] = (u32)(left <= 0 ? 0 : (left < RS_TILE ? left : RS_TILE));
            *reinterpret_cast<u64*>(s_tile + 2) = in_base;
            s_tile[4] = first_of;
        }
    }
    for (int j = tid; j < VW * STRIDE; j += RS_THREADS) s_whist[j] = 0;
    __syncthreads();
    const int64_t tile = s_tile[0];
    const int tile_n = (int)s_tile[1];
    if (tile_n == 0) return;                                 // a ticket beyond the last tile (segmented launches are
    const u64 in_base = *reinterpret_cast<u64*>(s_tile + 2); // sized by an upper bound): never published, never awaited
    const bool full = tile_n == RS_TILE;

    // -- A: load (warp-striped: every access is a coalesced 256 B row) ------------------------------
    u64 key[IPT];
    u32 val[IPT];
    const int lw = warp * 32 * IPT + lane;                                  // first slot of this lane in the tile
    const int64_t lbase = (int64_t)in_base + lw;
    int my_n = (tile_n - lw + 31) / 32;                                     // valid items of this lane (may be <= 0)
    if (my_n > IPT) my_n = IPT;
    if (full) load_rows<HAS_VALS, true>(keys_in, vals_in, lbase, my_n, key, val);
    else load_rows<HAS_VALS, false>(keys_in, vals_in, lbase, my_n, key, val);

    // -- B: stable rank of every key inside its (virtual) warp's segment; the per-warp digit counts fall
    //       out of the same pass.  Lanes holding the same digit are found with one ballot per digit
    //       bit (match.any is far slower than `bits` ballots on this part); the lowest such lane
    //       bumps the running counter for the digit, no atomics needed.
    u32* my_whist = s_whist + warp * SPLIT * STRIDE;
    const u32 lt = lanemask_lt();
    u32 rnk2[IPT / 2];                       // two 16-bit ranks per register
    if (full) rank_rows<NB, true, IPT, SPLIT>(key, shift, mask, my_n, lane, lt, my_whist, rnk2);
    else rank_rows<NB, false, IPT, SPLIT>(key, shift, mask, my_n, lane, lt, my_whist, rnk2);
    __syncthreads();

    // -- C: offsets over (virtual) warps, tile totals, digit starts; publish the aggregate --------------
    u32 total = 0;
    if (tid < (int)radix) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < VW; ++w) {
            const u32 c = s_whist[w * STRIDE + tid];
            s_whist[w * STRIDE + tid] = run;
            run += c;
        }
        total = run;
    }
    u32 blk_total;
    const u32 dstart = block_exclusive_scan<u32, RS_THREADS>(total, s_scan, &blk_total);
    const u64 tag = (u64)epoch << 56;
    u64* mine = status + (size_t)tile * radix + tid;          // [tile][digit]: a tile's words are contiguous
    if (tid < (int)radix) {
        st_volatile_u64(mine, (tile == 0 ? ST_FLAG_PREFIX : ST_FLAG_AGG) | tag | (u64)total);
        s_dstart[tid] = dstart;
#pragma unroll
        for (int w = 0; w < VW; ++w) s_whist[w * STRIDE + tid] += dstart;   // absolute slots
    }
    __syncthreads();

    // -- E: decoupled look-back, RS_WINDOW predecessors per round trip -----------------------------------
    if (tid < (int)radix) {
        u64 excl = 0;
        if (tile > 0) {
            int64_t t = tile - 1;
            while (true) {
                u64 v[RS_WINDOW];
#pragma unroll
                for (int j = 0; j < RS_WINDOW; ++j)
                    v[j] = (t - j >= 0) ? ld_volatile_u64(status + (size_t)(t - j) * radix + tid)
                                        : (ST_FLAG_PREFIX | tag);
                bool done = false;
                int consumed = 0;
#pragma unroll
                for (int j = 0; j < RS_WINDOW; ++j) {
                    if (done || consumed != j) continue;
                    const u64 x = v[j];
                    if ((u32)((x >> 56) & 0x3F) != epoch || (x >> 62) == 0) continue;   // not published yet
                    excl += x & ST_VALUE_MASK;
                    ++consumed;
                    if ((x >> 62) == 2) done = true;
                }
                if (done) break;
                t -= consumed;
                if (consumed == 0) __nanosleep(20);
            }
            st_volatile_u64(mine, ST_FLAG_PREFIX | tag | (excl + (u64)total));
        }
        // ptr_base (fused partition + exchange): digit d is a destination rank and its run starts at a
        // byte address inside that rank's receive buffer, mapped into this process over NVLink
        s_gptr[tid] = ptr_base ? ptr_base[tid] + (excl - (u64)dstart) * 8u
                               : reinterpret_cast<u64>(keys_out) + (digit_base[tid] + excl - (u64)dstart) * 8u;
        const u32 first_of = s_tile[4];
        if (first_of != ~0u)                                  // first tile of a segment: bucket bounds (see RsSegs)
            atomicMin(reinterpret_cast<unsigned long long*>(segs->bounds + (size_t)tid * segs->n_b + first_of / segs->n_a),
                      (unsigned long long)(digit_base[tid] + excl));
    }

    // -- D: scatter into the digit-ordered staging buffer ---------------------------------------------------
    if (full) stage_rows<HAS_VALS, true, IPT, SPLIT, NB>(key, val, rnk2, shift, mask, my_n, my_whist, s_keys, s_vals);
    else stage_rows<HAS_VALS, false, IPT, SPLIT, NB>(key, val, rnk2, shift, mask, my_n, my_whist, s_keys, s_vals);
    __syncthreads();

    // -- F: coalesced per-digit runs out --------------------------------------------------------------------
    if (full) write_rows<HAS_VALS, true, IPT>(s_keys, s_vals, s_gptr, shift, mask, tid, tile_n, keys_out, vals_out);
    else write_rows<HAS_VALS, false, IPT>(s_keys, s_vals, s_gptr, shift, mask, tid, tile_n, keys_out, vals_out);
}

static size_t rs_smem_bytes(bool has_vals, int nb) {
    size_t b = (size_t)rs_tile(has_vals) * 8 + (size_t)RS_WARPS * RS_SPLIT * ((size_t)1 << nb) * 4 + RS_RADIX * 4 +
               RS_RADIX * 8 + 48 * 4 + 16 * 4;
    if (has_vals) b += (size_t)rs_tile(true) * 4;
    return b;
}

// Opt-in to more than 48 KB of dynamic shared memory.  Function attributes belong to the CURRENT DEVICE, so the
// opt-in is tracked per context (one context per device), never in a process-wide static.
void cov_func_smem(ottocov_ctx* ctx, const void* func, size_t bytes) {
    auto it = ctx->func_smem.find(func);
    if (it != ctx->func_smem.end() && it->second >= bytes) return;
    CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctx->func_smem[func] = bytes;
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

PassList make_pass_list(const BitField* fields, int n_fields) {
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
    return pl;
}

// one distribution pass: grid = n_tiles CTAs (an upper bound when the input is segmented)
static void launch_pass(ottocov_ctx* ctx, int family, bool has_vals, const u64* keys, u64* out, const u32* vals, u32* vout,
                        int64_t n, int64_t n_tiles, int shift, int bits, const u64* digit_base, const u64* ptr_base,
                        u32* ticket, const RsSegs* segs, const u32* abort_flag = nullptr) {
    const u32 epoch = next_epoch(ctx);
    const double pass_bytes = (has_vals ? 24.0 : 16.0) * (double)n;
#define RS_LAUNCH(HV, NBITS)                                                                                          \
    do {                                                                                                              \
        auto kern = rs_onesweep_kernel<HV, NBITS>;                                                                    \
        const size_t smem = rs_smem_bytes(HV, NBITS);                                                                 \
        cov_func_smem(ctx, (const void*)kern, smem);                                                                  \
        COV_LAUNCH(ctx, family, pass_bytes, kern, (unsigned)n_tiles, RS_THREADS, smem, keys, out, vals, vout, n, shift, \
                   bits, digit_base, ptr_base, ctx->sweep_status, ticket, epoch, segs, abort_flag);                   \
    } while (0)
#define RS_DISPATCH(HV)                                                                                               \
    if (bits <= 4) RS_LAUNCH(HV, 4);                                                                                  \
    else if (bits <= 6) RS_LAUNCH(HV, 6);                                                                             \
    else if (bits == 7) RS_LAUNCH(HV, 7);                                                                             \
    else RS_LAUNCH(HV, 8)
    if (has_vals) { RS_DISPATCH(true); } else { RS_DISPATCH(false); }
#undef RS_DISPATCH
#undef RS_LAUNCH
}

// pre_hist (optional): device array [passes][RS_RADIX] of raw digit counts of exactly these keys and fields
// (make_pass_list order), e.g. accumulated by the kernel that wrote the keys; it is scanned in place and
// the histogram read of the keys is skipped.
// seg_cnt / seg_off (optional, keys only): the input is not one array but n_a * n_b regions inside `keys`
// (rs_seg_build_kernel's layout) holding seg_cnt[slot] keys from key offset seg_off[slot] (device arrays), n keys in
// total; the first pass reads them as segments and writes `alt` contiguously.  `keys` must hold at least n keys for
// the later passes.
int radix_sort_pairs(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n,
                     const BitField* fields, int n_fields, u64* pre_hist) {
    return radix_sort_passes(ctx, keys, alt, vals, valt, n, make_pass_list(fields, n_fields), pre_hist);
}

int radix_sort_passes(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n, const PassList& pl,
                      u64* pre_hist, const u64* seg_cnt, const u64* seg_off, int n_a, int n_b, const u32* abort_flag,
                      u64* bucket_bounds) {
    if (n <= 0 || pl.n == 0 || (n == 1 && !seg_cnt)) return 0;
    const bool has_vals = vals != nullptr;
    if (seg_cnt && has_vals) COV_THROW(OTTOCOV_ERR_ARG, "segmented radix input is keys-only");
    if (bucket_bounds && !(seg_cnt && pl.n == 1)) COV_THROW(OTTOCOV_ERR_ARG, "bucket bounds come from a single pass over segments");

    DevBuf<u64> ghist;
    u64* gh = pre_hist;
    if (!gh) {
        if (seg_cnt) COV_THROW(OTTOCOV_ERR_ARG, "segmented radix input needs the pass histograms");
        ghist.alloc(ctx, (size_t)pl.n * RS_RADIX);
        gh = ghist.p;
        CUDA_CHECK(cudaMemsetAsync(gh, 0, (size_t)pl.n * RS_RADIX * sizeof(u64), ctx->stream));
        int grid = (int)imin64(ceil_div64(n, 256 * 8), (int64_t)ctx->num_sms * 8);
        COV_LAUNCH(ctx, OTTOCOV_K_HIST, 8.0 * n, rs_histogram_kernel, grid, 256,
                   (size_t)pl.n * RS_RADIX * sizeof(u32), keys, n, pl, gh);
    }
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_scan_hist_kernel, pl.n, RS_RADIX, 0, gh);

    const int64_t tile = rs_tile(has_vals);
    const int n_segs = seg_cnt ? n_a * n_b : 0;
    const int64_t n_tiles = ceil_div64(n, tile);
    const int64_t n_tiles_first = seg_cnt ? n / tile + n_segs : n_tiles;       // upper bound: one partial tile per segment
    ensure_sweep_state(ctx, (size_t)(n_tiles_first > n_tiles ? n_tiles_first : n_tiles) * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ctx->sweep_ticket, 0, RS_MAX_PASSES * sizeof(u32), ctx->stream));

    DevBuf<unsigned char> segbuf, seg_tiles;
    const RsSegs* segs = nullptr;
    if (seg_cnt) {
        const size_t o_ts = 64, o_in = o_ts + (((size_t)n_segs + 1) * 4 + 7) / 8 * 8, o_cn = o_in + (size_t)n_segs * 8;
        segbuf.alloc(ctx, o_cn + (size_t)n_segs * 8);
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_seg_build_kernel, 1, 256, 0, seg_cnt, seg_off, n_a, n_b, (u32)tile,
                   reinterpret_cast<RsSegs*>(segbuf.p), reinterpret_cast<u32*>(segbuf.p + o_ts),
                   reinterpret_cast<u64*>(segbuf.p + o_in), reinterpret_cast<u64*>(segbuf.p + o_cn));
        segs = reinterpret_cast<const RsSegs*>(segbuf.p);
        seg_tiles.alloc(ctx, (size_t)n_tiles_first * 16 + 16);
        u64* t_in = reinterpret_cast<u64*>(seg_tiles.p);
        u32* t_n = reinterpret_cast<u32*>(seg_tiles.p + (size_t)n_tiles_first * 8);
        u32* t_first = t_n + n_tiles_first;
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_seg_tiles_kernel, (unsigned)ceil_div64(n_tiles_first, 256), 256, 0, segs, (u32)tile,
                   (u32)n_tiles_first, t_in, t_n, t_first);
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_seg_attach_kernel, 1, 1, 0, reinterpret_cast<RsSegs*>(segbuf.p), (const u64*)t_in,
                   (const u32*)t_n, (const u32*)t_first, bucket_bounds, (u32)n_a, (u32)n_b);
    }
    for (int p = 0; p < pl.n; ++p) {
        const bool first_seg = (p == 0 && segs);
        launch_pass(ctx, OTTOCOV_K_SORT_PASS, has_vals, keys, alt, vals, valt, n, first_seg ? n_tiles_first : n_tiles,
                    pl.shift[p], pl.bits[p], gh + (size_t)p * RS_RADIX, nullptr, ctx->sweep_ticket + p, first_seg ? segs : nullptr,
                    abort_flag);
        u64* tk = keys; keys = alt; alt = tk;
        if (has_vals) { u32* tv = vals; vals = valt; valt = tv; }
    }
    return pl.n;
}

// One stable distribution pass on key bits [shift, shift + bits) whose per-digit output runs start at the
// byte addresses in ptr_base_host[digit] (device addresses; digits are destination ranks and the
// addresses may be peer memory mapped over NVLink): partition and exchange fused into one kernel.
void radix_partition_push(ottocov_ctx* ctx, const u64* keys, int64_t n, int shift, int bits,
                          const u64* ptr_base_host, int n_digits) {
    if (n <= 0) return;
    if (bits < 1 || bits > RS_MAX_BITS || n_digits > (1 << bits)) COV_THROW(OTTOCOV_ERR_ARG, "bad digit width for push");
    DevBuf<u64> pb(ctx, RS_RADIX);
    u64 h[RS_RADIX];
    for (int d = 0; d < RS_RADIX; ++d) h[d] = d < n_digits ? ptr_base_host[d] : 0;
    // `h` is pageable: the runtime stages it before cudaMemcpyAsync returns, so the stack frame may go away
    CUDA_CHECK(cudaMemcpyAsync(pb.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t n_tiles = ceil_div64(n, rs_tile(false));
    ensure_sweep_state(ctx, (size_t)n_tiles * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ctx->sweep_ticket, 0, RS_MAX_PASSES * sizeof(u32), ctx->stream));
    launch_pass(ctx, OTTOCOV_K_PARTITION, false, keys, nullptr, nullptr, nullptr, n, n_tiles, shift, bits, nullptr, pb.p,
                ctx->sweep_ticket, nullptr);
}
