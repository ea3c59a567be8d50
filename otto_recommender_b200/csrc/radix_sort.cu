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
template <int NB>
__device__ __forceinline__ u32 match_digit_ballot(u32 d) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        u32 m;
        asm volatile(
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
// bumps the warp's running digit counters.  FULL = every row is valid (all tiles but the last).
template <int ALGO, int NB, bool FULL, int IPT>
__device__ __forceinline__ void rank_rows(const u64 (&key)[IPT], int shift, u32 mask, u32 radix, int my_n, int lane,
                                          u32 lt, u32* my_whist, u32* mm, u32 (&rnk2)[IPT / 2]) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const bool valid = FULL || i < my_n;
        const u32 d = (u32)(key[i] >> shift) & mask;
        u32 peers;
        if (ALGO == 0) {
            peers = __match_any_sync(0xffffffffu, valid ? d : radix);      // `radix` never matches a digit
        } else if (ALGO == 1) {
            peers = match_digit_ballot<NB>(d);
            if (!FULL) {
                const u32 vm = __ballot_sync(0xffffffffu, valid);
                peers &= valid ? vm : ~vm;
            }
        } else {
            if (valid) atomicOr(&mm[d], 1u << lane);
            __syncwarp();
            peers = valid ? mm[d] : (1u << lane);
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) mm[d] = 0;              // ready for the next row
        }
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader && valid) {
            old = my_whist[d];
            my_whist[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        const u32 r = old + __popc(peers & lt);
        if (i & 1) rnk2[i >> 1] |= r << 16; else rnk2[i >> 1] = r;
        __syncwarp();
    }
}

// ALGO selects how lanes holding the same digit find each other (tuning knob OTTOCOV_RS_ALGO, measured in
// experiments/README.md; 1 is the default):
//   0  match.any            1  NB ballots (one per digit bit)            2  atomicOr on a per-warp mask table
// phases A, D and F for one tile; FULL = all 4096 slots valid (every tile but the last): no predicates
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

template <bool HAS_VALS, bool FULL, int IPT>
__device__ __forceinline__ void stage_rows(const u64 (&key)[IPT], const u32 (&val)[IPT], const u32 (&rnk2)[IPT / 2],
                                           int shift, u32 mask, int my_n, const u32* my_whist, u64* s_keys, u32* s_vals) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        if (FULL || i < my_n) {
            const u32 d = (u32)(key[i] >> shift) & mask;
            const u32 r = (i & 1) ? (rnk2[i >> 1] >> 16) : (rnk2[i >> 1] & 0xFFFFu);
            const u32 pos = my_whist[d] + r;
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

#ifndef OTTOCOV_RS_MINB
#define OTTOCOV_RS_MINB 2          // keys only: 40 keys per thread need the 128-register budget of 2 CTAs/SM
#endif
#ifndef OTTOCOV_RS_MINB_PAIRS
#define OTTOCOV_RS_MINB_PAIRS 4    // (key, value) tiles of 4096: 64 registers, 4 CTAs/SM
#endif
template <bool HAS_VALS, int ALGO, int NB>
__global__ void __launch_bounds__(RS_THREADS, HAS_VALS ? OTTOCOV_RS_MINB_PAIRS : OTTOCOV_RS_MINB)
rs_onesweep_kernel(const u64* __restrict__ keys_in, u64* __restrict__ keys_out,
                   const u32* __restrict__ vals_in, u32* __restrict__ vals_out, int64_t n, int shift,
                   int bits, const u64* __restrict__ digit_base, const u64* __restrict__ ptr_base, u64* status,
                   u32* ticket, u32 epoch) {
    constexpr int IPT = HAS_VALS ? RS_IPT_PAIRS : RS_IPT_KEYS;
    constexpr int RS_TILE = RS_THREADS * IPT;
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_keys = reinterpret_cast<u64*>(s_raw);                              // [RS_TILE]
    u32* s_whist = reinterpret_cast<u32*>(s_keys + RS_TILE);                  // [RS_WARPS][RS_RADIX]
    u32* s_dstart = s_whist + RS_WARPS * RS_RADIX;                            // [RS_RADIX]
    u64* s_gptr = reinterpret_cast<u64*>(s_dstart + RS_RADIX);                // [RS_RADIX]
    u32* s_scan = reinterpret_cast<u32*>(s_gptr + RS_RADIX);                  // [RS_WARPS + 1] (+pad to 48)
    u32* s_tile = s_scan + 48;                                                // [1] (+pad to 16)
    u32* s_match = s_tile + 16;                                               // [RS_WARPS][RS_RADIX] if ALGO == 2
    u32* s_vals = s_match + (ALGO == 2 ? RS_WARPS * RS_RADIX : 0);            // [RS_TILE] if HAS_VALS

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 radix = 1u << bits, mask = radix - 1u;

    if (tid == 0) *s_tile = atomicAdd(ticket, 1u);
    for (int j = tid; j < RS_WARPS * RS_RADIX; j += RS_THREADS) {
        s_whist[j] = 0;
        if (ALGO == 2) s_match[j] = 0;
    }
    __syncthreads();
    const int64_t tile = *s_tile;
    const int64_t tile_base = tile * RS_TILE;
    const int tile_n = (int)min((int64_t)RS_TILE, n - tile_base);
    const bool full = tile_n == RS_TILE;

    // -- A: load (warp-striped: every access is a coalesced 256 B row) ------------------------------
    u64 key[IPT];
    u32 val[IPT];
    const int64_t lbase = tile_base + (int64_t)warp * 32 * IPT + lane;
    const int my_n = (int)min((int64_t)IPT, (n - lbase + 31) / 32);        // valid items of this lane (may be <= 0)
    if (full) load_rows<HAS_VALS, true>(keys_in, vals_in, lbase, my_n, key, val);
    else load_rows<HAS_VALS, false>(keys_in, vals_in, lbase, my_n, key, val);

    // -- B: stable rank of every key inside its warp's 512-key segment; the per-warp digit counts fall
    //       out of the same pass.  Lanes holding the same digit are found with one ballot per digit
    //       bit (match.any is far slower than `bits` ballots on this part); the lowest such lane
    //       bumps the warp's running counter for the digit, no atomics needed.
    u32* my_whist = s_whist + warp * RS_RADIX;
    const u32 lt = lanemask_lt();
    u32 rnk2[IPT / 2];                       // two 16-bit ranks per register
    u32* mm = (ALGO == 2) ? s_match + warp * RS_RADIX : nullptr;
    if (full) rank_rows<ALGO, NB, true>(key, shift, mask, radix, my_n, lane, lt, my_whist, mm, rnk2);
    else rank_rows<ALGO, NB, false>(key, shift, mask, radix, my_n, lane, lt, my_whist, mm, rnk2);
    __syncthreads();

    // -- C: offsets over warps, tile totals, digit starts; publish the aggregate ----------------------
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
    const u64 tag = (u64)epoch << 56;
    u64* mine = status + (size_t)tile * radix + tid;          // [tile][digit]: a tile's words are contiguous
    if (tid < (int)radix) {
        st_volatile_u64(mine, (tile == 0 ? ST_FLAG_PREFIX : ST_FLAG_AGG) | tag | (u64)total);
        s_dstart[tid] = dstart;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) s_whist[w * RS_RADIX + tid] += dstart;   // absolute slots
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
    }

    // -- D: scatter into the digit-ordered staging buffer ---------------------------------------------------
    if (full) stage_rows<HAS_VALS, true>(key, val, rnk2, shift, mask, my_n, my_whist, s_keys, s_vals);
    else stage_rows<HAS_VALS, false>(key, val, rnk2, shift, mask, my_n, my_whist, s_keys, s_vals);
    __syncthreads();

    // -- F: coalesced per-digit runs out --------------------------------------------------------------------
    if (full) write_rows<HAS_VALS, true, IPT>(s_keys, s_vals, s_gptr, shift, mask, tid, tile_n, keys_out, vals_out);
    else write_rows<HAS_VALS, false, IPT>(s_keys, s_vals, s_gptr, shift, mask, tid, tile_n, keys_out, vals_out);
}

static size_t rs_smem_bytes(bool has_vals, int algo) {
    size_t b = (size_t)rs_tile(has_vals) * 8 + (size_t)RS_WARPS * RS_RADIX * 4 + RS_RADIX * 4 + RS_RADIX * 8 +
               48 * 4 + 16 * 4;
    if (algo == 2) b += (size_t)RS_WARPS * RS_RADIX * 4;
    if (has_vals) b += (size_t)rs_tile(true) * 4;
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

// pre_hist (optional): device array [passes][RS_RADIX] of raw digit counts of exactly these keys and fields
// (make_pass_list order), e.g. accumulated by the kernel that wrote the keys; it is scanned in place and
// the histogram read of the keys is skipped.
int radix_sort_pairs(ottocov_ctx* ctx, u64*& keys, u64*& alt, u32*& vals, u32*& valt, int64_t n,
                     const BitField* fields, int n_fields, u64* pre_hist) {
    const PassList pl = make_pass_list(fields, n_fields);
    if (n <= 1 || pl.n == 0) return 0;
    const bool has_vals = vals != nullptr;

    DevBuf<u64> ghist;
    u64* gh = pre_hist;
    if (!gh) {
        ghist.alloc(ctx, (size_t)pl.n * RS_RADIX);
        gh = ghist.p;
        CUDA_CHECK(cudaMemsetAsync(gh, 0, (size_t)pl.n * RS_RADIX * sizeof(u64), ctx->stream));
        int grid = (int)imin64(ceil_div64(n, 256 * 8), (int64_t)ctx->num_sms * 8);
        COV_LAUNCH(ctx, OTTOCOV_K_HIST, 8.0 * n, rs_histogram_kernel, grid, 256,
                   (size_t)pl.n * RS_RADIX * sizeof(u32), keys, n, pl, gh);
    }
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, rs_scan_hist_kernel, pl.n, RS_RADIX, 0, gh);

    const int64_t n_tiles = ceil_div64(n, rs_tile(has_vals));
    ensure_sweep_state(ctx, (size_t)n_tiles * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ctx->sweep_ticket, 0, RS_MAX_PASSES * sizeof(u32), ctx->stream));

    static int algo = -1;
    if (algo < 0) {
        const char* e = getenv("OTTOCOV_RS_ALGO");            // tuning knob, default = ballots
        algo = e ? atoi(e) : 1;
        if (algo < 0 || algo > 2) algo = 1;
    }
    const size_t smem = rs_smem_bytes(has_vals, algo);
    const double pass_bytes = (has_vals ? 24.0 : 16.0) * (double)n;
    for (int p = 0; p < pl.n; ++p) {
        const u32 epoch = next_epoch(ctx);
        const int bits = pl.bits[p];
#define RS_LAUNCH(HV, AL, NBITS)                                                                            \
        do {                                                                                                \
            auto kern = rs_onesweep_kernel<HV, AL, NBITS>;                                                  \
            static bool attr_done = false;                                                                  \
            if (!attr_done) {                                                                               \
                CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(HV, 2))); \
                CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
                attr_done = true;                                                                           \
            }                                                                                               \
            COV_LAUNCH(ctx, OTTOCOV_K_SORT_PASS, pass_bytes, kern, (unsigned)n_tiles, RS_THREADS, smem, keys, alt, \
                       (const u32*)vals, valt, n, pl.shift[p], bits, gh + (size_t)p * RS_RADIX,        \
                       (const u64*)nullptr, ctx->sweep_status, ctx->sweep_ticket + p, epoch);                                    \
        } while (0)
#define RS_DISPATCH(HV)                                                                                     \
        if (algo == 0) RS_LAUNCH(HV, 0, 8);                                                                 \
        else if (algo == 2) RS_LAUNCH(HV, 2, 8);                                                            \
        else if (bits <= 4) RS_LAUNCH(HV, 1, 4);                                                            \
        else if (bits <= 6) RS_LAUNCH(HV, 1, 6);                                                            \
        else if (bits == 7) RS_LAUNCH(HV, 1, 7);                                                            \
        else RS_LAUNCH(HV, 1, 8)
        if (has_vals) { RS_DISPATCH(true); } else { RS_DISPATCH(false); }
#undef RS_DISPATCH
#undef RS_LAUNCH
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
    CUDA_CHECK(cudaMemcpyAsync(pb.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t n_tiles = ceil_div64(n, rs_tile(false));
    ensure_sweep_state(ctx, (size_t)n_tiles * RS_RADIX);
    CUDA_CHECK(cudaMemsetAsync(ctx->sweep_ticket, 0, RS_MAX_PASSES * sizeof(u32), ctx->stream));
    const u32 epoch = next_epoch(ctx);
    auto kern = (bits <= 4) ? rs_onesweep_kernel<false, 1, 4> : rs_onesweep_kernel<false, 1, 8>;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false, 2)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false, 1, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false, 2)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false, 1, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_done = true;
    }
    COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 16.0 * n, kern, (unsigned)n_tiles, RS_THREADS, rs_smem_bytes(false, 1), keys,
               (u64*)nullptr, (const u32*)nullptr, (u32*)nullptr, n, shift, bits, (const u64*)nullptr, (const u64*)pb.p,
               ctx->sweep_status, ctx->sweep_ticket, epoch);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));      // `h` lives on this stack frame
}
