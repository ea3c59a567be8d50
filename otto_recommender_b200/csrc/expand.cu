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

#ifndef OTTOCOV_EX_DIRECT
#define OTTOCOV_EX_DIRECT 0
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
    lo_out[j] = (u32)j + 1u;
    cnt_out[j] = hi - (u32)j - 1u;
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
        rec_lo[r] = lo[j];
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

// Persistent: CTA b writes tiles b, b + gridDim.x, ...  Per tile the records that own its outputs are staged in
// shared memory; thread t then produces the EX_PER consecutive outputs [EX_PER t, EX_PER t + EX_PER): one binary
// search for the first, a linear walk over the record offsets for the rest.  The keys go through a swizzled
// shared-memory transpose and leave as 128-bit stores, 512 contiguous bytes per warp instruction.
// ghist != nullptr: the digit histograms of the distribution passes that will sort these keys (pl) are
// accumulated here, in shared memory while the keys are still in registers, and flushed once per CTA -- the
// sort's own histogram kernel (one more read of every key) is not needed.
template <bool SELF, bool CANON, bool MIX>
__global__ void __launch_bounds__(EX_THREADS)
expand_kernel(const u32* __restrict__ rec_src, const u32* __restrict__ rec_lo,
              const u64* __restrict__ rec_off, const u32* __restrict__ tile_rec,
              const u32* __restrict__ aid_src, const u32* __restrict__ aid_tgt, u64 out_begin,
              u64 out_end, u64* __restrict__ dst, u32 n_dest, KeyMix mix, int64_t n_tiles, PassList pl,
              u64* __restrict__ ghist) {
    constexpr int PITCH = EX_TILE + 2;
    __shared__ __align__(16) u32 s_buf[3 * PITCH];
    __shared__ u32 s_src[SELF ? EX_TILE + 1 : 1];
    extern __shared__ u32 s_hist[];                   // [pl.n][RS_RADIX] when ghist
    u32* s_off = s_buf;
    u32* s_lo = s_buf + PITCH;
    u32* s_aid = s_buf + 2 * PITCH;
    u64* s_out = reinterpret_cast<u64*>(s_buf);       // [EX_TILE]: re-uses s_off / s_lo once the keys sit in registers
    const bool hist = ghist != nullptr;
    if (hist)
        for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += EX_THREADS) s_hist[j] = 0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
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
            u32 lo = rec_lo[r];
            u32 rel;
            if (off <= o0) { rel = 0; lo += (u32)(o0 - off); }   // only j == 0: skip outputs of earlier tiles
            else rel = (u32)(off - o0);
            s_off[j] = rel;
            s_lo[j] = lo;
            s_aid[j] = aid_src[src];
            if (SELF) s_src[j] = src;
        }
        __syncthreads();

        u64 key[EX_PER];
        const u32 k0 = (u32)EX_PER * threadIdx.x;
        if (k0 < n_out) {
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
                    key[q] = make_pair_key<CANON, MIX>(a, aid_tgt[tgt], n_dest, mix);
                }
            }
            if (hist) {
                for (int p = 0; p < pl.n; ++p) {          // pass parameters are read once per pass, not per key
                    const int sh = pl.shift[p];
                    const u32 msk = (1u << pl.bits[p]) - 1u;
                    u32* hrow = s_hist + p * RS_RADIX;
#pragma unroll
                    for (int q = 0; q < EX_PER; ++q)
                        if (k0 + q < n_out) atomicAdd(&hrow[(u32)(key[q] >> sh) & msk], 1u);
                }
            }
        }
#if OTTOCOV_EX_DIRECT
        // measured variant (experiments/README.md): 8 consecutive keys leave straight from the registers as four
        // 128-bit stores per thread (64 B per thread, sectors half-filled per instruction, merged in L2)
        {
            u64* tile_dst = dst + (size_t)tile * EX_TILE;
            if (k0 < n_out) {
#pragma unroll
                for (int q = 0; q < EX_PER; q += 2) {
                    if (k0 + q + 1 < n_out && vec_ok) {
                        ulonglong2 v; v.x = key[q]; v.y = key[q + 1];
                        __stcs(reinterpret_cast<ulonglong2*>(tile_dst + k0 + q), v);
                    } else {
                        if (k0 + q < n_out) __stcs(tile_dst + k0 + q, key[q]);
                        if (k0 + q + 1 < n_out) __stcs(tile_dst + k0 + q + 1, key[q + 1]);
                    }
                }
            }
        }
#else
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
#endif
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

struct Segment {
    int tgt_type;
    bool self;
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
    u32 W = 0;
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
        pl->W = (u32)(spec->window > 86400 ? 86400 : spec->window);   // count_co_events.py:33-36
        pl->aid_bits = ctx->info.aid_bits > 0 ? ctx->info.aid_bits : 1;
        pl->user_min = spec->min_count > 1 ? spec->min_count : 1;
        // symmetric shortcut: one canonical key per unordered event pair, mirrored after the reduce.  It
        // pays when few rows are left to mirror (a threshold) or when the keys are about to cross NVLink
        // anyway; OTTOCOV_SYM_OFF / OTTOCOV_SYM_ON force the choice.
        const bool sym_kind = spec->next_mask == (1u << pl->A);
        pl->sym = sym_kind && !(spec->flags & OTTOCOV_SYM_OFF) &&
                  ((spec->flags & OTTOCOV_SYM_ON) || pl->user_min > 1 || distributed);
        const TypeArray& src = ctx->ta[pl->A];
        for (int B = 0; B < 3; ++B) {
            if (!((spec->next_mask >> B) & 1)) continue;
            const TypeArray& tgt = ctx->ta[B];
            if (src.n == 0 || tgt.n == 0) continue;
            Segment* sg = new Segment();
            pl->segs.push_back(sg);
            sg->tgt_type = B;
            sg->self = (pl->A == B);
            const u32* xr = nullptr;
            if (!sg->self) xr = (B == (pl->A + 1) % 3) ? src.xrank[0] : src.xrank[1];
            DevBuf<u32> lo(ctx, src.n), cnt(ctx, src.n);
            if (pl->sym)
                COV_LAUNCH(ctx, OTTOCOV_K_WINDOW, 16.0 * src.n, window_fwd_kernel, (unsigned)ceil_div64(src.n, 256), 256, 0,
                           src.skey, src.n, pl->W, lo.p, cnt.p);
            else
                COV_LAUNCH(ctx, OTTOCOV_K_WINDOW, 20.0 * src.n, window_kernel, (unsigned)ceil_div64(src.n, 256), 256, 0,
                           src.skey, xr, src.n, tgt.skey, (u32)tgt.n, pl->W, lo.p, cnt.p);
            sg->rec_src.alloc(ctx, src.n); sg->rec_lo.alloc(ctx, src.n); sg->rec_off.alloc(ctx, src.n);
            WindowRecords f;
            f.lo = lo.p; f.cnt = cnt.p;
            f.rec_src = sg->rec_src.p; f.rec_lo = sg->rec_lo.p; f.rec_off = sg->rec_off.p;
            u64 tot[2];
            scan_apply(ctx, OTTOCOV_K_WINDOW, f, src.n, tot, 2.0 * 4.0 * src.n + 8.0 * src.n + 16.0 * src.n);
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
    const TypeArray& src = ctx->ta[pl->A];
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
            const TypeArray& tgt = ctx->ta[sg->tgt_type];
            u64* dst = dst_base + (a - c0);
            const double bytes = 8.0 * (double)(oe - ob);
            const unsigned grid = (unsigned)imin64(n_tiles, (int64_t)ctx->num_sms * 5);     // 5 CTAs / SM at 48 registers
#define EX_LAUNCH(SELF_, CANON_, MIX_)                                                                              \
            COV_LAUNCH(ctx, OTTOCOV_K_EXPAND, bytes, (expand_kernel<SELF_, CANON_, MIX_>), grid, EX_THREADS, hist_smem,  \
                       sg->rec_src.p, sg->rec_lo.p, sg->rec_off.p, tile_rec.p, src.aid, tgt.aid, ob, oe, dst, n_dest, mx, \
                       n_tiles, hp, ghist)
            if (mix) {
                if (pl->sym) EX_LAUNCH(false, true, true);
                else if (sg->self) EX_LAUNCH(true, false, true);
                else EX_LAUNCH(false, false, true);
            } else {
                if (pl->sym) EX_LAUNCH(false, true, false);
                else if (sg->self) EX_LAUNCH(true, false, false);
                else EX_LAUNCH(false, false, false);
            }
#undef EX_LAUNCH
        }
        seg_start = seg_end;
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
