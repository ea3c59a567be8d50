// hash_reduce.cu -- bucketed hash aggregation: the thresholded groupby(['aid','aid_next']).count() of
// model/count_co_events.py:70-71 + :131-132 / :172 without a full sort.
//
// A full LSD sort of the pair keys orders them by all 2 x aid_bits key bits (6 passes for 1.8 M aids)
// only to find equal keys next to each other.  Counting needs less: equal keys must MEET, nothing more.
//
//   1. the expansion writes each key through a bijective mix (KeyMix, internal.cuh): equal keys stay
//      equal, and the top bits of the mixed key are uniform whatever the popularity skew of the aids;
//   2. radix_sort.cu sorts on the top `bb` bits of the mixed key only (3 passes instead of 6 at the
//      headline size) -- the keys are now grouped into 2^bb buckets of a few hundred keys each, and
//      every copy of a key lies in one bucket;
//   3. hash_reduce_kernel: a CTA owns the buckets that START inside its 2048-key tile (it skips the
//      head of the tile that continues the previous CTA's bucket and reads past the tile end to finish
//      its last bucket), counts their keys in a shared-memory open-addressing table of packed words
//      tag(42) | count(22) -- one 64-bit CAS claims AND counts a new key, one 32-bit add counts a
//      repeated one -- then scans the table count-first: rows with count >= min_count are un-mixed and
//      appended to the output (one global atomic per CTA); for symmetric kinds the diagonal doubling
//      and the mirrored row (b, a, c) are emitted in the same step;
//   4. the few surviving rows are sorted by plain key (they arrive in bucket order).
// HBM traffic after the expansion: passes x 16 P + 8 P (the pass histograms come from the expansion),
// against 8 P + 6 x 16 P + 8 P.  The kernel is bound by shared-memory atomics: one per key.
//
// The table holds DISTINCT keys only, so hot pairs (one key repeated millions of times) cost nothing
// but the streaming read.  If a CTA ever meets more distinct keys than its table holds (needs an
// adversarial input: buckets are balanced by the hash), a bucket tail runs past 2 M keys, or a tag
// does not fit its field, a flag is raised and the host re-does the reduce with the full sort (the
// mixed keys are un-mixed in place first).  Keys too wide for the packed word use an unpacked table
// (64-bit key + 32-bit count).
//
// Fused first pass (round 2): the first distribution pass needs no stability (there is no earlier order to
// keep), and hashed digits are uniform, so the expansion kernel itself scatters the keys into one slack region
// per digit (expand.cu, expand_scatter_kernel).  hashed_reduce then starts from those regions (HashPre): the
// remaining passes read them as segments.  HBM traffic: 8 P + (passes - 1) x 16 P + 8 P.
//
// Whole buckets (round 2, HashPre::big): with 16 bucket bits instead of 21 the fused pass + ONE more pass are
// enough, and hash_reduce_buckets_kernel counts one whole ~11 k-key bucket per CTA (further down: "whole buckets").
// Taken where it saves a pass (hashed_big_bucket_bits); steps 3-4 above describe the tile kernel, which still serves
// everything else (small kinds, the owners' side of the multi-GPU exchange, the streamed ingest).
#include "internal.cuh"
#include "scan.cuh"

// CTA shape.  Measured on the headline run (742 M keys): 512 threads / 8192 slots / 3 CTAs per SM 5.47 ms,
// 256 threads / 4096 slots / 6 CTAs per SM 5.15 ms -- the same 48 warps per SM, but the CTA barriers (the largest stall
// reason in ncu: profiles/r02_top_ncu_details.txt) wait for half as many warps.
#ifndef OTTOCOV_HR_THREADS
#define OTTOCOV_HR_THREADS 256
#endif
#ifndef OTTOCOV_HR_CAP_LOG2
#define OTTOCOV_HR_CAP_LOG2 12
#endif
#ifndef OTTOCOV_HR_MINB
#define OTTOCOV_HR_MINB 6
#endif
constexpr int HR_THREADS = OTTOCOV_HR_THREADS;
constexpr int HR_IPT = 8;
constexpr int HR_TILE = HR_THREADS * HR_IPT;     // 2048 keys per CTA (+ the tail of its last bucket)
constexpr int HR_CAP_LOG2 = OTTOCOV_HR_CAP_LOG2;
constexpr int HR_CAP = 1 << HR_CAP_LOG2;         // 4096 slots
constexpr int HR_XT = 4;                         // keys per thread per slice of a long bucket tail
constexpr int HR_MAX_TAIL_ROUNDS = 1000;         // ~2 M keys: beyond that one CTA would serialise the reduce
constexpr u64 HR_NONE = ~0ull;                   // empty slot / no key (mixed keys have <= 56 bits)
// Packed table word: tag(42) | count(22).  Shared-memory atomics cost ~2 cycles per lane on this part and
// bound the kernel, so one atomic per key matters: a new key is claimed AND counted by one 64-bit CAS, a
// repeated key by one 32-bit add on the low word.  tag = mixed key - (first bucket of the tile << rem_bits);
// the count cannot overflow its field because a CTA never counts more than HR_TILE + (HR_MAX_TAIL_ROUNDS *
// HR_XT + 1) * HR_THREADS keys.
constexpr int HR_CB = 22;
constexpr int HR_TAG_BITS = 64 - HR_CB;
constexpr u64 HR_CMASK = (1ull << HR_CB) - 1ull;
static_assert((u64)HR_TILE + ((u64)HR_MAX_TAIL_ROUNDS * HR_XT + 1) * HR_THREADS < HR_CMASK, "count field too narrow");

// PACKED: s_key[slot] = tag << 22 | count (32 KB, 6 CTAs / SM).  Otherwise (keys too wide for a 42-bit tag):
// s_key[slot] = mixed key, s_cnt[slot] = count (48 KB, 2+ CTAs / SM, two atomics for a new key).
template <bool PACKED, int CL2 = HR_CAP_LOG2>
__device__ __forceinline__ void hr_insert(u64* s_key, u32* s_cnt, u64 h, u64 base, u32 times, u32* flags) {
    const u64 tag = h - base;
    if (PACKED && (tag >> HR_TAG_BITS)) { atomicOr(flags, 8u); return; }     // bucket span wider than the tag
    u32 slot = (((u32)tag ^ (u32)(tag >> 32)) * 0x9E3779B1u) >> (32 - CL2);
#pragma unroll 1
    for (int probes = 0; probes < (1 << CL2); ++probes) {
        if (PACKED) {
            // CAS first, no look: most keys are new, and the shared-memory atomic unit is far from busy (23 % in ncu)
            // while issue slots are what the kernel runs out of
            const u64 cur = atomicCAS(reinterpret_cast<unsigned long long*>(s_key + slot), HR_NONE, (tag << HR_CB) | (u64)times);
            if (cur == HR_NONE) return;                         // claimed and counted in one step
            if ((cur >> HR_CB) == tag) { atomicAdd(reinterpret_cast<u32*>(s_key + slot), times); return; }
        } else {
            u64 cur = *reinterpret_cast<volatile u64*>(s_key + slot);
            if (cur == HR_NONE) {
                cur = atomicCAS(reinterpret_cast<unsigned long long*>(s_key + slot), HR_NONE, tag);
                if (cur == HR_NONE) cur = tag;                  // this thread claimed the slot
            }
            if (cur == tag) { atomicAdd(s_cnt + slot, times); return; }
        }
        slot = (slot + 1) & ((1 << CL2) - 1);
    }
    atomicOr(flags, 1u);                                   // table full: the host falls back to the full sort
}

// Keys past the tile end (the rest of the CTA's last bucket).  A long tail means a hot pair repeated many
// times: when a whole warp holds one key, one lane adds 32 instead of 32 lanes queueing on one counter.
template <bool PACKED, int CL2 = HR_CAP_LOG2>
__device__ __forceinline__ void hr_insert_tail(u64* s_key, u32* s_cnt, u64 h, u64 base, bool mine, u32* flags) {
    const u64 h0 = __shfl_sync(0xffffffffu, h, 0);
    if (__all_sync(0xffffffffu, mine && h == h0)) {
        if ((threadIdx.x & 31) == 0) hr_insert<PACKED, CL2>(s_key, s_cnt, h0, base, 32u, flags);
    } else if (mine) {
        hr_insert<PACKED, CL2>(s_key, s_cnt, h, base, 1u, flags);
    }
}

// slot j -> (mixed key, count); false when the slot is empty
template <bool PACKED>
__device__ __forceinline__ bool hr_slot(const u64* s_key, const u32* s_cnt, int j, u64 base, u64* h, u32* c) {
    const u64 w = s_key[j];
    if (w == HR_NONE) return false;
    if (PACKED) { *h = (w >> HR_CB) + base; *c = (u32)(w & HR_CMASK); }
    else { *h = w; *c = s_cnt[j]; }
    return true;
}

// Table scan + emit, called by every thread of the CTA (contains barriers).  The kernel is bound by issued
// instructions, and at a threshold of 10 only a slot in a hundred survives: the first look at a slot is one
// 32-bit load of its count (an empty packed slot reads as the all-ones count, which no real count reaches) and a
// compare with the lowest count that could survive (half the threshold for a symmetric kind: a diagonal row
// counts twice); only candidates are decoded and un-mixed.  Survivors are appended to the output with one global
// atomic per CTA.  Returns false when the output buffer overflowed (flag raised; the caller gives up).
template <bool SYM, bool PACKED, int CL2 = HR_CAP_LOG2, int NT = HR_THREADS>
__device__ __forceinline__ bool hr_emit(const u64* s_key, const u32* s_cnt, u32* s_scan, unsigned long long* s_base, u32* s_over,
                                        u64 base, const KeyMix& mix, u32 min_count, int mirror, u64* __restrict__ out_keys,
                                        u32* __restrict__ out_count, unsigned long long* __restrict__ out_n, u64 out_cap,
                                        u32* __restrict__ flags) {
    constexpr int SPT = (1 << CL2) / NT;                   // slots per thread
    static_assert(2 * SPT <= 32, "two flag bits per scanned slot must fit one register");
    const int tid = threadIdx.x;
    const u32 cand = SYM ? (min_count + 1u) / 2u : min_count;
    // Pass 1, branch-free: which of this thread's slots hold a count that could survive.  Pass 2 decodes those only.
    // (One loop doing both made every warp run the decode + un-mix for every slot index at which ANY of its lanes had
    // a candidate -- most of them -- and that was the largest block of issued instructions in the kernel: ncu, round 2.)
    u32 candm = 0;
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        const int j = tid + q * NT;
        const u32 c = PACKED ? (reinterpret_cast<const u32*>(s_key + j)[0] & (u32)HR_CMASK) : s_cnt[j];
        candm |= ((c >= cand && (!PACKED || c != (u32)HR_CMASK) && c != 0u) ? 1u : 0u) << q;
    }
    u32 bits = 0, emit = 0;                                // two flag bits per slot: keep, also emit the mirrored row
    if (!SYM) {                                            // cand == min_count: every candidate survives
#pragma unroll
        for (int q = 0; q < SPT; ++q) bits |= ((candm >> q) & 1u) << (2 * q);
        emit = __popc(candm);
    } else {
        while (candm) {
            const int q = __ffs(candm) - 1;
            candm &= candm - 1u;
            u64 h; u32 c;
            hr_slot<PACKED>(s_key, s_cnt, tid + q * NT, base, &h, &c);
            const u64 plain = key_mix_inv(mix, h);
            const bool diag = (u32)(plain >> 32) == (u32)plain;
            const u64 total = diag ? 2ull * c : (u64)c;    // (a, a): both orders of each event pair
            if (total >= (u64)min_count) {
                const bool two = mirror && !diag;
                bits |= (two ? 3u : 1u) << (2 * q);
                emit += two ? 2u : 1u;
            }
        }
    }
    u32 blk_total;
    const u32 ex = block_exclusive_scan<u32, NT>(emit, s_scan, &blk_total);
    if (tid == 0) {
        unsigned long long b = 0;
        if (blk_total) {
            b = atomicAdd(out_n, (unsigned long long)blk_total);
            if (b + blk_total > out_cap) { atomicOr(flags, 2u); *s_over = 1; }
        }
        *s_base = b;
    }
    __syncthreads();
    if (*s_over) return false;
    u64 o = *s_base + ex;
    while (bits) {                                         // survivors only
        const int q = (__ffs(bits) - 1) >> 1;
        const u32 f = (bits >> (2 * q)) & 3u;
        bits &= ~(3u << (2 * q));
        const int j = tid + q * NT;
        u64 h; u32 c;
        hr_slot<PACKED>(s_key, s_cnt, j, base, &h, &c);
        const u64 plain = key_mix_inv(mix, h);
        u64 total = c;
        if (SYM && (u32)(plain >> 32) == (u32)plain) total *= 2;
        const u32 c32 = (u32)(total > 0xFFFFFFFFull ? 0xFFFFFFFFull : total);
        out_keys[o] = plain; out_count[o] = c32; ++o;
        if (f & 2u) { out_keys[o] = (plain << 32) | (plain >> 32); out_count[o] = c32; ++o; }
    }
    return true;
}

// The same scan + emit without CTA barriers (packed words only): every WARP reserves room for its survivors with its
// own global atomic (the rows are sorted afterwards, their order here is free), and every thread resets the slots it
// scanned, so the table is empty again when the function returns.  Used by the whole-bucket kernel, whose 16-warp
// CTAs (two per SM) pay dearly for each barrier.  An output overflow raises the flag (the host falls back).
template <bool SYM, int CL2, int NT>
__device__ __forceinline__ void hr_emit_warp(u64* s_key, u64 base, const KeyMix& mix, u32 min_count, int mirror,
                                             u64* __restrict__ out_keys, u32* __restrict__ out_count,
                                             unsigned long long* __restrict__ out_n, u64 out_cap, u32* __restrict__ flags) {
    constexpr int SPT = (1 << CL2) / NT;
    static_assert(2 * SPT <= 32, "two flag bits per scanned slot must fit one register");
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 cand = SYM ? (min_count + 1u) / 2u : min_count;
    u32 candm = 0;
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
        const u32 c = reinterpret_cast<const u32*>(s_key + tid + q * NT)[0] & (u32)HR_CMASK;
        candm |= ((c >= cand && c != (u32)HR_CMASK && c != 0u) ? 1u : 0u) << q;
    }
    u32 bits = 0, emit = 0;
    if (!SYM) {
#pragma unroll
        for (int q = 0; q < SPT; ++q) bits |= ((candm >> q) & 1u) << (2 * q);
        emit = __popc(candm);
    } else {
        while (candm) {
            const int q = __ffs(candm) - 1;
            candm &= candm - 1u;
            u64 h; u32 c;
            hr_slot<true>(s_key, nullptr, tid + q * NT, base, &h, &c);
            const u64 plain = key_mix_inv(mix, h);
            const bool diag = (u32)(plain >> 32) == (u32)plain;
            const u64 total = diag ? 2ull * c : (u64)c;
            if (total >= (u64)min_count) {
                const bool two = mirror && !diag;
                bits |= (two ? 3u : 1u) << (2 * q);
                emit += two ? 2u : 1u;
            }
        }
    }
    u32 inc = emit;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const u32 wtot = __shfl_sync(0xffffffffu, inc, 31);
    if (wtot) {                                            // warp-uniform
        unsigned long long b = 0;
        if (lane == 31) b = atomicAdd(out_n, (unsigned long long)wtot);
        b = __shfl_sync(0xffffffffu, b, 31);
        if (b + wtot > out_cap) {
            if (lane == 31) atomicOr(flags, 2u);
        } else {
            u64 o = b + inc - emit;
            while (bits) {
                const int q = (__ffs(bits) - 1) >> 1;
                const u32 f = (bits >> (2 * q)) & 3u;
                bits &= ~(3u << (2 * q));
                u64 h; u32 c;
                hr_slot<true>(s_key, nullptr, tid + q * NT, base, &h, &c);
                const u64 plain = key_mix_inv(mix, h);
                u64 total = c;
                if (SYM && (u32)(plain >> 32) == (u32)plain) total *= 2;
                const u32 c32 = (u32)(total > 0xFFFFFFFFull ? 0xFFFFFFFFull : total);
                out_keys[o] = plain; out_count[o] = c32; ++o;
                if (f & 2u) { out_keys[o] = (plain << 32) | (plain >> 32); out_count[o] = c32; ++o; }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < SPT; ++q) s_key[tid + q * NT] = HR_NONE;
}

template <bool PACKED, int CL2 = HR_CAP_LOG2, int NT = HR_THREADS>
__device__ __forceinline__ void hr_clear(u64* s_key, u32* s_cnt) {
    const int tid = threadIdx.x;
    uint4* kv = reinterpret_cast<uint4*>(s_key);           // two slots per 128-bit store
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
    for (int q = 0; q < (1 << CL2) / 2 / NT; ++q) kv[tid + q * NT] = ones;
    if (!PACKED) {
        uint4* cv = reinterpret_cast<uint4*>(s_cnt);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < (1 << CL2) / 4 / NT; ++q) cv[tid + q * NT] = zero;
    }
}

template <bool SYM, bool PACKED>
__global__ void __launch_bounds__(HR_THREADS, PACKED ? OTTOCOV_HR_MINB : 2)
hash_reduce_kernel(const u64* __restrict__ keys, int64_t n, int rem_bits, KeyMix mix, u32 min_count, int mirror,
                   u64* __restrict__ out_keys, u32* __restrict__ out_count, unsigned long long* __restrict__ out_n,
                   u64 out_cap, u32* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_key = reinterpret_cast<u64*>(s_raw);                              // [HR_CAP]
    u32* s_cnt = reinterpret_cast<u32*>(s_key + HR_CAP);                     // [HR_CAP] (!PACKED)
    u32* s_scan = s_cnt + (PACKED ? 0 : HR_CAP);                             // [HR_THREADS / 32 + 1]
    __shared__ unsigned long long s_base;
    __shared__ u32 s_over;

    if (*flags & HR_FLAG_FUSED_OVERFLOW) return;           // incomplete input (see rs_onesweep_kernel): the host repeats the step
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * HR_TILE;
    const int64_t t1 = min(n, t0 + (int64_t)HR_TILE);
    // bucket of the key just before the tile (it belongs to an earlier CTA, with every other key of that
    // bucket) and of the tile's last key (this CTA finishes that bucket beyond the tile end)
    const u64 pb = t0 > 0 ? (keys[t0 - 1] >> rem_bits) : HR_NONE;
    const u64 lb = keys[t1 - 1] >> rem_bits;
    if (lb == pb) return;                                  // the whole tile continues an earlier CTA's bucket
    // every key this CTA counts is >= the first key's bucket: tags are relative to it
    const u64 base = PACKED ? ((keys[t0] >> rem_bits) << rem_bits) : 0ull;

    u64 k[HR_IPT];
#pragma unroll
    for (int r = 0; r < HR_IPT; ++r) {
        const int64_t i = t0 + r * HR_THREADS + tid;
        k[r] = (i < t1) ? __ldcs(keys + i) : HR_NONE;
    }
    u64 kx = (t1 + tid < n) ? __ldcs(keys + t1 + tid) : HR_NONE;     // first slice past the tile end
    hr_clear<PACKED>(s_key, s_cnt);
    if (tid == 0) s_over = 0;
    __syncthreads();

#pragma unroll
    for (int r = 0; r < HR_IPT; ++r)
        if (k[r] != HR_NONE && (k[r] >> rem_bits) != pb) hr_insert<PACKED>(s_key, s_cnt, k[r], base, 1u, flags);
    // the keys of bucket lb are a prefix of what follows the tile: first the slice loaded up front, then
    // (rarely: a bucket that holds a hot pair) slices of HR_XT keys per thread until the bucket ends
    {
        const bool more = (kx != HR_NONE) && ((kx >> rem_bits) == lb);
        hr_insert_tail<PACKED>(s_key, s_cnt, kx, base, more, flags);
        if (__syncthreads_and(more ? 1 : 0)) {
            int64_t next = t1 + HR_THREADS;
            int rounds = 0;
            while (true) {
                u64 kk[HR_XT];
#pragma unroll
                for (int q = 0; q < HR_XT; ++q) {
                    const int64_t i = next + q * HR_THREADS + tid;
                    kk[q] = (i < n) ? __ldcs(keys + i) : HR_NONE;
                }
                bool all = true;
#pragma unroll
                for (int q = 0; q < HR_XT; ++q) {
                    const bool m = (kk[q] != HR_NONE) && ((kk[q] >> rem_bits) == lb);
                    hr_insert_tail<PACKED>(s_key, s_cnt, kk[q], base, m, flags);
                    all = all && m;
                }
                if (!__syncthreads_and(all ? 1 : 0)) break;
                next += HR_XT * HR_THREADS;
                if (++rounds >= HR_MAX_TAIL_ROUNDS) {      // block-uniform: a multi-million-key bucket is better
                    if (tid == 0) atomicOr(flags, 4u);     // served by the sort path (host falls back)
                    break;
                }
            }
        }
    }
    __syncthreads();

    hr_emit<SYM, PACKED>(s_key, s_cnt, s_scan, &s_base, &s_over, base, mix, min_count, mirror, out_keys, out_count, out_n,
                         out_cap, flags);
}

// ---- the same reduce over SEVERAL bucket-sorted arrays (streamed ingest, expand.cu::count_parts_impl) ----------------
// ottocov_count_parts buckets every group of parts on its own while the next group is still crossing PCIe, so when
// the last byte has arrived there are G arrays, each sorted by the 21-bit bucket id, and all that is left is the
// counting.  A CTA owns a RANGE of `rb` consecutive buckets (~2048 keys in total over the groups); a per-group table
// of range boundaries (hr_range_bounds_kernel, built right after the group's last pass) tells it which slice of every
// array is its own, so every bucket it counts is complete and the tile-edge logic of hash_reduce_kernel is not needed.
constexpr int HR_MAX_GROUPS = 16;
struct HrGroups {
    const u64* keys[HR_MAX_GROUPS];       // bucket-sorted keys of group g
    const u32* bounds[HR_MAX_GROUPS];     // [n_ranges + 1] first key of every bucket range in that array
    int n_groups;
};

__global__ void __launch_bounds__(256) hr_fill_u32_kernel(u32* p, int64_t n, u32 v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// bounds[r] = first index whose bucket range is >= r (bounds pre-filled with n); empty ranges get their successor's start
__global__ void __launch_bounds__(256) hr_range_bounds_kernel(const u64* __restrict__ keys, int64_t n, int rem_bits, u32 rb,
                                                              u32* __restrict__ bounds, const u32* __restrict__ abort_flag) {
    if (abort_flag && (*abort_flag & HR_FLAG_FUSED_OVERFLOW)) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 r = (u32)((keys[i] >> rem_bits) / rb);
    const int64_t first = (i == 0) ? 0 : (int64_t)((keys[i - 1] >> rem_bits) / rb) + 1;
    for (int64_t q = first; q <= (int64_t)r; ++q) bounds[q] = (u32)i;       // usually one iteration, none inside a range
}

template <bool SYM, bool PACKED>
__global__ void __launch_bounds__(HR_THREADS, PACKED ? OTTOCOV_HR_MINB : 2)
hash_reduce_ranges_kernel(HrGroups grp, int rem_bits, u32 rb, KeyMix mix, u32 min_count, int mirror,
                          u64* __restrict__ out_keys, u32* __restrict__ out_count, unsigned long long* __restrict__ out_n,
                          u64 out_cap, u32* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_key = reinterpret_cast<u64*>(s_raw);                              // [HR_CAP]
    u32* s_cnt = reinterpret_cast<u32*>(s_key + HR_CAP);                     // [HR_CAP] (!PACKED)
    u32* s_scan = s_cnt + (PACKED ? 0 : HR_CAP);                             // [HR_THREADS / 32 + 1]
    __shared__ unsigned long long s_base;
    __shared__ u32 s_over;
    if (*flags & HR_FLAG_FUSED_OVERFLOW) return;
    const int tid = threadIdx.x;
    const u64 r = blockIdx.x;
    const u64 base = PACKED ? ((r * rb) << rem_bits) : 0ull;               // every key of the range is >= its first bucket
    hr_clear<PACKED>(s_key, s_cnt);
    if (tid == 0) s_over = 0;
    __syncthreads();
    u32 total = 0;
    for (int g = 0; g < grp.n_groups; ++g) {
        const u32 lo = grp.bounds[g][r], hi = grp.bounds[g][r + 1];
        total += hi - lo;
        const u64* __restrict__ k = grp.keys[g];
        for (u32 i = lo + tid; i < hi; i += HR_THREADS) hr_insert<PACKED>(s_key, s_cnt, __ldcs(k + i), base, 1u, flags);
    }
    if (total > (u32)HR_CMASK / 2) { if (tid == 0) atomicOr(flags, 4u); return; }     // count field of the packed word (block-uniform)
    __syncthreads();
    hr_emit<SYM, PACKED>(s_key, s_cnt, s_scan, &s_base, &s_over, base, mix, min_count, mirror, out_keys, out_count, out_n,
                         out_cap, flags);
}

// ---- whole buckets (round 2): one distribution pass fewer ------------------------------------------------------------
// The tile kernel above wants buckets of ~512 keys: 21 bucket bits for the headline's 742 M keys = the fused pass + TWO
// more passes of 16 B/key each.  Here a CTA takes one WHOLE bucket of ~11 k keys -- 16 bucket bits = the fused pass + ONE
// more.  The bucket's [lo, hi) comes from the bounds table the last pass wrote (radix_sort.cu, RsSegs), so there is no
// tile-edge logic.
//   * The keys are staged once in shared memory as 32-bit tags (the bits below the bucket id), each thread's keys in
//     its own column: the keys of round 0 from the top of the column, the others from the bottom.
//   * 2^round_bits rounds (0 or 1 bit): round r clears the 8192-slot table, counts the keys whose top remaining bit is
//     r, and emits.  Half a bucket per round keeps the table half empty.
//   * The insert loop is LANE-PERSISTENT: one probe per iteration, and a lane whose key is placed takes its next key in
//     the same iteration.  (Measured with ncu on the first version, which called hr_insert once per key: 12 of 32 lanes
//     active on average -- every key cost the warp the LONGEST probe chain among its 32 lanes -- and the kernel was
//     bound by issued instructions, 236 per 32 keys; experiments/round2_variants/.)  Now the warp runs for the longest
//     SUM of chains over a lane's ~11 keys, which is close to the average.
// A bucket that does not fit the staging area (a hot pair repeated thousands of times) streams from HBM once per round.
constexpr int HRB_THREADS = 512;
constexpr int HRB_CAP_LOG2 = 13;                     // 8192 slots = 64 KB; + 48 KB of staged tags: 2 CTAs / SM
constexpr int HRB_CAP = 1 << HRB_CAP_LOG2;
constexpr int HRB_KPT = 24;
constexpr int HRB_STAGE_KEYS = HRB_THREADS * HRB_KPT;  // 12288
constexpr int HRB_MAX_REM_BITS = 31;                 // tags are 32-bit
constexpr int HRB_MAX_ROUND_BITS = 1;
constexpr int HRB_MAX_ITERS = 4 * HRB_CAP;           // probes per lane and round before the CTA gives up (typical: ~16)

// bounds[nb] = n; an entry no tile wrote (~0: empty bucket) takes the next written one to its right
__global__ void __launch_bounds__(1024) hr_bounds_fix_kernel(u64* __restrict__ bounds, u32 nb, u64 n, const u32* __restrict__ flags) {
    __shared__ u64 s_min[1024];
    if (*flags & HR_FLAG_FUSED_OVERFLOW) return;
    const u32 tid = threadIdx.x;
    const u32 per = (nb + 1 + 1023) / 1024;
    const u32 b0 = min(nb + 1, tid * per), b1 = min(nb + 1, b0 + per);
    u64 m = HR_NONE;
    for (u32 i = b1; i-- > b0;) {
        const u64 v = (i == nb) ? n : bounds[i];
        m = v < m ? v : m;
    }
    s_min[tid] = m;
    __syncthreads();
    if (tid == 0) {
        u64 c = HR_NONE;
        for (int t = 1023; t >= 0; --t) { const u64 x = s_min[t]; s_min[t] = c; c = x < c ? x : c; }
    }
    __syncthreads();
    u64 c = s_min[tid];                                    // first written bound to the right of this thread's span
    for (u32 i = b1; i-- > b0;) {
        const u64 v = (i == nb) ? n : bounds[i];
        if (v == HR_NONE) bounds[i] = c;
        else { c = v; if (i == nb) bounds[i] = n; }
    }
}

__device__ __forceinline__ u32 hrb_slot_of(u32 tag) { return (tag * 0x9E3779B1u) >> (32 - HRB_CAP_LOG2); }

template <bool SYM>
__global__ void __launch_bounds__(HRB_THREADS, 2)
hash_reduce_buckets_kernel(const u64* __restrict__ keys, const u64* __restrict__ bounds, int rem_bits, int round_bits,
                           KeyMix mix, u32 min_count, int mirror, u64* __restrict__ out_keys, u32* __restrict__ out_count,
                           unsigned long long* __restrict__ out_n, u64 out_cap, u32* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_key = reinterpret_cast<u64*>(s_raw);                              // [HRB_CAP] packed words tag | count
    u32* s_tags = reinterpret_cast<u32*>(s_key + HRB_CAP);                   // [HRB_THREADS / 32 warps][HRB_KPT * 32]
    if (*flags & HR_FLAG_FUSED_OVERFLOW) return;
    const int tid = threadIdx.x;
    const u64 lo = bounds[blockIdx.x], hi = bounds[blockIdx.x + 1];
    if (hi <= lo) return;
    const u64 nk = hi - lo;
    if (nk >= (u64)HR_CMASK / 2) { if (tid == 0) atomicOr(flags, 4u); return; }      // count field of the packed word
    const u64 base = (keys[lo] >> rem_bits) << rem_bits;                     // the bucket id: the same in every key
    const u32 low_mask = (u32)((1ull << rem_bits) - 1ull);
    const int rshift = rem_bits - round_bits;
    const bool staged = nk <= (u64)HRB_STAGE_KEYS;
    // Staging: the warp's keys of round 0 are packed from the front of the warp's queue, the others from its back
    // (ballot + popc), so lane l's keys of a round are entries l, l + 32, ... : every lane gets the same number +- 1.
    const int lane = tid & 31;
    u32* wq = s_tags + (tid >> 5) * (HRB_KPT * 32);
    u32 n_front = 0, n_back = 0;                                             // warp-uniform
    if (staged) {
        const u32 n32 = (u32)nk;
        const u32* kp = reinterpret_cast<const u32*>(keys + lo + tid);      // low words only: the tag has <= 31 bits
        const u32 lt = lanemask_lt();
        u32 t[HRB_KPT];
#pragma unroll
        for (int j = 0; j < HRB_KPT; ++j)
            t[j] = ((u32)(j * HRB_THREADS + tid) < n32) ? (__ldcs(kp + 2 * j * HRB_THREADS) & low_mask) : 0u;
#pragma unroll
        for (int j = 0; j < HRB_KPT; ++j) {
            const bool front = (t[j] >> rshift) == 0u;
            if ((u32)((j + 1) * HRB_THREADS) <= n32) {                      // a full row (block-uniform): one ballot
                const u32 mf = __ballot_sync(0xffffffffu, front);
                const u32 pf = __popc(mf & lt), cf = __popc(mf);
                wq[front ? n_front + pf : HRB_KPT * 32 - 1 - (n_back + (u32)lane - pf)] = t[j];
                n_front += cf;
                n_back += 32u - cf;
            } else if ((u32)(j * HRB_THREADS) < n32) {                       // the last, partial row
                const bool valid = (u32)(j * HRB_THREADS + tid) < n32;
                const u32 mv = __ballot_sync(0xffffffffu, valid), mf = __ballot_sync(0xffffffffu, valid && front);
                const u32 pf = __popc(mf & lt), pb = __popc(mv & lt) - pf;
                if (valid) wq[front ? n_front + pf : HRB_KPT * 32 - 1 - (n_back + pb)] = t[j];
                n_front += __popc(mf);
                n_back += __popc(mv) - __popc(mf);
            }
        }
        __syncwarp();
    }
    hr_clear<true, HRB_CAP_LOG2, HRB_THREADS>(s_key, nullptr);
    __syncthreads();
    for (int r = 0; r < (1 << round_bits); ++r) {
        if (staged) {
            const u32 n_r = r == 0 ? n_front : n_back;                       // the warp's keys of this round
            const int mine = n_r > (u32)lane ? (int)((n_r - (u32)lane + 31u) >> 5) : 0;
            const u32* first = r == 0 ? wq + lane : wq + (HRB_KPT * 32 - 1 - lane);
            const int stride = r == 0 ? 32 : -32;
            // One probe per iteration; a lane whose key is placed takes its next key in the same iteration.  The vote
            // makes the warp reconverge every time round (without it the lanes that took different branches ran the
            // loop in separate groups).  CAS first, no look: an empty slot is claimed AND counted by the one atomic; the
            // tag field of the word it returns says which of the three cases this was (empty = all ones, never a tag).
            // The body is kept to ~20 instructions: no per-key probe counter -- a table that fills up (more distinct
            // keys than slots; needs an adversarial input) shows as a warp that is still busy after HRB_MAX_ITERS.
            int left = mine;
            const u32* p = first;
            u32 tag = left > 0 ? *p : 0u;
            u32 slot = hrb_slot_of(tag);
            int it = 0;
            while (__any_sync(0xffffffffu, left > 0)) {
                if (left > 0) {
                    const u64 cur = atomicCAS(reinterpret_cast<unsigned long long*>(s_key + slot), HR_NONE,
                                              ((u64)tag << HR_CB) | 1ull);
                    const u32 ct = (u32)(cur >> HR_CB);
                    if (ct == tag) atomicAdd(reinterpret_cast<u32*>(s_key + slot), 1u);
                    if (ct == tag || ct == 0xFFFFFFFFu) {
                        if (--left > 0) {
                            p += stride;
                            tag = *p;
                            slot = hrb_slot_of(tag);
                        }
                    } else {
                        slot = (slot + 1) & (HRB_CAP - 1);
                    }
                }
                if (++it >= HRB_MAX_ITERS) break;              // warp-uniform
            }
            if (__any_sync(0xffffffffu, left > 0) && lane == 0) atomicOr(flags, 1u);      // the host falls back to the sort
        } else {
            for (u64 i0 = lo; i0 < hi; i0 += HRB_THREADS) {                  // block-uniform trip count
                const u64 i = i0 + (u64)tid;
                const u64 t = (i < hi) ? (u64)((u32)__ldcs(keys + i) & low_mask) : HR_NONE;
                const bool mine = (i < hi) && ((u32)t >> rshift) == (u32)r;
                hr_insert_tail<true, HRB_CAP_LOG2>(s_key, nullptr, t, 0ull, mine, flags);
            }
        }
        __syncthreads();
        hr_emit_warp<SYM, HRB_CAP_LOG2, HRB_THREADS>(s_key, base, mix, min_count, mirror, out_keys, out_count, out_n, out_cap,
                                                     flags);                 // leaves the table empty
        if (r + 1 < (1 << round_bits)) __syncthreads();
    }
}

__global__ void __launch_bounds__(256) unmix_kernel(u64* __restrict__ keys, int64_t n, KeyMix mix) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = key_mix_inv(mix, keys[i]);
}

// plain keys (optionally with a destination stamp) -> mixed keys in place; the digit histograms of the bucket
// passes (pl) are accumulated in the same read, so the sort needs no histogram pass of its own
__global__ void __launch_bounds__(256) mix_hist_kernel(u64* __restrict__ keys, int64_t n, KeyMix mix, u64 keep_mask,
                                                       PassList pl, u64* __restrict__ ghist) {
    extern __shared__ u32 s_hist[];                   // [pl.n][RS_RADIX]
    for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += blockDim.x) s_hist[j] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 k = keys[i] & keep_mask;
        const u64 h = key_mix_fwd(mix, (u32)(k >> 32), (u32)k);
        keys[i] = h;
        for (int p = 0; p < pl.n; ++p)
            atomicAdd(&s_hist[p * RS_RADIX + ((u32)(h >> pl.shift[p]) & ((1u << pl.bits[p]) - 1u))], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < pl.n * RS_RADIX; j += blockDim.x) {
        const u32 c = s_hist[j];
        if (c) atomicAdd(&ghist[j], (u64)c);
    }
}

// ghist: device [pl.n][RS_RADIX], zeroed here; pass it to hashed_reduce as pre_hist
void mix_keys_inplace(ottocov_ctx* ctx, u64* keys, int64_t n, const KeyMix& mix, bool strip_dest, const PassList& pl,
                      u64* ghist) {
    if (n <= 0) return;
    if (pl.n > 0) CUDA_CHECK(cudaMemsetAsync(ghist, 0, (size_t)pl.n * RS_RADIX * sizeof(u64), ctx->stream));
    const int grid = (int)imin64(ceil_div64(n, 256 * 8), (int64_t)ctx->num_sms * 8);
    COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 16.0 * n, mix_hist_kernel, grid, 256, (size_t)pl.n * RS_RADIX * sizeof(u32), keys, n,
               mix, strip_dest ? 0x00FFFFFFFFFFFFFFull : ~0ull, pl, ghist);
}

bool hashed_reduce_supported(int aid_bits) { return aid_bits >= 1 && 2 * aid_bits <= 56; }

static size_t hr_smem_bytes(bool packed) { return (size_t)HR_CAP * 8 + (packed ? 0 : (size_t)HR_CAP * 4) + 32 * 4; }

// How n mixed keys of kb bits are bucketed: <= OTTOCOV_HR_AVG (default 512) keys per bucket on average.
// (A big-bucket variant that saves a pass at the price of four table rounds per bucket was measured slower in
// round 1: experiments/round1_variants/hash_reduce_big_kernel.cu.txt.)
int hashed_bucket_bits(int64_t n, int kb) {
    static int64_t avg_target = 0;
    if (!avg_target) {
        const char* e = getenv("OTTOCOV_HR_AVG");                      // tuning knob
        avg_target = e ? atoll(e) : 512;
        if (avg_target < 16 || avg_target > 2048) avg_target = 512;
    }
    int bb = 0;
    while (bb < kb && (n >> bb) > avg_target) ++bb;
    return bb;
}

// Whole-bucket reduce: <= OTTOCOV_HRB_AVG (default 11500: the registers of a CTA hold 12288 keys, and a Poisson bucket of
// 11500 stays below that by five sigma) keys per bucket on average.  Used when that takes fewer distribution passes
// than the tile kernel's small buckets (OTTOCOV_HRB=0 turns it off, OTTOCOV_HRB=2 uses it wherever it can run: tests).
static int64_t hrb_avg() {
    static int64_t v = 0;
    if (!v) { const char* e = getenv("OTTOCOV_HRB_AVG"); v = e ? atoll(e) : 11500; if (v < 16 || v > 11500) v = 11500; }
    return v;
}
static int hrb_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OTTOCOV_HRB"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
    return v;
}
static int hrb_round_bits() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OTTOCOV_HRB_ROUND_BITS"); v = e ? atoi(e) : 1; if (v < 0 || v > HRB_MAX_ROUND_BITS) v = 1; }
    return v;
}
int hashed_big_bucket_bits(int64_t n, int kb) {
    const int mode = hrb_mode();
    if (mode == 0) return 0;
    int bb = 0;
    while (bb < kb && (n >> bb) > hrb_avg()) ++bb;
    if (bb <= RS_MAX_BITS || bb > 2 * RS_MAX_BITS) return 0;           // the fused pass + exactly one more
    if (kb - bb > HRB_MAX_REM_BITS || kb - bb < hrb_round_bits()) return 0;
    const int small = hashed_bucket_bits(n, kb);
    const int p_small = (small + RS_MAX_BITS - 1) / RS_MAX_BITS;
    if (mode == 1 && p_small <= 2) return 0;                             // no pass saved
    return bb;
}

ottocov_table* hashed_reduce(ottocov_ctx* ctx, u64* keys, u64* alt, int64_t n, const KeyMix& mix, u32 min_count,
                             bool sym, bool mirror, int* passes_out, u64* pre_hist, const HashPre* pre) {
    ottocov_table* out = new ottocov_table();
    out->aid_bits = mix.ab;
    if (passes_out) *passes_out = 0;
    if (n <= 0) return out;
    if (min_count < 1) min_count = 1;
    try {
        // ---- group the keys into buckets of <= avg_target keys on average (top bits of the mixed key) ----
        static int force_fallback = -1, no_packed = 0;
        if (force_fallback < 0) {
            const char* f = getenv("OTTOCOV_HR_FORCE_FALLBACK");        // test knob: exercise the overflow path
            force_fallback = (f && atoi(f)) ? 1 : 0;
            const char* g = getenv("OTTOCOV_HR_NO_PACKED");             // test / tuning knob: wide table words
            no_packed = (g && atoi(g)) ? 1 : 0;
        }
        const int bb = pre ? pre->bb : hashed_bucket_bits(n, mix.kb);
        const int rem_bits = mix.kb - bb;
        BitField bucket_field[1] = {{rem_bits, mix.kb}};
        PassList pl = make_pass_list(bucket_field, 1);
        u64* k = keys; u64* ka = alt; u32* v = nullptr; u32* va = nullptr;
        int passes;
        const bool big = pre && pre->big;
        DevBuf<u64> bucket_bounds;
        size_t n_buckets = 0;
        if (pre) {                      // pass 0 was done by whoever wrote the keys: they sit in its digit regions
            const int b0 = pre->first_bits > 0 ? pre->first_bits : (pl.n > 0 ? pl.bits[0] : 0);
            BitField rest_field[1] = {{rem_bits + b0, mix.kb}};
            const PassList rest = make_pass_list(rest_field, 1);
            if (rest.n < 1) COV_THROW(OTTOCOV_ERR_ARG, "fused first pass needs at least one more pass");
            if (big) {
                if (rest.n != 1 || rem_bits > HRB_MAX_REM_BITS || !pre->ctr)
                    COV_THROW(OTTOCOV_ERR_ARG, "whole-bucket reduce: one pass after the fused one, <= 31 key bits below the bucket");
                n_buckets = ((size_t)1 << rest.bits[0]) * (size_t)pre->n_b;
                bucket_bounds.alloc(ctx, n_buckets + 1);
                CUDA_CHECK(cudaMemsetAsync(bucket_bounds.p, 0xFF, (n_buckets + 1) * sizeof(u64), ctx->stream));
            }
            passes = 1 + radix_sort_passes(ctx, k, ka, v, va, n, rest, pre_hist, pre->seg_cnt, pre->seg_off, pre->n_a, pre->n_b,
                                           pre->ctr ? reinterpret_cast<const u32*>(pre->ctr + 1) : nullptr, bucket_bounds.p);
        } else {
            passes = radix_sort_passes(ctx, k, ka, v, va, n, pl, pre_hist);
        }
        if (passes_out) *passes_out = passes;

        // ---- count inside the buckets ------------------------------------------------------------------------
        const u64 cap = (u64)(sym ? 2 : 1) * ((u64)n / min_count) + 1024;
        DevBuf<u64> ok(ctx, cap);
        DevBuf<u32> oc(ctx, cap);
        DevBuf<unsigned long long> ctr_own;
        unsigned long long* ctr = pre ? pre->ctr : nullptr;   // [0] rows written, [1] flags (low 32 bits)
        if (!ctr) {
            ctr_own.alloc(ctx, 2);
            ctr = ctr_own.p;
            CUDA_CHECK(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), ctx->stream));
        }
        // packed table words need the tag (mixed key relative to the tile's first bucket) to fit 42 bits: always
        // true for keys of <= 42 bits; wider keys qualify when 256 consecutive buckets fit
        const bool packed = (mix.kb <= HR_TAG_BITS || rem_bits + 8 <= HR_TAG_BITS) && !no_packed;
        const unsigned grid = (unsigned)ceil_div64(n, HR_TILE);
        u32* flags = reinterpret_cast<u32*>(ctr + 1);
        if (big) {
            COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, hr_bounds_fix_kernel, 1, 1024, 0, bucket_bounds.p, (u32)n_buckets, (u64)n,
                       (const u32*)flags);
            const size_t smem = (size_t)HRB_CAP * 8 + (size_t)HRB_STAGE_KEYS * 4;
#define HRB_LAUNCH(SYM_)                                                                                               \
            do {                                                                                                       \
                auto kern = hash_reduce_buckets_kernel<SYM_>;                                                          \
                cov_func_smem(ctx, (const void*)kern, smem);                                                           \
                COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n, kern, (unsigned)n_buckets, HRB_THREADS, smem, k,               \
                           (const u64*)bucket_bounds.p, rem_bits, hrb_round_bits(), mix, min_count,                    \
                           (SYM_ && mirror) ? 1 : 0, ok.p, oc.p, ctr, cap, flags);                                     \
            } while (0)
            if (sym) HRB_LAUNCH(true); else HRB_LAUNCH(false);
#undef HRB_LAUNCH
        } else {
#define HR_LAUNCH(SYM_, PACKED_)                                                                                      \
        do {                                                                                                          \
            auto kern = hash_reduce_kernel<SYM_, PACKED_>;                                                            \
            cov_func_smem(ctx, (const void*)kern, hr_smem_bytes(PACKED_));                                            \
            COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n, kern, grid, HR_THREADS, hr_smem_bytes(PACKED_), k, n, rem_bits,    \
                       mix, min_count, (SYM_ && mirror) ? 1 : 0, ok.p, oc.p, ctr, cap, flags);                        \
        } while (0)
        if (sym) {
            if (packed) HR_LAUNCH(true, true); else HR_LAUNCH(true, false);
        } else {
            if (packed) HR_LAUNCH(false, true); else HR_LAUNCH(false, false);
        }
#undef HR_LAUNCH
        }
        unsigned long long h[2];
        cov_readback(ctx, h, ctr, sizeof(h));
        const int64_t rows = (int64_t)h[0];
        BitField full[2] = {{0, mix.ab}, {32, 32 + mix.ab}};

        if (h[1] & (unsigned long long)HR_FLAG_FUSED_OVERFLOW) throw FusedOverflow{};
        if ((h[1] & 0xFFFFFFFFull) != 0 || force_fallback) {
            // ---- fallback: plain keys back, full sort, run-length reduce ---------------------------------------
            ok.release(); oc.release();
            COV_LAUNCH(ctx, OTTOCOV_K_MISC, 16.0 * n, unmix_kernel, (unsigned)ceil_div64(n, 256), 256, 0, k, n, mix);
            const int p2 = radix_sort_pairs(ctx, k, ka, v, va, n, full, 2);
            if (passes_out) *passes_out += p2;
            reduce_sorted(ctx, k, nullptr, n, min_count, sym, &out->keys, &out->count, &out->n);
            if (sym && mirror) {
                ottocov_table* fullt = mirror_table_impl(ctx, out, false);
                dev_free(ctx, out->keys); dev_free(ctx, out->count); delete out;
                out = fullt;
            }
            return out;
        }
        ctx->stats[OTTOCOV_K_RLE].algo_bytes += 12.0 * (double)rows;
        if (rows == 0) return out;

        if (pre && pre->skip_sort) {     // e.g. a half table on its way into ottocov_mirror_collect, which sorts the union
            DevBuf<u64> tk(ctx, rows);
            DevBuf<u32> tc(ctx, rows);
            CUDA_CHECK(cudaMemcpyAsync(tk.p, ok.p, rows * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(tc.p, oc.p, rows * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
            out->keys = tk.take(); out->count = tc.take(); out->n = rows;
            return out;
        }
        // ---- the survivors arrive in bucket order: sort them by plain key -----------------------------------
        DevBuf<u64> ok2(ctx, rows);
        DevBuf<u32> oc2(ctx, rows);
        u64* sk = ok.p; u64* ska = ok2.p; u32* sv = oc.p; u32* sva = oc2.p;
        radix_sort_pairs(ctx, sk, ska, sv, sva, rows, full, 2);
        if (sk != ok2.p) {               // result sits in the (over-sized) output buffers: keep the tight copy
            CUDA_CHECK(cudaMemcpyAsync(ok2.p, sk, rows * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(oc2.p, sv, rows * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        out->keys = ok2.take();
        out->count = oc2.take();
        out->n = rows;
    } catch (...) {
        dev_free(ctx, out->keys); dev_free(ctx, out->count);
        delete out;
        throw;
    }
    return out;
}


// ---- host side of the multi-array reduce ----------------------------------------------------------------------------------
u32 hashed_range_buckets(int64_t n_est, int bb) {
    const double per_bucket = (double)n_est / (double)((u64)1 << bb);
    double rb = (double)HR_TILE / (per_bucket > 1e-9 ? per_bucket : 1e-9);
    if (rb < 1.0) rb = 1.0;
    if (rb > 4096.0) rb = 4096.0;
    return (u32)(rb + 0.5);
}

int64_t hashed_n_ranges(int bb, u32 rb) { return (int64_t)((((u64)1 << bb) + rb - 1) / rb); }

void hashed_range_bounds(ottocov_ctx* ctx, const u64* keys, int64_t n, int rem_bits, u32 rb, int64_t n_ranges, u32* bounds,
                         const u32* abort_flag) {
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, hr_fill_u32_kernel, (unsigned)ceil_div64(n_ranges + 1, 256), 256, 0, bounds, n_ranges + 1, (u32)n);
    if (n > 0)
        COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n, hr_range_bounds_kernel, (unsigned)ceil_div64(n, 256), 256, 0, keys, n, rem_bits, rb,
                   bounds, abort_flag);
}

// keys[g]: n[g] mixed keys sorted by their bucket field [kb - bb, kb); bounds[g]: hashed_range_bounds of that array.
// ctr: device [2] (rows, flags) zeroed before the keys were made (HR_FLAG_FUSED_OVERFLOW may be set by their writer).
// Throws FusedOverflow when any flag is raised (the caller owns the fallback: it still has the events).
ottocov_table* hashed_reduce_groups(ottocov_ctx* ctx, const u64* const* keys, const u32* const* bounds, int n_groups,
                                    int64_t n_total, int bb, u32 rb, const KeyMix& mix, u32 min_count, bool sym, bool mirror,
                                    unsigned long long* ctr) {
    if (n_groups < 1 || n_groups > HR_MAX_GROUPS) COV_THROW(OTTOCOV_ERR_ARG, "1..%d bucket-sorted arrays", HR_MAX_GROUPS);
    ottocov_table* out = new ottocov_table();
    out->aid_bits = mix.ab;
    if (min_count < 1) min_count = 1;
    try {
        static int no_packed = -1;
        if (no_packed < 0) { const char* g = getenv("OTTOCOV_HR_NO_PACKED"); no_packed = (g && atoi(g)) ? 1 : 0; }
        const int rem_bits = mix.kb - bb;
        int span_bits = 0;
        while (((u64)1 << span_bits) < rb) ++span_bits;
        const bool packed = (mix.kb <= HR_TAG_BITS || rem_bits + span_bits <= HR_TAG_BITS) && !no_packed;
        const int64_t n_ranges = hashed_n_ranges(bb, rb);
        const u64 cap = (u64)(sym ? 2 : 1) * ((u64)n_total / min_count) + 1024;
        DevBuf<u64> ok(ctx, cap);
        DevBuf<u32> oc(ctx, cap);
        HrGroups grp;
        memset(&grp, 0, sizeof(grp));
        grp.n_groups = n_groups;
        for (int g = 0; g < n_groups; ++g) { grp.keys[g] = keys[g]; grp.bounds[g] = bounds[g]; }
        u32* flags = reinterpret_cast<u32*>(ctr + 1);
#define HRG_LAUNCH(SYM_, PACKED_)                                                                                      \
        do {                                                                                                           \
            auto kern = hash_reduce_ranges_kernel<SYM_, PACKED_>;                                                      \
            cov_func_smem(ctx, (const void*)kern, hr_smem_bytes(PACKED_));                                             \
            COV_LAUNCH(ctx, OTTOCOV_K_RLE, 8.0 * n_total, kern, (unsigned)n_ranges, HR_THREADS, hr_smem_bytes(PACKED_), grp, \
                       rem_bits, rb, mix, min_count, (SYM_ && mirror) ? 1 : 0, ok.p, oc.p, ctr, cap, flags);           \
        } while (0)
        if (sym) { if (packed) HRG_LAUNCH(true, true); else HRG_LAUNCH(true, false); }
        else { if (packed) HRG_LAUNCH(false, true); else HRG_LAUNCH(false, false); }
#undef HRG_LAUNCH
        unsigned long long h[2];
        cov_readback(ctx, h, ctr, sizeof(h));
        if ((h[1] & 0xFFFFFFFFull) != 0) {
            if (getenv("OTTOCOV_TRACE")) fprintf(stderr, "[trace] hashed_reduce_groups: flags 0x%llx -> fallback\n", h[1] & 0xFFFFFFFFull);
            throw FusedOverflow{};
        }
        const int64_t rows = (int64_t)h[0];
        ctx->stats[OTTOCOV_K_RLE].algo_bytes += 12.0 * (double)rows;
        if (rows == 0) return out;
        BitField full[2] = {{0, mix.ab}, {32, 32 + mix.ab}};
        DevBuf<u64> ok2(ctx, rows);
        DevBuf<u32> oc2(ctx, rows);
        u64* sk = ok.p; u64* ska = ok2.p; u32* sv = oc.p; u32* sva = oc2.p;
        radix_sort_pairs(ctx, sk, ska, sv, sva, rows, full, 2);
        if (sk != ok2.p) {
            CUDA_CHECK(cudaMemcpyAsync(ok2.p, sk, rows * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(oc2.p, sv, rows * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        out->keys = ok2.take(); out->count = oc2.take(); out->n = rows;
    } catch (...) {
        dev_free(ctx, out->keys); dev_free(ctx, out->count);
        delete out;
        throw;
    }
    return out;
}
