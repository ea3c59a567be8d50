// events.cu -- the columnar loader: rows (session, aid, ts, type) in any order -> per-type arrays
// sorted by (session, ts), exact duplicates removed.
//
// Replaces, for the GPU path, pl.read_parquet(...).unique() (model/count_co_events.py:91-92) and the
// hash-join's implicit grouping by session (:19).  Input schema: etl/jsonl_to_parquet.py:23-29.
//
// Layout produced (struct of arrays, one set per event type t in {click, cart, order}):
//   skey[t][j]   u64   (session - session_min) << 32 | (ts - ts_min)   ascending
//   aid[t][j]    u32
//   xrank[t][0/1][j] u32  number of type-(t+1)%3 / (t+2)%3 events that precede event j in the
//                         combined (session, ts) order = where a window search into that other
//                         type starts.
// Because each type is sorted by (session, ts), "same session and |dt| <= W" is one contiguous index
// range per source event: this is the CSR-by-session the expansion kernel walks.
#include "internal.cuh"
#include "scan.cuh"

struct EvStats {
    int smin, smax, tmin, tmax, amin, amax;
    unsigned int bad_type;     // some type outside 0..2
    unsigned int unsorted;     // some adjacent row pair out of (session, ts) order
    unsigned long long n_type[3];
};

__global__ void ev_stats_init_kernel(EvStats* st) {
    st->smin = st->tmin = st->amin = 2147483647;
    st->smax = st->tmax = st->amax = -2147483647 - 1;
    st->bad_type = 0; st->unsorted = 0;
    st->n_type[0] = st->n_type[1] = st->n_type[2] = 0;
}

__global__ void __launch_bounds__(256) ev_stats_kernel(const int32_t* __restrict__ session,
                                                       const int32_t* __restrict__ aid,
                                                       const int32_t* __restrict__ ts,
                                                       const int8_t* __restrict__ type, int64_t n,
                                                       EvStats* st) {
    int smin = 2147483647, smax = -2147483647 - 1, tmin = smin, tmax = smax, amin = smin, amax = smax;
    unsigned bad = 0, uns = 0;
    unsigned c0 = 0, c1 = 0, c2 = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int s = session[i], t = ts[i], a = aid[i], y = type[i];
        smin = min(smin, s); smax = max(smax, s);
        tmin = min(tmin, t); tmax = max(tmax, t);
        amin = min(amin, a); amax = max(amax, a);
        if (y < 0 || y > 2) bad = 1;
        c0 += (y == 0); c1 += (y == 1); c2 += (y == 2);
        if (i > 0) {
            const int ps = session[i - 1], pt = ts[i - 1];
            if (ps > s || (ps == s && pt > t)) uns = 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, o));
        amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        uns |= __shfl_xor_sync(0xffffffffu, uns, o);
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st->smin, smin); atomicMax(&st->smax, smax);
        atomicMin(&st->tmin, tmin); atomicMax(&st->tmax, tmax);
        atomicMin(&st->amin, amin); atomicMax(&st->amax, amax);
        if (bad) atomicOr(&st->bad_type, 1u);
        if (uns) atomicOr(&st->unsorted, 1u);
        if (c0) atomicAdd(&st->n_type[0], (unsigned long long)c0);
        if (c1) atomicAdd(&st->n_type[1], (unsigned long long)c1);
        if (c2) atomicAdd(&st->n_type[2], (unsigned long long)c2);
    }
}

// skey[i] and the identity permutation
__global__ void __launch_bounds__(256) ev_make_keys_kernel(const int32_t* __restrict__ session,
                                                           const int32_t* __restrict__ ts, int64_t n,
                                                           int smin, int tmin, u64* __restrict__ skey,
                                                           u32* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 s = (u64)((int64_t)session[i] - (int64_t)smin);
    const u64 t = (u64)((int64_t)ts[i] - (int64_t)tmin);
    skey[i] = (s << 32) | t;
    if (idx) idx[i] = (u32)i;
}

// bring aid/type into sorted order
__global__ void __launch_bounds__(256) ev_gather_kernel(const u32* __restrict__ idx,
                                                        const int32_t* __restrict__ aid,
                                                        const int8_t* __restrict__ type, int64_t n,
                                                        u32* __restrict__ aid_s, int8_t* __restrict__ type_s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 j = idx[i];
    aid_s[i] = (u32)aid[j];
    type_s[i] = type[j];
}

// Scan functor: drop exact duplicates, split by type, record cross-type ranks.
// FROM_COLUMNS: the rows are already in (session, ts) order, so the search key is built on the fly from
// the session and ts columns (no key array is ever written or re-read); otherwise `skey` holds the sorted keys.
template <bool FROM_COLUMNS>
struct SplitByType {
    static constexpr int NC = 3;
    const u64* skey;        // sorted (!FROM_COLUMNS)
    const int32_t* session; // FROM_COLUMNS
    const int32_t* ts;      // FROM_COLUMNS
    int smin, tmin;
    const u32* aid;         // in sorted order
    const int8_t* type;     // in sorted order
    TypeArray out[3];
    __device__ __forceinline__ u64 key_at(int64_t i) const {
        if (!FROM_COLUMNS) return skey[i];
        return ((u64)((int64_t)session[i] - (int64_t)smin) << 32) | (u64)((int64_t)ts[i] - (int64_t)tmin);
    }
    // An event is a duplicate iff an EARLIER row of its equal-(session, ts) run has the same aid and
    // type; runs are short (mostly 1), so the backward walk is O(1) amortised.
    __device__ bool keep(int64_t i) const {
        const u64 k = key_at(i);
        const u32 a = aid[i];
        const int8_t y = type[i];
        for (int64_t j = i - 1; j >= 0 && key_at(j) == k; --j)
            if (aid[j] == a && type[j] == y) return false;
        return true;
    }
    __device__ u64 value(int64_t i) const { return keep(i) ? (1ull << (21 * (int)type[i])) : 0ull; }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const {
        if (!v) return;
        const int t = type[i];
        // selects instead of pre[(t + k) % 3]: a dynamically indexed array would live in local memory
        const u64 p0 = pre[0], p1 = pre[1], p2 = pre[2];
        const u64 pos = t == 0 ? p0 : (t == 1 ? p1 : p2);
        out[t].skey[pos] = key_at(i);
        out[t].aid[pos] = aid[i];
        out[t].xrank[0][pos] = (u32)(t == 0 ? p1 : (t == 1 ? p2 : p0));
        out[t].xrank[1][pos] = (u32)(t == 0 ? p2 : (t == 1 ? p0 : p1));
    }
};

// The split of rows that are (optimistically) already in (session, ts) order, with the loader's range / order statistics
// accumulated while the rows are read anyway (scan.cuh, `Acc`): saves the separate statistics pass over all four columns.
struct EvAcc { int tmin, tmax, amax; unsigned bad; };      // bad: bit 0 = a row out of order, bit 1 = a negative aid
struct SplitSortedStats : SplitByType<true> {
    typedef EvAcc Acc;
    EvStats* st;
    int64_t n_rows;
    __device__ void acc_init(Acc& a) const { a.tmin = 2147483647; a.tmax = -2147483647 - 1; a.amax = -2147483647 - 1; a.bad = 0; }
    __device__ u64 value(int64_t i, Acc& a) const {
        const int s = session[i], t = ts[i], ai = (int)aid[i];
        a.tmin = min(a.tmin, t); a.tmax = max(a.tmax, t); a.amax = max(a.amax, ai);
        if (ai < 0) a.bad |= 2u;
        if (i > 0) {
            const int ps = session[i - 1], pt = ts[i - 1];
            if (ps > s || (ps == s && pt > t)) a.bad |= 1u;
        } else st->smin = s;                                   // sorted rows: the session range sits at the two ends
        if (i == n_rows - 1) st->smax = s;
        return keep(i) ? (1ull << (21 * (int)type[i])) : 0ull;
    }
    __device__ void acc_flush(Acc& a) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a.tmin = min(a.tmin, __shfl_xor_sync(0xffffffffu, a.tmin, o));
            a.tmax = max(a.tmax, __shfl_xor_sync(0xffffffffu, a.tmax, o));
            a.amax = max(a.amax, __shfl_xor_sync(0xffffffffu, a.amax, o));
            a.bad |= __shfl_xor_sync(0xffffffffu, a.bad, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&st->tmin, a.tmin); atomicMax(&st->tmax, a.tmax); atomicMax(&st->amax, a.amax);
            if (a.bad & 1u) atomicOr(&st->unsorted, 1u);
            if (a.bad & 2u) atomicMin(&st->amin, -1);
        }
    }
};

static int bit_width_u64(u64 v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// rows [base, base + n) of a functor defined on global row indices (one launch of a chained scan)
template <class F>
struct RowOffset {
    static constexpr int NC = F::NC;
    F f;
    int64_t base;
    __device__ u64 value(int64_t i) const { return f.value(i + base); }
    __device__ void apply(int64_t i, u64 v, const u64* pre) const { f.apply(i + base, v, pre); }
};

// out[0..2] += rows of type 0..2, out[3] |= some type outside 0..2
__global__ void __launch_bounds__(256) ev_type_count_kernel(const int8_t* __restrict__ type, int64_t n,
                                                            unsigned long long* __restrict__ out) {
    unsigned c0 = 0, c1 = 0, c2 = 0, bad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int y = type[i];
        c0 += (y == 0); c1 += (y == 1); c2 += (y == 2);
        if (y < 0 || y > 2) bad = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(&out[0], (unsigned long long)c0);
        if (c1) atomicAdd(&out[1], (unsigned long long)c1);
        if (c2) atomicAdd(&out[2], (unsigned long long)c2);
        if (bad) atomicOr(&out[3], 1ull);
    }
}


// ---- host columns: copy and load overlapped ----------------------------------------------------------------
// A 220 M-row load is 2.9 GB over PCIe (~52 ms) followed by ~4.5 ms of kernels.  The kernels only need rows
// that have arrived, so for host input the copy is cut into chunks on a second stream and each chunk is
// validated and split as soon as it lands:
//   * the type column goes first (1 B/row) and is counted at once: that validates every type and gives the
//     capacities of the per-type outputs before any other column has arrived;
//   * chunk c of (session, aid, ts) is copied; behind an event, the main stream runs the statistics kernel and
//     one launch of the dedup + split scan on its rows.  The scan is chained over the launches (scan.cuh
//     carry_in), so output positions and cross-type ranks are exactly those of a single scan;
//   * search keys use the fixed offset INT32_MIN for session and ts (the minima are not known yet; only
//     differences of keys matter to the window search).
// The split is optimistic: it assumes rows ordered by (session, ts).  If the statistics say otherwise (or the
// data is invalid) its outputs are dropped and the caller continues with the general path on the device copies.
constexpr int64_t EV_PIPE_MIN_ROWS = 4 << 20;
constexpr int EV_PIPE_CHUNKS = 12;

static bool load_events_pipelined(ottocov_ctx* ctx, const int32_t* h_session, const int32_t* h_aid, const int32_t* h_ts,
                                  const int8_t* h_type, int64_t n, int32_t* d_session, int32_t* d_aid, int32_t* d_ts,
                                  int8_t* d_type) {
    if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream;
    auto new_event = [&]() {
        cudaEvent_t e;
        if (!ctx->sync_events.empty()) { e = ctx->sync_events.back(); ctx->sync_events.pop_back(); }
        else CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return e;
    };
    std::vector<cudaEvent_t> used;
    struct EventReturn {
        ottocov_ctx* c; std::vector<cudaEvent_t>& v;
        ~EventReturn() { for (cudaEvent_t e : v) c->sync_events.push_back(e); }
    } event_return{ctx, used};

    // the destination blocks may have been handed over by earlier work of the main stream: order the copies behind it
    cudaEvent_t e0 = new_event(); used.push_back(e0);
    CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(cs, e0, 0));

    ctx->begin(OTTOCOV_K_LOAD);
    CUDA_CHECK(cudaMemcpyAsync(d_type, h_type, n, cudaMemcpyHostToDevice, cs));
    cudaEvent_t et = new_event(); used.push_back(et);
    CUDA_CHECK(cudaEventRecord(et, cs));
    int64_t b[EV_PIPE_CHUNKS + 1];
    cudaEvent_t ec[EV_PIPE_CHUNKS];
    for (int c = 0; c <= EV_PIPE_CHUNKS; ++c) b[c] = (n * c / EV_PIPE_CHUNKS) & ~(int64_t)31;
    b[EV_PIPE_CHUNKS] = n;
    for (int c = 0; c < EV_PIPE_CHUNKS; ++c) {
        const int64_t m = b[c + 1] - b[c];
        if (m > 0) {
            CUDA_CHECK(cudaMemcpyAsync(d_session + b[c], h_session + b[c], m * 4, cudaMemcpyHostToDevice, cs));
            CUDA_CHECK(cudaMemcpyAsync(d_aid + b[c], h_aid + b[c], m * 4, cudaMemcpyHostToDevice, cs));
            CUDA_CHECK(cudaMemcpyAsync(d_ts + b[c], h_ts + b[c], m * 4, cudaMemcpyHostToDevice, cs));
        }
        ec[c] = new_event(); used.push_back(ec[c]);
        CUDA_CHECK(cudaEventRecord(ec[c], cs));
    }
    ctx->end(OTTOCOV_K_LOAD, 13.0 * n);
    ctx->stats[OTTOCOV_K_LOAD].launches -= 1;   // copies, not kernels
    // From here on every exit path must leave the main stream behind the last copy (the caller goes on to use
    // or free the device columns on it); an exception must not return to the caller while a copy still reads its
    // host buffers.
    struct JoinCopies {
        ottocov_ctx* c; cudaEvent_t last; bool done = false;
        ~JoinCopies() {
            cudaStreamWaitEvent(c->stream, last, 0);
            if (!done) cudaStreamSynchronize(c->copy_stream);
        }
    } join_copies{ctx, ec[EV_PIPE_CHUNKS - 1]};

    // -- types: validity and per-type capacities -------------------------------------------------------------------
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, et, 0));
    DevBuf<unsigned long long> d_tc(ctx, 4);
    CUDA_CHECK(cudaMemsetAsync(d_tc.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 1.0 * n, ev_type_count_kernel, (int)imin64(ceil_div64(n, 1024), (int64_t)ctx->num_sms * 8), 256, 0,
               d_type, n, d_tc.p);
    unsigned long long tc[4];
    cov_readback(ctx, tc, d_tc.p, sizeof(tc));
    if (tc[3]) { join_copies.done = true; return false; }       // the general path reports the error

    DevBuf<u64> o_skey[3];
    DevBuf<u32> o_aid[3], o_x0[3], o_x1[3];
    SplitByType<true> f;
    f.skey = nullptr; f.session = d_session; f.ts = d_ts; f.smin = INT32_MIN; f.tmin = INT32_MIN;
    f.aid = reinterpret_cast<const u32*>(d_aid); f.type = d_type;
    for (int t = 0; t < 3; ++t) {
        const size_t cap = (size_t)tc[t];
        o_skey[t].alloc(ctx, cap); o_aid[t].alloc(ctx, cap); o_x0[t].alloc(ctx, cap); o_x1[t].alloc(ctx, cap);
        f.out[t].skey = o_skey[t].p; f.out[t].aid = o_aid[t].p;
        f.out[t].xrank[0] = o_x0[t].p; f.out[t].xrank[1] = o_x1[t].p;
        f.out[t].n = 0;
    }
    DevBuf<EvStats> d_st(ctx, 1);
    DevBuf<u64> chain(ctx, 8);                  // two sets of 3 running totals, used alternately
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, ev_stats_init_kernel, 1, 1, 0, d_st.p);
    const u64* carry = nullptr;
    int last = -1;
    for (int c = 0; c < EV_PIPE_CHUNKS; ++c) {
        const int64_t m = b[c + 1] - b[c];
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ec[c], 0));
        if (m <= 0) continue;
        const int grid = (int)imin64(ceil_div64(m, 256), (int64_t)ctx->num_sms * 16);
        // one row of overlap with the previous chunk: the order check compares every row with its predecessor
        // (the extra row only repeats values the minima / maxima have already seen)
        const int64_t o = b[c] > 0 ? b[c] - 1 : 0;
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 13.0 * m, ev_stats_kernel, grid, 256, 0, d_session + o, d_aid + o, d_ts + o, d_type + o,
                   b[c + 1] - o, d_st.p);
        RowOffset<SplitByType<true>> fo;
        fo.f = f; fo.base = b[c];
        u64* tout = chain.p + 4 * (c & 1);
        scan_apply(ctx, OTTOCOV_K_LOAD, fo, m, nullptr, 13.0 * m + 20.0 * m, carry, tout);
        carry = tout;
        last = c;
    }
    EvStats st;
    cov_readback(ctx, &st, d_st.p, sizeof(st));
    join_copies.done = true;                                       // every copy was waited for and has completed
    if (st.unsorted || st.amin < 0 || last < 0) return false;      // general path: sort first / report the error
    u64 totals[3];
    cov_readback(ctx, totals, chain.p + 4 * (last & 1), sizeof(totals));

    ottocov_events_info& info = ctx->info;
    info.session_min = st.smin; info.session_max = st.smax;
    info.ts_min = st.tmin; info.ts_max = st.tmax;
    info.aid_max = st.amax;
    info.aid_bits = bit_width_u64((u64)st.amax);
    if (info.aid_bits == 0) info.aid_bits = 1;
    info.was_sorted = 1;
    for (int t = 0; t < 3; ++t) {
        ctx->ta[t].skey = o_skey[t].take();
        ctx->ta[t].aid = o_aid[t].take();
        ctx->ta[t].xrank[0] = o_x0[t].take();
        ctx->ta[t].xrank[1] = o_x1[t].take();
        ctx->ta[t].n = (int64_t)totals[t];
        info.n_by_type[t] = (int64_t)totals[t];
    }
    info.n_events = (int64_t)(totals[0] + totals[1] + totals[2]);
    ctx->loaded = true;
    return true;
}

void free_events(ottocov_ctx* ctx) {
    for (int t = 0; t < 3; ++t) {
        dev_free(ctx, ctx->ta[t].skey);
        dev_free(ctx, ctx->ta[t].aid);
        dev_free(ctx, ctx->ta[t].xrank[0]);
        dev_free(ctx, ctx->ta[t].xrank[1]);
        ctx->ta[t] = TypeArray();
    }
    ctx->loaded = false;
}

void load_events_impl(ottocov_ctx* ctx, const int32_t* session, const int32_t* aid,
                      const int32_t* ts, const int8_t* type, int64_t n, int where) {
    free_events(ctx);
    ottocov_events_info& info = ctx->info;
    memset(&info, 0, sizeof(info));
    info.n_rows_in = n;
    if (n == 0) {
        info.was_sorted = 1;
        ctx->loaded = true;
        return;
    }
    if (n >= (int64_t)0xFFFFFFFFll) COV_THROW(OTTOCOV_ERR_ARG, "at most 2^32-2 rows per load (got %lld)", (long long)n);

    // -- columns onto the device -----------------------------------------------------------------
    DevBuf<int32_t> d_session, d_aid, d_ts;
    DevBuf<int8_t> d_type;
    static int no_pipeline = -1;
    if (no_pipeline < 0) { const char* e = getenv("OTTOCOV_NO_LOAD_PIPELINE"); no_pipeline = (e && atoi(e)) ? 1 : 0; }
    if (where == OTTOCOV_HOST && n >= EV_PIPE_MIN_ROWS && !no_pipeline) {
        d_session.alloc(ctx, n); d_aid.alloc(ctx, n); d_ts.alloc(ctx, n); d_type.alloc(ctx, n);
        if (load_events_pipelined(ctx, session, aid, ts, type, n, d_session.p, d_aid.p, d_ts.p, d_type.p)) return;
        session = d_session.p; aid = d_aid.p; ts = d_ts.p; type = d_type.p;     // all four columns have arrived
    } else if (where == OTTOCOV_HOST) {
        d_session.alloc(ctx, n); d_aid.alloc(ctx, n); d_ts.alloc(ctx, n); d_type.alloc(ctx, n);
        ctx->begin(OTTOCOV_K_LOAD);
        CUDA_CHECK(cudaMemcpyAsync(d_session.p, session, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_aid.p, aid, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_ts.p, ts, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_type.p, type, n, cudaMemcpyHostToDevice, ctx->stream));
        ctx->end(OTTOCOV_K_LOAD, 13.0 * n);
        ctx->stats[OTTOCOV_K_LOAD].launches -= 1;   // copies, not kernels
        session = d_session.p; aid = d_aid.p; ts = d_ts.p; type = d_type.p;
    }

    // -- fast path: rows already in (session, ts) order (what the ETL writes).  One tiny pass over the type column gives
    //    the per-type capacities; then ONE pass validates, finds the ranges, checks the order, removes duplicates and
    //    splits by type.  If the rows turn out unordered (or an aid is negative) its outputs are dropped and the
    //    general path below runs (and reports errors).
    static int no_fused_loader = -1;
    if (no_fused_loader < 0) { const char* e = getenv("OTTOCOV_NO_FUSED_LOADER"); no_fused_loader = (e && atoi(e)) ? 1 : 0; }
    if (!no_fused_loader) {
        DevBuf<unsigned long long> d_tc(ctx, 4);
        CUDA_CHECK(cudaMemsetAsync(d_tc.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 1.0 * n, ev_type_count_kernel, (int)imin64(ceil_div64(n, 1024), (int64_t)ctx->num_sms * 8), 256, 0,
                   type, n, d_tc.p);
        unsigned long long tc[4];
        cov_readback(ctx, tc, d_tc.p, sizeof(tc));
        if (tc[3]) COV_THROW(OTTOCOV_ERR_DATA, "event type outside {0,1,2}");
        DevBuf<u64> f_skey[3];
        DevBuf<u32> f_aid[3], f_x0[3], f_x1[3];
        DevBuf<EvStats> f_st(ctx, 1);
        COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, ev_stats_init_kernel, 1, 1, 0, f_st.p);
        SplitSortedStats f;
        f.skey = nullptr; f.session = session; f.ts = ts; f.smin = INT32_MIN; f.tmin = INT32_MIN;
        f.aid = reinterpret_cast<const u32*>(aid); f.type = type;
        f.st = f_st.p; f.n_rows = n;
        for (int t = 0; t < 3; ++t) {
            const size_t cap = (size_t)tc[t];
            f_skey[t].alloc(ctx, cap); f_aid[t].alloc(ctx, cap); f_x0[t].alloc(ctx, cap); f_x1[t].alloc(ctx, cap);
            f.out[t].skey = f_skey[t].p; f.out[t].aid = f_aid[t].p;
            f.out[t].xrank[0] = f_x0[t].p; f.out[t].xrank[1] = f_x1[t].p;
            f.out[t].n = 0;
        }
        u64 ftot[3];
        scan_apply(ctx, OTTOCOV_K_LOAD, f, n, ftot, 13.0 * n + 20.0 * n);
        EvStats fs;
        cov_readback(ctx, &fs, f_st.p, sizeof(fs));
        if (!fs.unsorted && fs.amin >= 0) {
            info.session_min = fs.smin; info.session_max = fs.smax;
            info.ts_min = fs.tmin; info.ts_max = fs.tmax;
            info.aid_max = fs.amax;
            info.aid_bits = bit_width_u64((u64)fs.amax);
            if (info.aid_bits == 0) info.aid_bits = 1;
            info.was_sorted = 1;
            for (int t = 0; t < 3; ++t) {
                ctx->ta[t].skey = f_skey[t].take();
                ctx->ta[t].aid = f_aid[t].take();
                ctx->ta[t].xrank[0] = f_x0[t].take();
                ctx->ta[t].xrank[1] = f_x1[t].take();
                ctx->ta[t].n = (int64_t)ftot[t];
                info.n_by_type[t] = (int64_t)ftot[t];
            }
            info.n_events = (int64_t)(ftot[0] + ftot[1] + ftot[2]);
            ctx->loaded = true;
            return;
        }
    }

    // -- validate + ranges + sortedness ------------------------------------------------------------
    DevBuf<EvStats> d_st(ctx, 1);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, ev_stats_init_kernel, 1, 1, 0, d_st.p);
    {
        int grid = (int)imin64(ceil_div64(n, 256), (int64_t)ctx->num_sms * 16);
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 13.0 * n, ev_stats_kernel, grid, 256, 0, session, aid, ts, type, n, d_st.p);
    }
    EvStats st;
    cov_readback(ctx, &st, d_st.p, sizeof(st));
    if (st.bad_type) COV_THROW(OTTOCOV_ERR_DATA, "event type outside {0,1,2}");
    if (st.amin < 0) COV_THROW(OTTOCOV_ERR_DATA, "negative aid %d", st.amin);
    info.session_min = st.smin; info.session_max = st.smax;
    info.ts_min = st.tmin; info.ts_max = st.tmax;
    info.aid_max = st.amax;
    info.aid_bits = bit_width_u64((u64)st.amax);
    if (info.aid_bits == 0) info.aid_bits = 1;
    info.was_sorted = st.unsorted ? 0 : 1;

    // -- (session, ts) keys; sort if the rows are not already in that order ------------------------
    DevBuf<u64> skey, skey_alt;
    DevBuf<u32> idx, idx_alt, aid_sorted;
    DevBuf<int8_t> type_sorted;
    const unsigned g1 = (unsigned)ceil_div64(n, 256);
    const u32* aid_s;
    const int8_t* type_s;
    u64* skey_p = nullptr;
    if (st.unsorted) {
        skey.alloc(ctx, n);
        idx.alloc(ctx, n); idx_alt.alloc(ctx, n); skey_alt.alloc(ctx, n);
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 20.0 * n, ev_make_keys_kernel, g1, 256, 0, session, ts, n, st.smin, st.tmin, skey.p, idx.p);
        BitField f[2];
        f[0].lo = 0;  f[0].hi = bit_width_u64((u64)((int64_t)st.tmax - (int64_t)st.tmin));
        f[1].lo = 32; f[1].hi = 32 + bit_width_u64((u64)((int64_t)st.smax - (int64_t)st.smin));
        u64* k = skey.p; u64* ka = skey_alt.p; u32* v = idx.p; u32* va = idx_alt.p;
        radix_sort_pairs(ctx, k, ka, v, va, n, f, 2);
        skey_p = k;
        aid_sorted.alloc(ctx, n); type_sorted.alloc(ctx, n);
        COV_LAUNCH(ctx, OTTOCOV_K_LOAD, 14.0 * n, ev_gather_kernel, g1, 256, 0, v, aid, type, n, aid_sorted.p, type_sorted.p);
        aid_s = aid_sorted.p; type_s = type_sorted.p;
    } else {
        aid_s = reinterpret_cast<const u32*>(aid); type_s = type;
    }

    // -- dedup + split by type (capacity = per-type row counts before dedup) -----------------------
    DevBuf<u64> o_skey[3];
    DevBuf<u32> o_aid[3], o_x0[3], o_x1[3];
    TypeArray outs[3];
    for (int t = 0; t < 3; ++t) {
        const size_t cap = (size_t)st.n_type[t];
        o_skey[t].alloc(ctx, cap); o_aid[t].alloc(ctx, cap); o_x0[t].alloc(ctx, cap); o_x1[t].alloc(ctx, cap);
        outs[t].skey = o_skey[t].p; outs[t].aid = o_aid[t].p;
        outs[t].xrank[0] = o_x0[t].p; outs[t].xrank[1] = o_x1[t].p;
        outs[t].n = 0;
    }
    u64 totals[3];
    if (st.unsorted) {
        SplitByType<false> f;
        f.skey = skey_p; f.session = nullptr; f.ts = nullptr; f.smin = st.smin; f.tmin = st.tmin;
        f.aid = aid_s; f.type = type_s;
        for (int t = 0; t < 3; ++t) f.out[t] = outs[t];
        scan_apply(ctx, OTTOCOV_K_LOAD, f, n, totals, 13.0 * n + 20.0 * n);
    } else {
        SplitByType<true> f;
        f.skey = nullptr; f.session = session; f.ts = ts; f.smin = st.smin; f.tmin = st.tmin;
        f.aid = aid_s; f.type = type_s;
        for (int t = 0; t < 3; ++t) f.out[t] = outs[t];
        scan_apply(ctx, OTTOCOV_K_LOAD, f, n, totals, 13.0 * n + 20.0 * n);
    }
    for (int t = 0; t < 3; ++t) {
        ctx->ta[t].skey = o_skey[t].take();
        ctx->ta[t].aid = o_aid[t].take();
        ctx->ta[t].xrank[0] = o_x0[t].take();
        ctx->ta[t].xrank[1] = o_x1[t].take();
        ctx->ta[t].n = (int64_t)totals[t];
        info.n_by_type[t] = (int64_t)totals[t];
    }
    info.n_events = (int64_t)(totals[0] + totals[1] + totals[2]);
    ctx->loaded = true;
}
