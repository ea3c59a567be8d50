// api.cu -- the C ABI of libottocov.so (include/ottocov.h): argument checks, exception -> status
// translation, context life cycle, per-family launch accounting.
#include "internal.cuh"
#include <cstring>

static thread_local std::string g_create_error;

static const char* k_family_names[OTTOCOV_K_FAMILIES] = {
    "load", "window", "expand", "histogram", "sort_pass", "reduce", "filter", "topk", "order", "partition", "misc"};

void ottocov_ctx::begin(int family) {
    stats[family].launches += 1;
    if (!((profiling >> family) & 1u)) return;
    ProfEvent pe;
    pe.family = family;
    auto grab = [&]() {
        cudaEvent_t e;
        if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) { e = nullptr; }
        return e;
    };
    pe.a = grab(); pe.b = grab();
    if (pe.a) cudaEventRecord(pe.a, stream);
    prof_pending.push_back(pe);
}

void ottocov_ctx::end(int family, double algo_bytes) {
    stats[family].algo_bytes += algo_bytes;
    if (!((profiling >> family) & 1u) || prof_pending.empty()) return;
    ProfEvent& pe = prof_pending.back();
    if (pe.family == family && pe.b) cudaEventRecord(pe.b, stream);
}

#include <chrono>
void cov_trace(ottocov_ctx* ctx, const char* what) {
    static int on = -1;
    static std::chrono::steady_clock::time_point last;
    if (on < 0) { const char* e = getenv("OTTOCOV_TRACE"); on = (e && atoi(e)) ? 1 : 0; last = std::chrono::steady_clock::now(); }
    if (!on) return;
    auto t0 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(ctx->stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[trace] %-28s host %8.3f ms  (+drain %7.3f ms)\n", what,
            std::chrono::duration<double, std::milli>(t0 - last).count(),
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    last = t1;
}

void cov_readback(ottocov_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    if (bytes > 4096) COV_THROW(OTTOCOV_ERR_ARG, "read-back larger than the pinned pad");
    CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(host_dst, ctx->pinned, bytes);
}

// ---- caching allocator --------------------------------------------------------------------------------
void* cov_alloc(ottocov_ctx* ctx, size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    // best fit among cached blocks, at most 25 % (+1 MiB) larger than asked
    int best = -1;
    const size_t limit = bytes + bytes / 4 + (1u << 20);
    for (int i = 0; i < (int)ctx->cache.size(); ++i) {
        const size_t b = ctx->cache[i].bytes;
        if (b >= bytes && b <= limit && (best < 0 || b < ctx->cache[best].bytes)) best = i;
    }
    void* p = nullptr;
    size_t got = bytes;
    if (best >= 0) {
        p = ctx->cache[best].p; got = ctx->cache[best].bytes;
        ctx->cached_bytes -= got;
        ctx->cache[best] = ctx->cache.back();
        ctx->cache.pop_back();
    } else {
        cudaError_t e = cudaMallocAsync(&p, bytes, ctx->stream);
        if (e != cudaSuccess) {           // give the cache back to the driver and retry once
            (void)cudaGetLastError();
            cov_trim(ctx);
            e = cudaMallocAsync(&p, bytes, ctx->stream);
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            COV_THROW(e == cudaErrorMemoryAllocation ? OTTOCOV_ERR_NOMEM : OTTOCOV_ERR_CUDA,
                      "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    ctx->live[p] = got;
    ctx->live_bytes += got;
    if (ctx->live_bytes + ctx->cached_bytes > ctx->peak_bytes) ctx->peak_bytes = ctx->live_bytes + ctx->cached_bytes;
    return p;
}

void cov_free(ottocov_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->live.find(p);
    if (it == ctx->live.end()) { cudaFreeAsync(p, ctx->stream); return; }   // not ours: plain free
    const size_t b = it->second;
    ctx->live.erase(it);
    ctx->live_bytes -= b;
    ctx->cache.push_back(CacheBlock{p, b});
    ctx->cached_bytes += b;
}

void cov_trim(ottocov_ctx* ctx) {
    ctx->budget_cache = 0;
    for (CacheBlock& c : ctx->cache) cudaFreeAsync(c.p, ctx->stream);
    ctx->cache.clear();
    ctx->cached_bytes = 0;
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    (void)cudaGetLastError();
}

static void resolve_profile(ottocov_ctx* ctx) {
    if (ctx->prof_pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (ProfEvent& pe : ctx->prof_pending) {
        float ms = 0.f;
        if (pe.a && pe.b && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) ctx->stats[pe.family].ms += ms;
        else (void)cudaGetLastError();
        if (pe.a) ctx->event_pool.push_back(pe.a);
        if (pe.b) ctx->event_pool.push_back(pe.b);
    }
    ctx->prof_pending.clear();
}

#define API_BEGIN(ctx_)                                                          \
    if (!(ctx_)) return OTTOCOV_ERR_ARG;                                         \
    try {                                                                        \
        CUDA_CHECK(cudaSetDevice((ctx_)->device));

#define API_END(ctx_)                                                            \
        return OTTOCOV_OK;                                                       \
    } catch (const CovError& e) {                                                \
        (ctx_)->err = e.msg;                                                     \
        (void)cudaGetLastError();                                                \
        return e.code;                                                           \
    } catch (const std::bad_alloc&) {                                            \
        (ctx_)->err = "host allocation failed";                                  \
        return OTTOCOV_ERR_NOMEM;                                                \
    } catch (...) {                                                              \
        (ctx_)->err = "unknown internal error";                                  \
        return OTTOCOV_ERR_CUDA;                                                 \
    }

extern "C" {

int ottocov_version(void) { return OTTOCOV_VERSION; }

const char* ottocov_kernel_family_name(int family) {
    return (family >= 0 && family < OTTOCOV_K_FAMILIES) ? k_family_names[family] : "?";
}

int ottocov_create(int device, ottocov_ctx** out) {
    if (!out) return OTTOCOV_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (libottocov has no CPU fallback)";
        (void)cudaGetLastError();
        return OTTOCOV_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) {
        g_create_error = "device index out of range";
        return OTTOCOV_ERR_ARG;
    }
    ottocov_ctx* ctx = nullptr;
    try {
        ctx = new ottocov_ctx();
        ctx->device = device;
        memset(ctx->stats, 0, sizeof(ctx->stats));
        memset(&ctx->info, 0, sizeof(ctx->info));
        memset(&ctx->last_count, 0, sizeof(ctx->last_count));
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) COV_THROW(OTTOCOV_ERR_CUDA, "device %d is sm_%d%d; libottocov is built for sm_100a only", device, prop.major, prop.minor);
        ctx->num_sms = prop.multiProcessorCount;
        CUDA_CHECK(cudaHostAlloc(&ctx->pinned, 4096, cudaHostAllocDefault));
        // keep freed blocks in the stream-ordered pool: the pipeline re-allocates the same sizes
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    } catch (const CovError& e2) {
        g_create_error = e2.msg;
        delete ctx;
        (void)cudaGetLastError();
        return e2.code;
    } catch (...) {
        g_create_error = "unknown error in ottocov_create";
        delete ctx;
        return OTTOCOV_ERR_CUDA;
    }
    *out = ctx;
    return OTTOCOV_OK;
}

int ottocov_destroy(ottocov_ctx* ctx) {
    if (!ctx) return OTTOCOV_OK;
    cudaSetDevice(ctx->device);
    free_events(ctx);
    free_topk(ctx);
    free_popularity(ctx);
    free_features(ctx);
    free_plan(ctx);
    if (ctx->sweep_status) cudaFreeAsync(ctx->sweep_status, ctx->stream);
    if (ctx->scan_status) cudaFreeAsync(ctx->scan_status, ctx->stream);
    for (auto& kv : ctx->live) cudaFreeAsync(kv.first, ctx->stream);    // tables the caller never freed
    ctx->live.clear();
    cov_trim(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->sweep_ticket) cudaFree(ctx->sweep_ticket);
    if (ctx->scan_ticket) cudaFree(ctx->scan_ticket);
    if (ctx->scan_totals) cudaFree(ctx->scan_totals);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
    for (ProfEvent& pe : ctx->prof_pending) { if (pe.a) cudaEventDestroy(pe.a); if (pe.b) cudaEventDestroy(pe.b); }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->sync_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    (void)cudaGetLastError();
    delete ctx;
    return OTTOCOV_OK;
}

const char* ottocov_last_error(const ottocov_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int ottocov_set_stream(ottocov_ctx* ctx, void* cuda_stream) {
    API_BEGIN(ctx)
    if ((cudaStream_t)cuda_stream != ctx->stream) {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // pool frees are ordered on the old stream
        ctx->stream = (cudaStream_t)cuda_stream;
    }
    API_END(ctx)
}

int ottocov_trim(ottocov_ctx* ctx) {
    API_BEGIN(ctx)
    cov_trim(ctx);
    API_END(ctx)
}

int ottocov_memory_info(ottocov_ctx* ctx, int64_t* live_bytes, int64_t* cached_bytes, int64_t* peak_bytes) {
    API_BEGIN(ctx)
    if (live_bytes) *live_bytes = (int64_t)ctx->live_bytes;
    if (cached_bytes) *cached_bytes = (int64_t)ctx->cached_bytes;
    if (peak_bytes) *peak_bytes = (int64_t)ctx->peak_bytes;
    API_END(ctx)
}

int ottocov_synchronize(ottocov_ctx* ctx) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

int ottocov_set_profiling(ottocov_ctx* ctx, int on) {
    API_BEGIN(ctx)
    resolve_profile(ctx);
    ctx->profiling = on == 1 ? 0xFFFFFFFFu : (unsigned)on;      // 1 = every family, else a bit mask << 0
    API_END(ctx)
}

int ottocov_kernel_stats(ottocov_ctx* ctx, ottocov_kernel_stat* out, int reset) {
    API_BEGIN(ctx)
    resolve_profile(ctx);
    if (out) memcpy(out, ctx->stats, sizeof(ctx->stats));
    if (reset) memset(ctx->stats, 0, sizeof(ctx->stats));
    API_END(ctx)
}

int ottocov_load_events(ottocov_ctx* ctx, const int32_t* session, const int32_t* aid, const int32_t* ts,
                        const int8_t* type, int64_t n, int where) {
    API_BEGIN(ctx)
    if (n < 0) COV_THROW(OTTOCOV_ERR_ARG, "n < 0");
    if (n > 0 && (!session || !aid || !ts || !type)) COV_THROW(OTTOCOV_ERR_ARG, "NULL column");
    if (where != OTTOCOV_HOST && where != OTTOCOV_DEVICE) COV_THROW(OTTOCOV_ERR_ARG, "bad `where`");
    ctx->info_only = false;
    load_events_impl(ctx, session, aid, ts, type, n, where);
    API_END(ctx)
}

int ottocov_count_parts(ottocov_ctx* ctx, int n_parts, const int32_t* const* session, const int32_t* const* aid,
                        const int32_t* const* ts, const int8_t* const* type, const int64_t* rows,
                        const ottocov_spec* specs, int n_specs, ottocov_table** tables_out) {
    API_BEGIN(ctx)
    if (n_parts < 0 || n_specs < 1 || !specs || !tables_out) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    if (n_parts > 0 && (!session || !aid || !ts || !type || !rows)) COV_THROW(OTTOCOV_ERR_ARG, "NULL part table");
    ctx->info_only = false;
    try {
        count_parts_impl(ctx, n_parts, session, aid, ts, type, rows, specs, n_specs, tables_out);
    } catch (...) {
        for (int k = 0; k < n_specs; ++k)
            if (tables_out[k]) { dev_free(ctx, tables_out[k]->keys); dev_free(ctx, tables_out[k]->count); delete tables_out[k]; tables_out[k] = nullptr; }
        free_events(ctx);
        throw;
    }
    API_END(ctx)
}

int ottocov_get_events_info(ottocov_ctx* ctx, ottocov_events_info* out) {
    API_BEGIN(ctx)
    if (!out) COV_THROW(OTTOCOV_ERR_ARG, "NULL out");
    if (!ctx->loaded && !ctx->info_only) COV_THROW(OTTOCOV_ERR_STATE, "no events loaded");
    *out = ctx->info;
    API_END(ctx)
}

int ottocov_count(ottocov_ctx* ctx, const ottocov_spec* spec, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!spec || !out) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out = count_impl(ctx, spec);
    API_END(ctx)
}

int ottocov_expand_prepare(ottocov_ctx* ctx, const ottocov_spec* spec, int64_t* n_keys, int* symmetric) {
    API_BEGIN(ctx)
    if (!spec || !n_keys || !symmetric) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    expand_prepare_impl(ctx, spec, n_keys, symmetric);
    API_END(ctx)
}

int ottocov_expand_run(ottocov_ctx* ctx, int n_ranks, uint64_t* buf_a_dev, uint64_t* buf_b_dev, int* result_in_b,
                       int64_t* rows_per_dest) {
    API_BEGIN(ctx)
    if (!result_in_b || !rows_per_dest) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    expand_run_impl(ctx, n_ranks, (u64*)buf_a_dev, (u64*)buf_b_dev, result_in_b, rows_per_dest);
    API_END(ctx)
}

int ottocov_push_keys(ottocov_ctx* ctx, const uint64_t* keys_dev, int64_t n, int n_ranks, const uint64_t* dest_ptrs) {
    API_BEGIN(ctx)
    if (n < 0 || !dest_ptrs || (n > 0 && !keys_dev)) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    push_keys_impl(ctx, (const u64*)keys_dev, n, n_ranks, (const u64*)dest_ptrs);
    API_END(ctx)
}

int ottocov_xplan_make(int n_ranks, int aid_bits, int64_t max_local_keys, int64_t total_keys, int64_t stripe_cap,
                       int64_t mirror_cap, ottocov_xplan* out) {
    if (!out) return OTTOCOV_ERR_ARG;
    try {
        xplan_make_impl(n_ranks, aid_bits, max_local_keys, total_keys, stripe_cap, mirror_cap, out);
        return OTTOCOV_OK;
    } catch (const CovError& e) {
        g_create_error = e.msg;
        return e.code;
    }
}

int ottocov_expand_scatter(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const uint64_t* peer_base) {
    API_BEGIN(ctx)
    if (!plan || !peer_base) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    expand_scatter_impl(ctx, plan, rank, (const u64*)peer_base);
    API_END(ctx)
}

int ottocov_reduce_received(ottocov_ctx* ctx, const ottocov_xplan* plan, uint64_t recv_area_dev, uint32_t min_count,
                            int symmetric, ottocov_table** out, int64_t* need_cap) {
    API_BEGIN(ctx)
    if (!plan || !out || !need_cap || !recv_area_dev) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    const int64_t n_pairs = ctx->last_count.n_pairs;
    *out = reduce_received_impl(ctx, plan, (u64)recv_area_dev, min_count, symmetric, need_cap);
    ctx->last_count.n_pairs = n_pairs;
    API_END(ctx)
}

int ottocov_mirror_push(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                        const uint64_t* peer_base) {
    API_BEGIN(ctx)
    if (!plan || !half || !peer_base) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    mirror_push_impl(ctx, plan, rank, half, (const u64*)peer_base);
    API_END(ctx)
}

int ottocov_mirror_collect(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                           uint64_t recv_area_dev, ottocov_table** out, int64_t* need_rows) {
    API_BEGIN(ctx)
    if (!plan || !half || !out || !need_rows || !recv_area_dev) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out = mirror_collect_impl(ctx, plan, rank, half, (u64)recv_area_dev, need_rows);
    ctx->last_count.n_unique = *out ? (*out)->n : 0;
    API_END(ctx)
}

int ottocov_reduce_pairs(ottocov_ctx* ctx, uint64_t* keys_dev, int64_t n, int aid_bits, uint32_t min_count,
                         int symmetric, int strip_dest, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!out || n < 0 || (n > 0 && !keys_dev)) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    *out = nullptr;
    *out = reduce_pairs_impl(ctx, (u64*)keys_dev, n, aid_bits, min_count, symmetric, strip_dest);
    API_END(ctx)
}

int ottocov_table_mirror(ottocov_ctx* ctx, const ottocov_table* t, int transpose_only, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!t || !out) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out = mirror_table_impl(ctx, t, transpose_only != 0);
    API_END(ctx)
}

int ottocov_get_count_info(ottocov_ctx* ctx, ottocov_count_info* out) {
    API_BEGIN(ctx)
    if (!out) COV_THROW(OTTOCOV_ERR_ARG, "NULL out");
    *out = ctx->last_count;
    API_END(ctx)
}

int ottocov_table_from_arrays(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next, const uint32_t* count,
                              int64_t n, int where, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!out || n < 0 || (n > 0 && (!aid || !aid_next || !count))) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    *out = nullptr;
    *out = table_from_arrays_impl(ctx, aid, aid_next, count, n, where);
    API_END(ctx)
}

int ottocov_table_from_packed(ottocov_ctx* ctx, const uint64_t* keys, const uint32_t* count, int64_t n, int where,
                              ottocov_table** out) {
    API_BEGIN(ctx)
    if (!out || n < 0 || (n > 0 && (!keys || !count))) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    *out = nullptr;
    *out = table_from_packed_impl(ctx, (const u64*)keys, count, n, where);
    API_END(ctx)
}

int ottocov_table_free(ottocov_ctx* ctx, ottocov_table* t) {
    API_BEGIN(ctx)
    if (t) {
        dev_free(ctx, t->keys);
        dev_free(ctx, t->count);
        delete t;
    }
    API_END(ctx)
}

int ottocov_table_rows(const ottocov_table* t, int64_t* n_rows) {
    if (!t || !n_rows) return OTTOCOV_ERR_ARG;
    *n_rows = t->n;
    return OTTOCOV_OK;
}

int ottocov_table_total(ottocov_ctx* ctx, const ottocov_table* t, int64_t* sum_of_counts) {
    API_BEGIN(ctx)
    if (!t || !sum_of_counts) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *sum_of_counts = table_total_impl(ctx, t);
    API_END(ctx)
}

int ottocov_table_merge(ottocov_ctx* ctx, ottocov_table* const* tabs, int n_tabs, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!out || n_tabs < 0 || (n_tabs > 0 && !tabs)) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    *out = nullptr;
    *out = merge_tables_impl(ctx, tabs, n_tabs);
    API_END(ctx)
}

int ottocov_table_filter(ottocov_ctx* ctx, const ottocov_table* t, uint32_t min_count, ottocov_table** out) {
    API_BEGIN(ctx)
    if (!t || !out) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out = filter_table_impl(ctx, t, min_count);
    API_END(ctx)
}

int ottocov_table_fetch(ottocov_ctx* ctx, const ottocov_table* t, int order, int64_t head, int32_t* aid,
                        int32_t* aid_next, int32_t* count, int64_t cap, int where, int64_t* n_out) {
    API_BEGIN(ctx)
    if (!t) COV_THROW(OTTOCOV_ERR_ARG, "NULL table");
    if (cap > 0 && (!aid || !aid_next || !count)) COV_THROW(OTTOCOV_ERR_ARG, "NULL output");
    fetch_table_impl(ctx, t, order, head, aid, aid_next, count, cap, where, n_out);
    API_END(ctx)
}

int ottocov_table_device_ptrs(const ottocov_table* t, const uint64_t** keys, const uint32_t** count) {
    if (!t) return OTTOCOV_ERR_ARG;
    if (keys) *keys = (const uint64_t*)t->keys;
    if (count) *count = t->count;
    return OTTOCOV_OK;
}

int ottocov_table_topk(ottocov_ctx* ctx, const ottocov_table* t, int k, int64_t* n_aids) {
    API_BEGIN(ctx)
    if (!t) COV_THROW(OTTOCOV_ERR_ARG, "NULL table");
    topk_impl(ctx, t, k);
    if (n_aids) *n_aids = ctx->topk_n;
    API_END(ctx)
}

int ottocov_topk_fetch(ottocov_ctx* ctx, int32_t* aid_x, int32_t* n_valid, int32_t* aid_y, int32_t* cnt,
                       int64_t cap_aids, int where) {
    API_BEGIN(ctx)
    if (ctx->topk_k == 0) COV_THROW(OTTOCOV_ERR_STATE, "ottocov_topk_fetch before ottocov_table_topk");
    const int64_t n = ctx->topk_n;
    const int k = ctx->topk_k;
    if (cap_aids < n) COV_THROW(OTTOCOV_ERR_CAPACITY, "top-k fetch needs room for %lld aids", (long long)n);
    if (n > 0) {
        if (!aid_x || !n_valid || !aid_y || !cnt) COV_THROW(OTTOCOV_ERR_ARG, "NULL output");
        const cudaMemcpyKind kind = (where == OTTOCOV_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        CUDA_CHECK(cudaMemcpyAsync(aid_x, ctx->topk_aid_x, n * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(n_valid, ctx->topk_nvalid, n * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(aid_y, ctx->topk_aid_y, n * k * 4, kind, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(cnt, ctx->topk_cnt, n * k * 4, kind, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

int ottocov_topk_lookup(ottocov_ctx* ctx, const int32_t* aids, int64_t n, int where, int32_t* n_valid, int32_t* aid_y,
                        int32_t* cnt) {
    API_BEGIN(ctx)
    if (n < 0 || (n > 0 && (!aids || !n_valid || !aid_y || !cnt))) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    topk_lookup_impl(ctx, aids, n, where, n_valid, aid_y, cnt);
    API_END(ctx)
}

int ottocov_count_features(ottocov_ctx* ctx, const int32_t* aid, const int32_t* aid_next, const int32_t* count, int64_t n,
                           int where, int first_n, int64_t quantile_row, int64_t* n_rows) {
    API_BEGIN(ctx)
    if (n < 0 || (n > 0 && (!aid || !aid_next || !count))) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    if (where != OTTOCOV_HOST && where != OTTOCOV_DEVICE) COV_THROW(OTTOCOV_ERR_ARG, "bad `where`");
    count_features_impl(ctx, aid, aid_next, count, n, where, first_n, quantile_row);
    if (n_rows) *n_rows = ctx->feat_n;
    API_END(ctx)
}

int ottocov_count_features_fetch(ottocov_ctx* ctx, int32_t* aid, int32_t* aid_next, int32_t* count, int16_t* count_pop,
                                 int16_t* perc_pop, int16_t* rank, int8_t* count_rel, int64_t cap_rows, int where) {
    API_BEGIN(ctx)
    count_features_fetch_impl(ctx, aid, aid_next, count, count_pop, perc_pop, rank, count_rel, cap_rows, where);
    API_END(ctx)
}

int ottocov_count_weighted(ottocov_ctx* ctx, const ottocov_spec* spec, ottocov_wtable** out) {
    API_BEGIN(ctx)
    if (!spec || !out) COV_THROW(OTTOCOV_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out = count_weighted_impl(ctx, spec);
    API_END(ctx)
}

int ottocov_wtable_rows(const ottocov_wtable* t, int64_t* n_rows) {
    if (!t || !n_rows) return OTTOCOV_ERR_ARG;
    *n_rows = t->n;
    return OTTOCOV_OK;
}

int ottocov_wtable_free(ottocov_ctx* ctx, ottocov_wtable* t) {
    API_BEGIN(ctx)
    if (t) { dev_free(ctx, t->keys); dev_free(ctx, t->count); dev_free(ctx, t->score_fx); delete t; }
    API_END(ctx)
}

int ottocov_wtable_fetch(ottocov_ctx* ctx, const ottocov_wtable* t, int32_t* aid, int32_t* aid_next, double* score,
                         int32_t* count, int64_t cap, int where, int64_t* n_out) {
    API_BEGIN(ctx)
    if (!t) COV_THROW(OTTOCOV_ERR_ARG, "NULL table");
    if (cap > 0 && (!aid || !aid_next || !score)) COV_THROW(OTTOCOV_ERR_ARG, "NULL output");
    wtable_fetch_impl(ctx, t, aid, aid_next, score, count, cap, where, n_out);
    API_END(ctx)
}

int ottocov_wtable_topk(ottocov_ctx* ctx, const ottocov_wtable* t, int k, int32_t* aid, int32_t* aid_next, double* score,
                        int32_t* rank, int64_t cap, int where, int64_t* n_out) {
    API_BEGIN(ctx)
    if (!t) COV_THROW(OTTOCOV_ERR_ARG, "NULL table");
    if (cap > 0 && (!aid || !aid_next || !score || !rank)) COV_THROW(OTTOCOV_ERR_ARG, "NULL output");
    wtable_topk_impl(ctx, t, k, aid, aid_next, score, rank, cap, where, n_out);
    API_END(ctx)
}

int ottocov_count_popularity(ottocov_ctx* ctx, const int32_t* cluster, const int32_t* aid, const int32_t* ts,
                             const int8_t* type, int64_t n, int where, int32_t ts_recent, int keep_top_k, int64_t* n_rows) {
    API_BEGIN(ctx)
    if (n < 0 || keep_top_k < 1) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    if (n > 0 && (!cluster || !aid || !ts || !type)) COV_THROW(OTTOCOV_ERR_ARG, "NULL column");
    if (where != OTTOCOV_HOST && where != OTTOCOV_DEVICE) COV_THROW(OTTOCOV_ERR_ARG, "bad `where`");
    if (n >= (int64_t)0xFFFFFFFFll) COV_THROW(OTTOCOV_ERR_ARG, "at most 2^32-2 rows per call");
    count_popularity_impl(ctx, cluster, aid, ts, type, n, where, ts_recent, keep_top_k);
    if (n_rows) *n_rows = ctx->pop_n;
    API_END(ctx)
}

int ottocov_popularity_fetch(ottocov_ctx* ctx, int32_t* aid, int32_t* cluster, int16_t* ranks, int64_t cap_rows, int where) {
    API_BEGIN(ctx)
    popularity_fetch_impl(ctx, aid, cluster, ranks, cap_rows, where);
    API_END(ctx)
}

int ottocov_table_partition(ottocov_ctx* ctx, const ottocov_table* t, int n_ranks, uint64_t* keys_out_dev,
                            uint32_t* count_out_dev, int64_t* rows_per_dest) {
    API_BEGIN(ctx)
    if (!t || !rows_per_dest || n_ranks < 1 || n_ranks > 1024) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    if (t->n > 0 && (!keys_out_dev || !count_out_dev)) COV_THROW(OTTOCOV_ERR_ARG, "NULL send buffer");
    partition_table_impl(ctx, t, n_ranks, (u64*)keys_out_dev, count_out_dev, rows_per_dest);
    if (t->n == 0) for (int r = 0; r < n_ranks; ++r) rows_per_dest[r] = 0;
    API_END(ctx)
}

uint32_t ottocov_hash_dest(uint32_t aid, uint32_t n_ranks) {
    uint32_t h = aid * 0x9E3779B1u;
    h ^= h >> 15;
    return n_ranks ? h % n_ranks : 0;
}

uint64_t ottocov_key_mix(int aid_bits, uint32_t aid, uint32_t aid_next) {
    if (aid_bits < 1 || aid_bits > 28) return ~0ull;
    const KeyMix m = make_key_mix(aid_bits);
    return key_mix_fwd(m, aid, aid_next);
}

uint64_t ottocov_key_unmix(int aid_bits, uint64_t mixed) {
    if (aid_bits < 1 || aid_bits > 28) return ~0ull;
    const KeyMix m = make_key_mix(aid_bits);
    return key_mix_inv(m, mixed);
}

int ottocov_sort_u64(ottocov_ctx* ctx, uint64_t* keys_dev, uint32_t* vals_dev, int64_t n, int lo_bit, int hi_bit) {
    API_BEGIN(ctx)
    if (n < 0 || lo_bit < 0 || hi_bit > 64 || lo_bit > hi_bit) COV_THROW(OTTOCOV_ERR_ARG, "bad argument");
    if (n > 1 && hi_bit > lo_bit) {
        if (!keys_dev) COV_THROW(OTTOCOV_ERR_ARG, "NULL keys");
        DevBuf<u64> alt(ctx, n);
        DevBuf<u32> valt;
        if (vals_dev) valt.alloc(ctx, n);
        u64* k = (u64*)keys_dev; u64* ka = alt.p; u32* v = vals_dev; u32* va = valt.p;
        BitField f[1] = {{lo_bit, hi_bit}};
        radix_sort_pairs(ctx, k, ka, v, va, n, f, 1);
        if (k != (u64*)keys_dev) {
            CUDA_CHECK(cudaMemcpyAsync(keys_dev, k, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            if (vals_dev) CUDA_CHECK(cudaMemcpyAsync(vals_dev, v, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    API_END(ctx)
}

}  // extern "C"
