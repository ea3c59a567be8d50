/*
 * cov_oracle.c -- plain-C brute-force oracle for the co-event counting path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product library (libottocov.so) never calls into this file.
 *
 * PARITY UNPINNED by the reference's own tests (it has none, and its arithmetic runs inside
 * polars, which is not installable here).  Pinned by SURVEY.md App. B.1 golden vectors and by
 * agreement with oracle/ref_restatement.py (an independent dataframe-shaped restatement).
 *
 * Semantics restated (file:line into /root/reference):
 *   model/count_co_events.py:92      exact duplicate events (session,aid,ts,type) are dropped
 *   model/count_co_events.py:19      every ordered pair (i,j) of events of one session ...
 *   model/count_co_events.py:23-27   ... except the event with itself
 *   model/count_co_events.py:30-36   -86400 <= ts_j - ts_i <= 86400          (config.py:41-42)
 *   model/count_co_events.py:66-71   type_i == this, type_j in next, |dt| <= W, count per (aid_i,aid_j)
 *
 * The algorithm is deliberately the dumb one: sort rows, dedup neighbours, O(n^2) double loop
 * per session, sort the emitted keys, run-length count.  No windows-as-ranges cleverness, so it
 * shares no logic with the CUDA path it checks.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t session, ts, aid; int8_t type; } ev_t;

static int ev_cmp(const void* pa, const void* pb) {
    const ev_t* a = (const ev_t*)pa; const ev_t* b = (const ev_t*)pb;
    if (a->session != b->session) return a->session < b->session ? -1 : 1;
    if (a->ts != b->ts) return a->ts < b->ts ? -1 : 1;
    if (a->aid != b->aid) return a->aid < b->aid ? -1 : 1;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    return 0;
}

/* LSD radix sort of u64 keys, 16-bit digits. */
static int sort_u64(uint64_t* keys, int64_t n) {
    if (n < 2) return 0;
    uint64_t* tmp = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
    int64_t* hist = (int64_t*)malloc(65536 * sizeof(int64_t));
    if (!tmp || !hist) { free(tmp); free(hist); return -1; }
    uint64_t* src = keys; uint64_t* dst = tmp;
    for (int pass = 0; pass < 4; ++pass) {
        int shift = pass * 16;
        memset(hist, 0, 65536 * sizeof(int64_t));
        for (int64_t i = 0; i < n; ++i) hist[(src[i] >> shift) & 0xFFFF]++;
        int64_t run = 0;
        for (int d = 0; d < 65536; ++d) { int64_t c = hist[d]; hist[d] = run; run += c; }
        for (int64_t i = 0; i < n; ++i) dst[hist[(src[i] >> shift) & 0xFFFF]++] = src[i];
        uint64_t* t = src; src = dst; dst = t;
    }
    /* 4 passes: result is back in keys */
    free(tmp); free(hist);
    return 0;
}

typedef struct { uint64_t* v; int64_t n, cap; } vec_t;
static int vec_push(vec_t* x, uint64_t k) {
    if (x->n == x->cap) {
        int64_t nc = x->cap ? x->cap * 2 : (1 << 16);
        uint64_t* nv = (uint64_t*)realloc(x->v, (size_t)nc * sizeof(uint64_t));
        if (!nv) return -1;
        x->v = nv; x->cap = nc;
    }
    x->v[x->n++] = k;
    return 0;
}

void cov_oracle_free(void* p) { free(p); }

/*
 * Count co-events of one kind.  type_this in {0,1,2}; next_mask bit t set <=> type t is a
 * "next" type; window = W(name) in seconds.  On success returns the number U of distinct
 * (aid, aid_next) pairs and hands back three malloc'd arrays of length U sorted by
 * (aid, aid_next); *n_emitted = total ordered co-event pairs (sum of counts).  Returns -1 on
 * allocation failure.  Free the arrays with cov_oracle_free.
 */
int64_t cov_oracle_count_range(int64_t n, const int32_t* session, const int32_t* aid, const int32_t* ts,
                               const int8_t* type, int type_this, int next_mask, int64_t window,
                               int64_t dt_min, int64_t dt_max,
                               int32_t** out_aid, int32_t** out_aid_next, uint32_t** out_count,
                               int64_t* n_emitted, int64_t* n_events_after_dedup);

int64_t cov_oracle_count(int64_t n, const int32_t* session, const int32_t* aid, const int32_t* ts,
                         const int8_t* type, int type_this, int next_mask, int64_t window,
                         int32_t** out_aid, int32_t** out_aid_next, uint32_t** out_count,
                         int64_t* n_emitted, int64_t* n_events_after_dedup) {
    /* config.py:41-42: MIN_TIME_TO_NEXT = -24 h, MAX_TIME_TO_NEXT = +24 h */
    return cov_oracle_count_range(n, session, aid, ts, type, type_this, next_mask, window, -86400, 86400,
                                  out_aid, out_aid_next, out_count, n_emitted, n_events_after_dedup);
}

/* The same with the pre-filter bounds of self_merge (count_co_events.py:33-36) as parameters:
 * dt_min <= ts_j - ts_i <= dt_max (config.MIN_TIME_TO_NEXT / MAX_TIME_TO_NEXT). */
int64_t cov_oracle_count_range(int64_t n, const int32_t* session, const int32_t* aid, const int32_t* ts,
                               const int8_t* type, int type_this, int next_mask, int64_t window,
                               int64_t dt_min, int64_t dt_max,
                               int32_t** out_aid, int32_t** out_aid_next, uint32_t** out_count,
                               int64_t* n_emitted, int64_t* n_events_after_dedup) {
    *out_aid = NULL; *out_aid_next = NULL; *out_count = NULL; *n_emitted = 0;
    ev_t* ev = (ev_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(ev_t));
    if (!ev) return -1;
    for (int64_t i = 0; i < n; ++i) {
        ev[i].session = session[i]; ev[i].ts = ts[i]; ev[i].aid = aid[i]; ev[i].type = type[i];
    }
    qsort(ev, (size_t)n, sizeof(ev_t), ev_cmp);
    int64_t m = 0;                                   /* df.unique() */
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || ev_cmp(&ev[i], &ev[i - 1]) != 0) ev[m++] = ev[i];
    if (n_events_after_dedup) *n_events_after_dedup = m;

    vec_t keys = {0, 0, 0};
    for (int64_t s0 = 0; s0 < m;) {
        int64_t s1 = s0;
        while (s1 < m && ev[s1].session == ev[s0].session) ++s1;
        for (int64_t i = s0; i < s1; ++i) {
            if (ev[i].type != type_this) continue;
            for (int64_t j = s0; j < s1; ++j) {
                if (j == i) continue;                /* the event joined with itself */
                if (!((next_mask >> ev[j].type) & 1)) continue;
                int64_t dt = (int64_t)ev[j].ts - (int64_t)ev[i].ts;
                if (dt < dt_min || dt > dt_max) continue;
                int64_t adt = dt < 0 ? -dt : dt;
                if (adt > window) continue;
                uint64_t k = ((uint64_t)(uint32_t)ev[i].aid << 32) | (uint32_t)ev[j].aid;
                if (vec_push(&keys, k)) { free(ev); free(keys.v); return -1; }
            }
        }
        s0 = s1;
    }
    free(ev);
    *n_emitted = keys.n;
    if (sort_u64(keys.v, keys.n)) { free(keys.v); return -1; }
    int64_t u = 0;
    for (int64_t i = 0; i < keys.n; ++i) if (i == 0 || keys.v[i] != keys.v[i - 1]) ++u;
    int32_t* oa = (int32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(int32_t));
    int32_t* ob = (int32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(int32_t));
    uint32_t* oc = (uint32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(uint32_t));
    if (!oa || !ob || !oc) { free(oa); free(ob); free(oc); free(keys.v); return -1; }
    int64_t w = -1;
    for (int64_t i = 0; i < keys.n; ++i) {
        if (i == 0 || keys.v[i] != keys.v[i - 1]) {
            ++w; oa[w] = (int32_t)(keys.v[i] >> 32); ob[w] = (int32_t)(keys.v[i] & 0xFFFFFFFFu); oc[w] = 0;
        }
        oc[w]++;
    }
    free(keys.v);
    *out_aid = oa; *out_aid_next = ob; *out_count = oc;
    return u;
}


/* ---- EXTENSION oracle: time-decay weighted scores (no reference counterpart) -------------------------
 * Same pair set as cov_oracle_count_range; every pair contributes w = max(0.10, 1 - |dt| / window) (window 0: w = 1),
 * evaluated and summed in float64 in key order.  Returns U distinct keys sorted by (aid, aid_next) with their float64
 * score and integer count. */
typedef struct { uint64_t key; double w; } kw_t;
static int kw_cmp(const void* pa, const void* pb) {
    const kw_t* a = (const kw_t*)pa; const kw_t* b = (const kw_t*)pb;
    return a->key < b->key ? -1 : (a->key > b->key ? 1 : 0);
}

int64_t cov_oracle_score(int64_t n, const int32_t* session, const int32_t* aid, const int32_t* ts, const int8_t* type,
                         int type_this, int next_mask, int64_t window, int64_t dt_min, int64_t dt_max,
                         int32_t** out_aid, int32_t** out_aid_next, double** out_score, uint32_t** out_count) {
    *out_aid = NULL; *out_aid_next = NULL; *out_score = NULL; *out_count = NULL;
    ev_t* ev = (ev_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(ev_t));
    if (!ev) return -1;
    for (int64_t i = 0; i < n; ++i) { ev[i].session = session[i]; ev[i].ts = ts[i]; ev[i].aid = aid[i]; ev[i].type = type[i]; }
    qsort(ev, (size_t)n, sizeof(ev_t), ev_cmp);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || ev_cmp(&ev[i], &ev[i - 1]) != 0) ev[m++] = ev[i];
    kw_t* kw = NULL; int64_t nk = 0, cap = 0;
    for (int64_t s0 = 0; s0 < m;) {
        int64_t s1 = s0;
        while (s1 < m && ev[s1].session == ev[s0].session) ++s1;
        for (int64_t i = s0; i < s1; ++i) {
            if (ev[i].type != type_this) continue;
            for (int64_t j = s0; j < s1; ++j) {
                if (j == i) continue;
                if (!((next_mask >> ev[j].type) & 1)) continue;
                int64_t dt = (int64_t)ev[j].ts - (int64_t)ev[i].ts;
                if (dt < dt_min || dt > dt_max) continue;
                int64_t adt = dt < 0 ? -dt : dt;
                if (adt > window) continue;
                if (nk == cap) {
                    cap = cap ? cap * 2 : (1 << 16);
                    kw_t* nv = (kw_t*)realloc(kw, (size_t)cap * sizeof(kw_t));
                    if (!nv) { free(kw); free(ev); return -1; }
                    kw = nv;
                }
                double w = window > 0 ? 1.0 - (double)adt / (double)window : 1.0;
                if (w < 0.10) w = 0.10;
                kw[nk].key = ((uint64_t)(uint32_t)ev[i].aid << 32) | (uint32_t)ev[j].aid;
                kw[nk].w = w;
                ++nk;
            }
        }
        s0 = s1;
    }
    free(ev);
    qsort(kw, (size_t)nk, sizeof(kw_t), kw_cmp);
    int64_t u = 0;
    for (int64_t i = 0; i < nk; ++i) if (i == 0 || kw[i].key != kw[i - 1].key) ++u;
    int32_t* oa = (int32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(int32_t));
    int32_t* ob = (int32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(int32_t));
    double* os = (double*)malloc((size_t)(u > 0 ? u : 1) * sizeof(double));
    uint32_t* oc = (uint32_t*)malloc((size_t)(u > 0 ? u : 1) * sizeof(uint32_t));
    if (!oa || !ob || !os || !oc) { free(oa); free(ob); free(os); free(oc); free(kw); return -1; }
    int64_t w = -1;
    for (int64_t i = 0; i < nk; ++i) {
        if (i == 0 || kw[i].key != kw[i - 1].key) {
            ++w; oa[w] = (int32_t)(kw[i].key >> 32); ob[w] = (int32_t)(kw[i].key & 0xFFFFFFFFu); os[w] = 0.0; oc[w] = 0;
        }
        os[w] += kw[i].w; oc[w]++;
    }
    free(kw);
    *out_aid = oa; *out_aid_next = ob; *out_score = os; *out_count = oc;
    return u;
}
