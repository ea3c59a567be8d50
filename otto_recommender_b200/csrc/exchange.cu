// exchange.cu -- the second, small exchange of the multi-GPU path: mirrored rows of symmetric kinds.
//
// Symmetric kinds (click_to_click, cart_to_cart, buy_to_buy) are expanded as canonical keys (min aid, max aid), so the
// reduced, thresholded HALF table of rank r holds rows (a, b, c), a <= b, hash(a) % R == r.  The full table also has
// the transposed rows (b, a, c), and they belong to rank hash(b) % R (every row of one aid lives on one rank, so the
// per-aid top-K stays rank-local).  Round 1 re-sharded them with a partition pass + three NCCL all-to-alls + a host
// read-back of the split sizes + two merges.  Here the rows are stored straight into their owner's receive stripe
// (peer memory over NVLink), the per-source row counts are published next to them, and after the ranks synchronise
// each rank concatenates what it received with its own half rows and sorts once.  No reference counterpart: the
// reference is single-process; the table it must reproduce is concat_files_w_stats' (count_co_events.py:168-175).
#include "internal.cuh"

__global__ void __launch_bounds__(256) mirror_push_kernel(const u64* __restrict__ keys, const u32* __restrict__ count, int64_t n,
                                                          PeerBases pb, ottocov_xplan plan, int rank,
                                                          unsigned long long* __restrict__ mcur) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    u64 k = 0;
    u32 c = 0;
    bool act = i < n;
    if (act) { k = keys[i]; c = count[i]; }
    const u32 a = (u32)(k >> 32), b = (u32)k;
    act = act && a != b;                                        // the diagonal is its own transpose
    const u32 dest = act ? hash_dest(b, (u32)plan.n_ranks) : 0xFFFFFFFFu;
    const u32 peers = __match_any_sync(0xffffffffu, dest);      // one reservation per destination per warp
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (act && lane == leader) base = atomicAdd(&mcur[dest], (unsigned long long)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!act) return;
    const u64 slot = base + __popc(peers & lanemask_lt());
    if (slot >= (u64)plan.mirror_cap) return;                   // counted, not stored: the collect step asks for more room
    const u64 peer = pb.p[dest];
    reinterpret_cast<u64*>(peer + plan.off_mkeys)[(size_t)rank * plan.mirror_cap + slot] = ((u64)b << 32) | (u64)a;
    reinterpret_cast<u32*>(peer + plan.off_mcnt)[(size_t)rank * plan.mirror_cap + slot] = c;
}

__global__ void publish_mirror_kernel(PeerBases pb, ottocov_xplan plan, int rank, const unsigned long long* __restrict__ mcur) {
    const int dest = threadIdx.x;
    if (dest >= plan.n_ranks) return;
    u64 need = 0;
    for (int d = 0; d < plan.n_ranks; ++d) need = mcur[d] > need ? mcur[d] : need;
    u64* st = reinterpret_cast<u64*>(pb.p[dest] + plan.off_mstatus) + (size_t)rank * 4;
    st[0] = mcur[dest];
    st[1] = need;
    st[2] = 0; st[3] = 0;
}

void mirror_push_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half, const u64* peer_base_host) {
    const int R = plan->n_ranks;
    if (R < 2 || R > XCH_MAX_RANKS || rank < 0 || rank >= R) COV_THROW(OTTOCOV_ERR_ARG, "bad rank / n_ranks");
    PeerBases pb;
    memset(&pb, 0, sizeof(pb));
    for (int d = 0; d < R; ++d) pb.p[d] = peer_base_host[d];
    DevBuf<unsigned long long> mcur(ctx, XCH_MAX_RANKS);
    CUDA_CHECK(cudaMemsetAsync(mcur.p, 0, XCH_MAX_RANKS * sizeof(unsigned long long), ctx->stream));
    if (half->n > 0)
        COV_LAUNCH(ctx, OTTOCOV_K_PARTITION, 24.0 * half->n, mirror_push_kernel, (unsigned)ceil_div64(half->n, 256), 256, 0,
                   half->keys, half->count, half->n, pb, *plan, rank, mcur.p);
    COV_LAUNCH(ctx, OTTOCOV_K_MISC, 0, publish_mirror_kernel, 1, XCH_MAX_RANKS, 0, pb, *plan, rank, mcur.p);
}

ottocov_table* mirror_collect_impl(ottocov_ctx* ctx, const ottocov_xplan* plan, int rank, const ottocov_table* half,
                                   u64 recv_area, int64_t* need_rows) {
    (void)rank;
    const int R = plan->n_ranks;
    *need_rows = 0;
    u64 st[XCH_MAX_RANKS * 4];
    cov_readback(ctx, st, reinterpret_cast<const void*>(recv_area + plan->off_mstatus), (size_t)R * 4 * 8);
    u64 need = 0;
    int64_t total = half->n;
    for (int r = 0; r < R; ++r) { need = st[r * 4 + 1] > need ? st[r * 4 + 1] : need; total += (int64_t)st[r * 4]; }
    if (need > (u64)plan->mirror_cap) { *need_rows = (int64_t)need; return nullptr; }
    ottocov_table* out = new ottocov_table();
    out->aid_bits = plan->aid_bits;
    if (total == 0) return out;
    try {
        DevBuf<u64> k(ctx, total), ka(ctx, total);
        DevBuf<u32> v(ctx, total), va(ctx, total);
        int64_t at = 0;
        if (half->n > 0) {
            CUDA_CHECK(cudaMemcpyAsync(k.p, half->keys, half->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(v.p, half->count, half->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            at = half->n;
        }
        for (int r = 0; r < R; ++r) {
            const int64_t m = (int64_t)st[r * 4];
            if (m == 0) continue;
            CUDA_CHECK(cudaMemcpyAsync(k.p + at, reinterpret_cast<const u64*>(recv_area + plan->off_mkeys) + (size_t)r * plan->mirror_cap,
                                       m * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_CHECK(cudaMemcpyAsync(v.p + at, reinterpret_cast<const u32*>(recv_area + plan->off_mcnt) + (size_t)r * plan->mirror_cap,
                                       m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            at += m;
        }
        // half rows (a <= b) and transposed rows (b > a) never collide, and every canonical key lives on one rank:
        // the keys are distinct, one sort finishes the table
        BitField fields[2] = {{0, plan->aid_bits}, {32, 32 + plan->aid_bits}};
        u64* pk = k.p; u64* pka = ka.p; u32* pv = v.p; u32* pva = va.p;
        radix_sort_pairs(ctx, pk, pka, pv, pva, total, fields, 2);
        if (pk == k.p) { out->keys = k.take(); out->count = v.take(); }
        else { out->keys = ka.take(); out->count = va.take(); }
        out->n = total;
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}
