"""Multi-GPU co-event counting: sessions sharded across ranks, counts re-sharded by hash(aid).

The reference is single-process; this is the one place the path gains an
exchange.  Sessions are independent (pairs never cross a session; parts never split one,
etl/jsonl_to_parquet.py:35-40) and counts are an associative integer sum keyed by (aid, aid_next).

Default path (round 2), `count_exchange_scatter`: the first distribution pass of the bucketed hash reduce
runs inside the pair expansion, and its digit is (owner rank of the key, low hash-bucket bits), so the
expansion kernel stores every key straight into a stripe of its OWNER's HBM (torch symmetric memory =
CUDA IPC mappings over NVLink).  A step is: barrier - expansion/scatter + publish - barrier - remaining
passes + hash reduce on what arrived [- mirrored rows pushed the same way - barrier - one sort].  No key
touches local HBM before it crosses, no NCCL data collective, no per-destination counts on the host.

Earlier paths, kept as baselines (`bench.py --exchange nccl|push`):

    rank r:  load its session shard -> expand -> local reduce-by-key        (no communication)
             stable partition of the local table by dest = hash(aid) % R    (ottocov_table_partition)
             all-to-all of (key u64, count u32) records                      (NCCL over NVLink)
             sort + segmented sum of what it received                        (ottocov_table_from_packed)
    => rank r owns every row whose aid hashes to r; threshold and top-K are then rank-local, and
       the union over ranks equals the single-GPU table bit for bit (integer sums commute).

One process per GPU (torch.distributed, backend nccl).  The exchange itself is backend-agnostic
(`exchange_records`), which is what the world_size-2 gloo tests on CPU exercise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def hash_dest(aid: np.ndarray, n_ranks: int) -> np.ndarray:
    """numpy twin of ottocov_hash_dest (csrc/internal.cuh): dest rank of an aid."""
    h = (aid.astype(np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return (h % np.uint64(n_ranks)).astype(np.int64)


def shard_bounds(session_lengths: np.ndarray, n_ranks: int) -> np.ndarray:
    """Contiguous session ranges per rank balanced by a pair-work proxy.

    Work per session grows between n (short, window-limited) and n^2 (all events inside the
    window); n * min(n, 32) tracks the emitted pairs of OTTO-shaped sessions closely enough to
    balance ranks within a few percent.  Returns R+1 session indices.
    """
    n = np.asarray(session_lengths, dtype=np.int64)
    w = n * np.minimum(n, 32)
    c = np.concatenate([[0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, n_ranks) / n_ranks
    cuts = np.searchsorted(c, targets, side="left")
    return np.concatenate([[0], cuts, [len(n)]]).astype(np.int64)


def exchange_records(send_keys: torch.Tensor, send_cnt: torch.Tensor, rows_per_dest: Sequence[int],
                     group=None) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """All-to-all of records already grouped by destination rank.

    send_keys int64 [n], send_cnt int32 [n] (rows of dest 0 first, then dest 1, ...).
    Returns (recv_keys, recv_cnt, rows_per_source).  Works on CUDA tensors with nccl and on CPU
    tensors with gloo.
    """
    world = dist.get_world_size(group)
    assert len(rows_per_dest) == world
    dev = send_keys.device
    send_sizes = torch.tensor(list(rows_per_dest), dtype=torch.int64, device=dev)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_sizes, send_sizes, group=group)
    recv_list = [int(x) for x in recv_sizes.tolist()]
    send_list = [int(x) for x in rows_per_dest]
    n_recv = sum(recv_list)
    recv_keys = torch.empty(n_recv, dtype=send_keys.dtype, device=dev)
    recv_cnt = torch.empty(n_recv, dtype=send_cnt.dtype, device=dev)
    dist.all_to_all_single(recv_keys, send_keys, recv_list, send_list, group=group)
    dist.all_to_all_single(recv_cnt, send_cnt, recv_list, send_list, group=group)
    return recv_keys, recv_cnt, recv_list


def reshard_table(engine, table, group=None):
    """Local (aid, aid_next, count) table -> this rank's shard of the global table."""
    world = dist.get_world_size(group)
    if world == 1:
        return table
    n = table.rows
    dev = torch.device("cuda", engine.device)
    send_keys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    send_cnt = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    rows = engine.partition(table, world, send_keys.data_ptr(), send_cnt.data_ptr())
    recv_keys, recv_cnt, _ = exchange_records(send_keys[:n], send_cnt[:n], rows, group)
    return engine.table_from_packed(recv_keys, recv_cnt)


def count_distributed(engine, name: str, group=None, free_local: bool = True):
    """ottocov_count on this rank's events, then the hash(aid) exchange.  Returns this rank's shard."""
    local = engine.count(name)
    shard = reshard_table(engine, local, group)
    if shard is not local and free_local:
        local.free()
    return shard


def exchange_keys(send_keys: torch.Tensor, rows_per_dest: Sequence[int], group=None) -> torch.Tensor:
    """All-to-all of raw 8-byte co-event keys already grouped by destination rank."""
    world = dist.get_world_size(group)
    dev = send_keys.device
    send_sizes = torch.tensor(list(rows_per_dest), dtype=torch.int64, device=dev)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_sizes, send_sizes, group=group)
    recv_list = [int(x) for x in recv_sizes.tolist()]
    recv = torch.empty(sum(recv_list), dtype=send_keys.dtype, device=dev)
    dist.all_to_all_single(recv, send_keys, recv_list, [int(x) for x in rows_per_dest], group=group)
    return recv


def global_aid_bits(engine, group=None) -> int:
    """Significant aid bits over all ranks' loaded events (bounds the radix passes).  One all-reduce per
    load_events call (every rank loads the same number of times, so all ranks take the same branch)."""
    gen = engine.load_generation
    cached = getattr(engine, "_global_aid_bits", None)
    if cached is not None and cached[0] == gen:
        return cached[1]
    bits = torch.tensor([engine.events_info()["aid_bits"]], dtype=torch.int64, device=torch.device("cuda", engine.device))
    if dist.get_world_size(group) > 1:
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    engine._global_aid_bits = (gen, int(bits.item()))
    return engine._global_aid_bits[1]


def count_exchange_first(engine, name: str, min_count: int = 1, group=None, aid_bits: Optional[int] = None):
    """Exchange-before-reduce: the raw keys of this rank's sessions cross NVLink once (8 B per pair, half
    of them for symmetric kinds), then every rank sorts and reduces only the pairs it owns -- per-rank
    work is P / R, and the count threshold can be fused into the reduce because sums are complete.

    Returns this rank's shard of the global (thresholded) table: all rows whose aid hashes to it."""
    world = dist.get_world_size(group)
    dev = torch.device("cuda", engine.device)
    n_keys, sym = engine.expand_prepare(name, min_count=min_count)
    buf_a = torch.empty(max(n_keys, 1), dtype=torch.int64, device=dev)
    buf_b = torch.empty(max(n_keys, 1), dtype=torch.int64, device=dev)
    grouped, rows = engine.expand_run(world, buf_a, buf_b)
    recv = exchange_keys(grouped[:n_keys], rows, group) if world > 1 else grouped[:n_keys]
    del buf_a, buf_b, grouped
    if aid_bits is None:                 # pass it when the catalogue size is known: saves an all-reduce
        aid_bits = global_aid_bits(engine, group)
    half = engine.reduce_pairs(recv, recv.numel(), aid_bits, min_count, symmetric=sym, strip_dest=world > 1)
    if not sym:
        return half
    if world == 1:
        full = engine.mirror(half)
        half.free()
        return full
    mirrored = engine.mirror(half, transpose_only=True)       # rows (b, a, c) live on rank hash(b)
    theirs = reshard_table(engine, mirrored, group)
    full = engine.merge([half, theirs])
    for t in (half, mirrored, theirs):
        t.free()
    return full


class PushExchange:
    """Receive buffers in symmetric memory (torch.distributed._symmetric_memory): every rank's buffer is
    mapped into every other rank's address space over NVLink, so the partition kernel of
    ottocov_push_keys stores each key directly where its owner will sort it.  One instance per
    (engine, process group); grows by collective agreement when a step would not fit."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", engine.device)
        self.capacity = 0            # allocated on first use: the size must be identical on every rank

    def _alloc(self, capacity: int):
        import torch.distributed._symmetric_memory as symm_mem
        self.capacity = int(capacity)
        self.recv = symm_mem.empty(self.capacity, dtype=torch.int64, device=self.dev)
        gname = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
        self.handle = symm_mem.rendezvous(self.recv, gname)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def exchange(self, keys: torch.Tensor, n_keys: int, rows_per_dest: Sequence[int]) -> torch.Tensor:
        """keys: stamped, NOT grouped (expand_run with buf_b=None).  Returns this rank's received keys."""
        mine = torch.tensor(list(rows_per_dest), dtype=torch.int64, device=self.dev)
        allc = torch.empty(self.world * self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        m = allc.view(self.world, self.world).cpu().numpy()          # m[src, dst]
        incoming = m.sum(axis=0)
        need = int(incoming.max())
        if need > self.capacity:                                      # same decision and size on every rank
            self._alloc(int(need * 1.25) + 4096)
        offs = m[: self.rank, :].sum(axis=0)                          # my slot range in every destination
        dest_ptrs = [self.peer_ptrs[d] + int(offs[d]) * 8 for d in range(self.world)]
        self.handle.barrier(channel=0)                                # nobody still reads its receive buffer
        self.engine.push_keys(keys, n_keys, dest_ptrs)
        self.handle.barrier(channel=1)                                # every push has landed
        return self.recv[: int(incoming[self.rank])]


def count_exchange_push(engine, name: str, min_count: int = 1, group=None, aid_bits: Optional[int] = None):
    """count_exchange_first with the NCCL all-to-all replaced by the fused partition + peer-store kernel."""
    world = dist.get_world_size(group)
    if world == 1:
        return count_exchange_first(engine, name, min_count, group, aid_bits)
    dev = torch.device("cuda", engine.device)
    n_keys, sym = engine.expand_prepare(name, min_count=min_count)
    buf_a = torch.empty(max(n_keys, 1), dtype=torch.int64, device=dev)
    keys, rows = engine.expand_run(world, buf_a, None)
    cache = engine.__dict__.setdefault("_push_exchanges", {})      # lives and dies with the engine
    ex = cache.get(id(group))
    if ex is None:
        ex = cache[id(group)] = PushExchange(engine, group)
    recv = ex.exchange(keys, n_keys, rows)
    if aid_bits is None:
        aid_bits = global_aid_bits(engine, group)
    half = engine.reduce_pairs(recv, recv.numel(), aid_bits, min_count, symmetric=sym, strip_dest=True)
    if not sym:
        return half
    mirrored = engine.mirror(half, transpose_only=True)
    theirs = reshard_table(engine, mirrored, group)
    full = engine.merge([half, theirs])
    for t in (half, mirrored, theirs):
        t.free()
    return full


class ScatterExchange:
    """State of the fused expansion + exchange for one (engine, process group): the agreed plan per co-event kind
    and the receive area in symmetric memory.  Collective decisions are taken from values every rank reads
    identically (the status words each source publishes to ALL ranks), so the ranks never diverge and a step needs
    no NCCL call; the only collective outside the first step of a kind is re-allocating a larger receive area."""

    def __init__(self, engine, group=None, device=None):
        self.engine, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = device if device is not None else torch.device("cuda", engine.device)
        self.plans = {}              # name -> (plan, total_keys it was made for)
        self.capacity = 0
        self.regrows = 0

    def _ensure(self, nbytes: int):
        if nbytes <= self.capacity:
            return
        self.capacity = int(nbytes * 1.1) + (1 << 20)           # the same arithmetic on every rank
        self._alloc(self.capacity)

    def _alloc(self, nbytes: int):
        """Receive area of `nbytes` on every rank, each mapped into every other rank (collective).  Sets peer_ptrs."""
        import torch.distributed._symmetric_memory as symm_mem
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
        gname = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
        self.handle = symm_mem.rendezvous(self.buf, gname)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def _barrier(self, channel: int):
        """Stream-ordered barrier over the ranks: work enqueued before it on any rank is visible after it on all."""
        self.handle.barrier(channel=channel)

    def _agree(self, n_keys: int):
        """max and sum of the ranks' key counts (first step of a kind only: one small collective + host read-back)"""
        mine = torch.tensor([n_keys], dtype=torch.int64, device=self.dev)
        allc = torch.empty(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        v = allc.cpu()
        return int(v.max()), int(v.sum())

    def count(self, name: str, min_count: int, aid_bits: int):
        eng = self.engine
        n_keys, sym = eng.expand_prepare(name, min_count=min_count)
        entry = self.plans.get(name)
        if entry is None:
            mx, tot = self._agree(n_keys)
            entry = (eng.make_xplan(self.world, aid_bits, mx, tot), tot)
        plan, tot = entry
        while True:
            self._ensure(plan.total_bytes)
            self._barrier(0)                                  # nobody still works in its receive area
            eng.expand_scatter(plan, self.rank, self.peer_ptrs)
            self._barrier(1)                                  # every key and every status word has landed
            half, need = eng.reduce_received(plan, self.peer_ptrs[self.rank], min_count, sym)
            if half is not None:
                break
            # some stripe overflowed; `need` is the same on every rank: grow and repeat the step
            self.regrows += 1
            plan = eng.make_xplan(self.world, aid_bits, 0, tot, stripe_cap=int(need * 1.1) + 4096, mirror_cap=plan.mirror_cap)
            n_keys, sym = eng.expand_prepare(name, min_count=min_count)
        if sym:
            while True:
                eng.mirror_push(plan, self.rank, half, self.peer_ptrs)
                self._barrier(2)
                full, need = eng.mirror_collect(plan, self.rank, half, self.peer_ptrs[self.rank])
                if full is not None:
                    break
                self.regrows += 1
                plan = eng.make_xplan(self.world, aid_bits, 0, tot, stripe_cap=plan.stripe_cap, mirror_cap=int(need * 1.25) + 4096)
                self._ensure(plan.total_bytes)
                self._barrier(0)                              # the layout moved: nobody may still read the old stripes
            half.free()
            half = full
        self.plans[name] = (plan, tot)
        return half


def count_exchange_scatter(engine, name: str, min_count: int = 1, group=None, aid_bits: Optional[int] = None):
    """This rank's shard (all rows whose aid hashes to it) of the global thresholded table of one co-event kind:
    expansion fused with the first bucket pass AND the exchange (see the module docstring)."""
    world = dist.get_world_size(group)
    if world == 1:
        return engine.count(name, min_count=min_count)
    if aid_bits is None:
        aid_bits = global_aid_bits(engine, group)
    cache = engine.__dict__.setdefault("_scatter_exchanges", {})   # lives and dies with the engine
    ex = cache.get(id(group))
    if ex is None:
        ex = cache[id(group)] = ScatterExchange(engine, group)
    return ex.count(name, min_count, aid_bits)


def count_scatter_emulated(engine, shards, name: str, min_count: int, aid_bits: int, stripe_cap: int = 0,
                           mirror_cap: int = 0):
    """The same flow with R "ranks" played one after the other by ONE engine on one GPU (tests, smoke): every
    rank's receive area is a local buffer, so the peer stores of the expansion land in ordinary HBM.  `shards`
    = list of (session, aid, ts, type) column tuples, one per rank.  Returns (tables per rank, regrows)."""
    R = len(shards)
    dev = torch.device("cuda", engine.device)
    counts = []
    for cols in shards:
        engine.load_events(*cols)
        counts.append(engine.expand_prepare(name, min_count=min_count)[0])
    plan = engine.make_xplan(R, aid_bits, max(counts), sum(counts), stripe_cap=stripe_cap, mirror_cap=mirror_cap)
    regrows = 0
    while True:
        bufs = [torch.zeros(plan.total_bytes, dtype=torch.uint8, device=dev) for _ in range(R)]
        bases = [b.data_ptr() for b in bufs]
        sym = False
        for r, cols in enumerate(shards):
            engine.load_events(*cols)
            _, sym = engine.expand_prepare(name, min_count=min_count)
            engine.expand_scatter(plan, r, bases)
        halves, need = [], 0
        for r in range(R):
            h, nd = engine.reduce_received(plan, bases[r], min_count, sym)
            halves.append(h)
            need = max(need, nd)
        if need == 0:
            break
        assert all(h is None for h in halves), "every rank must see the overflow"
        regrows += 1
        plan = engine.make_xplan(R, aid_bits, 0, sum(counts), stripe_cap=int(need * 1.1) + 4096, mirror_cap=plan.mirror_cap)
    if not sym:
        return halves, regrows
    while True:
        for r in range(R):
            engine.mirror_push(plan, r, halves[r], bases)
        fulls, need = [], 0
        for r in range(R):
            f, nd = engine.mirror_collect(plan, r, halves[r], bases[r])
            fulls.append(f)
            need = max(need, nd)
        if need == 0:
            break
        assert all(f is None for f in fulls)
        regrows += 1
        # a new layout: the key stripes are no longer needed, only the mirror stripes move
        plan = engine.make_xplan(R, aid_bits, 0, sum(counts), stripe_cap=plan.stripe_cap, mirror_cap=int(need * 1.25) + 4096)
        bufs = [torch.zeros(plan.total_bytes, dtype=torch.uint8, device=dev) for _ in range(R)]
        bases = [b.data_ptr() for b in bufs]
    for h in halves:
        h.free()
    return fulls, regrows


def gather_table(table, group=None, dst: int = 0):
    """Collect every rank's shard on `dst` as numpy (aid, aid_next, count); None elsewhere."""
    a, b, c = table.fetch(order="key")
    world = dist.get_world_size(group)
    if world == 1:
        return a, b, c
    objs = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object((a, b, c), objs, dst=dst, group=group)
    if objs is None:
        return None
    return tuple(np.concatenate([o[i] for o in objs]) for i in range(3))
