"""Multi-GPU co-event counting: sessions sharded across ranks, counts re-sharded by hash(aid).

The reference is single-process; this is the one place the path gains a
collective.  Sessions are independent (pairs never cross a session; parts never split one,
etl/jsonl_to_parquet.py:35-40) and counts are an associative integer sum keyed by (aid, aid_next),
so:

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
