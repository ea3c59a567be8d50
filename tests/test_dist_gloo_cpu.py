"""world_size-2 gloo test of the multi-GPU host logic (session sharding, partition by hash(aid),
all-to-all exchange, union of shards == single-process table).  The device kernels are replaced by
the CPU oracle here -- this checks the plumbing, the GPU parity tests check the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import small_events

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, seed, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from otto_recommender_b200.dist import exchange_records, hash_dest, shard_bounds
    s, a, t, y = small_events(seed, n_sessions=300, shuffle=False)
    sess_ids, lens = np.unique(s, return_counts=True)
    b = shard_bounds(lens, world)
    lo, hi = sess_ids[b[rank]], (sess_ids[b[rank + 1]] if b[rank + 1] < len(sess_ids) else sess_ids[-1] + 1)
    m = (s >= lo) & (s < hi)
    oa, ob, oc, _, _ = c_oracle.count_name(s[m], a[m], t[m], y[m], "click_to_click")    # local table
    dest = hash_dest(oa, world)
    order = np.argsort(dest, kind="stable")                                            # stable partition
    keys = ((oa.astype(np.int64) << 32) | ob.astype(np.int64))[order]
    cnts = oc.astype(np.int32)[order]
    rows = np.bincount(dest, minlength=world).tolist()
    rk, rc, rows_from = exchange_records(torch.from_numpy(keys), torch.from_numpy(cnts), rows)
    assert sum(rows_from) == len(rk)
    rk, rc = rk.numpy(), rc.numpy()
    assert np.all(hash_dest(rk >> 32, world) == rank)         # every received row belongs here
    uk, inv = np.unique(rk, return_inverse=True)
    sc = np.zeros(len(uk), np.int64); np.add.at(sc, inv, rc)
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), k=uk, c=sc)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_exchange_union_equals_single_process(tmp_path, world):
    from oracle import c_oracle
    seed = 11
    mp.spawn(_worker, args=(world, _free_port(), seed, str(tmp_path)), nprocs=world, join=True)
    s, a, t, y = small_events(seed, n_sessions=300, shuffle=False)
    oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, "click_to_click")
    want = dict(zip(((oa.astype(np.int64) << 32) | ob).tolist(), oc.tolist()))
    got = {}
    for r in range(world):
        z = np.load(tmp_path / f"shard{r}.npz")
        for k, c in zip(z["k"].tolist(), z["c"].tolist()):
            assert k not in got                                # shards are disjoint
            got[k] = c
    assert got == want


# ---- the fused expansion + exchange protocol (dist.ScatterExchange) with the device replaced by the oracle ----------
def _scatter_worker(rank, world, port, seed, out_dir, stripe_cap, mirror_cap):
    """ScatterExchange.count over gloo.  The engine is a stand-in that keeps every rank's receive area in a file-backed
    array visible to all processes and does the kernels' work with numpy on the plain-C oracle's local table; the
    plan arithmetic is the real ottocov_xplan_make (host code of libottocov.so).  What is under test: the first-step
    agreement, that every rank takes the same grow-and-repeat decisions from the published status words, the
    barrier discipline, and that the union of the shards is the global table."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes
    from oracle import c_oracle
    from otto_recommender_b200 import _lib
    from otto_recommender_b200.dist import ScatterExchange, hash_dest, shard_bounds
    lib = _lib.load_library()
    s, a, t, y = small_events(seed, n_sessions=300, shuffle=False)
    sess_ids, lens = np.unique(s, return_counts=True)
    b = shard_bounds(lens, world)
    lo, hi = sess_ids[b[rank]], (sess_ids[b[rank + 1]] if b[rank + 1] < len(sess_ids) else sess_ids[-1] + 1)
    m = (s >= lo) & (s < hi)
    WORDS = 1 << 16

    class FakeEngine:
        device = 0
        areas = None                                            # filled by the exchange: one int64 array per rank

        def expand_prepare(self, name, min_count=1):
            oa, ob, oc, _, _ = c_oracle.count_name(s[m], a[m], t[m], y[m], name)
            self.sym = name == "click_to_click"
            if self.sym:                                        # canonical half pairs, diagonal counted once per unordered pair
                keep = oa <= ob
                oa, ob, oc = oa[keep], ob[keep], np.where(oa[keep] == ob[keep], oc[keep] // 2, oc[keep])
            self.keys = np.repeat((oa.astype(np.int64) << 32) | ob, oc.astype(np.int64))
            return len(self.keys), self.sym

        def make_xplan(self, n_ranks, aid_bits, mx, tot, stripe_cap=0, mirror_cap=0):
            plan = _lib.XPlan()
            assert lib.ottocov_xplan_make(n_ranks, aid_bits, mx, tot, stripe_cap, mirror_cap, ctypes.byref(plan)) == 0
            return plan

        # receive area of a rank (int64 words): [0, W) status [src][4]; per source a stripe of `cap` keys and a
        # mirror stripe of `mcap` (key, count) rows -- a miniature of the ottocov_xplan layout
        def _layout(self, plan):
            cap, mcap = int(plan.stripe_cap), int(plan.mirror_cap)
            assert 16 + world * (cap + 2 * mcap) <= WORDS
            return cap, mcap, (lambda src: 16 + src * cap), (lambda src: 16 + world * cap + src * 2 * mcap)

        def expand_scatter(self, plan, rank_, peers):
            cap, mcap, koff, _ = self._layout(plan)
            dest = hash_dest(self.keys >> 32, world)
            need = max(int(np.sum(dest == d)) for d in range(world)) if len(self.keys) else 0
            for d in range(world):
                mine = self.keys[dest == d]
                area = self.areas[d]
                area[koff(rank_):koff(rank_) + min(len(mine), cap)] = mine[:cap]
                area[rank_ * 4:rank_ * 4 + 4] = [int(need > cap), need, len(mine), 0]       # status published to EVERY rank

        def reduce_received(self, plan, recv, min_count, sym):
            cap, mcap, koff, _ = self._layout(plan)
            area = self.areas[rank]
            st = area[:world * 4].reshape(world, 4)
            if st[:, 0].any() or st[:, 1].max() > cap:
                return None, int(st[:, 1].max())
            keys = np.concatenate([area[koff(src):koff(src) + st[src, 2]] for src in range(world)])
            uk, cnt = np.unique(keys, return_counts=True)
            if sym:
                cnt = np.where((uk >> 32) == (uk & 0xFFFFFFFF), 2 * cnt, cnt)
            k = cnt >= min_count
            return Tab(uk[k], cnt[k]), 0

        def mirror_push(self, plan, rank_, half, peers):
            cap, mcap, _, moff = self._layout(plan)
            off = (half.k >> 32) != (half.k & 0xFFFFFFFF)
            tk = ((half.k[off] & 0xFFFFFFFF) << 32) | (half.k[off] >> 32)
            tc = half.c[off]
            dest = hash_dest(tk >> 32, world)
            need = max(int(np.sum(dest == d)) for d in range(world)) if len(tk) else 0
            for d in range(world):
                sel = dest == d
                n = min(int(sel.sum()), mcap)
                area = self.areas[d]
                area[moff(rank_):moff(rank_) + n] = tk[sel][:n]
                area[moff(rank_) + mcap:moff(rank_) + mcap + n] = tc[sel][:n]
                area[8 + rank_ * 2:8 + rank_ * 2 + 2] = [int(sel.sum()), need]

        def mirror_collect(self, plan, rank_, half, recv):
            cap, mcap, _, moff = self._layout(plan)
            area = self.areas[rank]
            st = area[8:8 + 2 * world].reshape(world, 2)
            if st[:, 1].max() > mcap:
                return None, int(st[:, 1].max())
            ks = [half.k] + [area[moff(src):moff(src) + st[src, 0]] for src in range(world)]
            cs = [half.c] + [area[moff(src) + mcap:moff(src) + mcap + st[src, 0]] for src in range(world)]
            k, c = np.concatenate(ks), np.concatenate(cs)
            o = np.argsort(k)
            return Tab(k[o], c[o]), 0

    class Tab:
        def __init__(self, k, c): self.k, self.c = np.asarray(k, np.int64), np.asarray(c, np.int64)
        def free(self): pass

    class FileExchange(ScatterExchange):
        def _alloc(self, nbytes):                                # "symmetric memory": one file per rank, mapped by all
            dist.barrier()
            mine = np.lib.format.open_memmap(os.path.join(out_dir, f"area{self.rank}_{self.capacity}.npy"), mode="w+",
                                             dtype=np.int64, shape=(WORDS,))
            mine[:] = 0; mine.flush()
            dist.barrier()
            self.engine.areas = [np.load(os.path.join(out_dir, f"area{r}_{self.capacity}.npy"), mmap_mode="r+") for r in range(self.world)]
            self.peer_ptrs = list(range(self.world))

        def _barrier(self, channel):
            for ar in self.engine.areas:
                ar.flush()
            dist.barrier()

    eng = FakeEngine()
    ex = FileExchange(eng, device=torch.device("cpu"))
    real_make = eng.make_xplan
    first = {"done": False}

    def make_small(n_ranks, aid_bits, mx, tot, stripe_cap=0, mirror_cap=0):   # first plan far too small: forces the regrows
        if not first["done"]:
            first["done"] = True
            return real_make(n_ranks, aid_bits, mx, tot, stripe_cap or 8, mirror_cap or 4)
        return real_make(n_ranks, aid_bits, mx, tot, stripe_cap, mirror_cap)
    eng.make_xplan = make_small
    out = {}
    for name, mc in (("click_to_click", 2), ("click_to_cart_or_buy", 1)):
        first["done"] = False
        ex.plans.pop(name, None)
        tab = ex.count(name, mc, aid_bits=8)
        assert np.all(hash_dest(tab.k >> 32, world) == rank)
        out[name] = (tab.k, tab.c)
    np.savez(os.path.join(out_dir, f"scatter{rank}.npz"), regrows=ex.regrows, **{f"{n}_{x}": v[i] for n, v in out.items() for i, x in enumerate("kc")})
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_exchange_protocol_over_gloo(tmp_path):
    from oracle import c_oracle
    seed, world = 12, 2
    mp.spawn(_scatter_worker, args=(world, _free_port(), seed, str(tmp_path), 8, 4), nprocs=world, join=True)
    s, a, t, y = small_events(seed, n_sessions=300, shuffle=False)
    shards = [np.load(tmp_path / f"scatter{r}.npz") for r in range(world)]
    assert all(int(z["regrows"]) >= 2 for z in shards) and len({int(z["regrows"]) for z in shards}) == 1
    for name, mc in (("click_to_click", 2), ("click_to_cart_or_buy", 1)):
        oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, name)
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
        want = dict(zip(((ka.astype(np.int64) << 32) | kb).tolist(), kc.tolist()))
        got = {}
        for z in shards:
            for k, c in zip(z[f"{name}_k"].tolist(), z[f"{name}_c"].tolist()):
                assert k not in got
                got[k] = c
        assert got == want, name
