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
