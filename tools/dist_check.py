"""Multi-GPU parity check (torchrun, nccl): every rank counts its session shard, the tables are
re-sharded by hash(aid) with an all-to-all, and the union over ranks must equal the single-process
table bit for bit.  Run: torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from otto_recommender_b200 import Engine
from otto_recommender_b200.dist import (count_distributed, count_exchange_first, count_exchange_push, count_exchange_scatter,
                                        gather_table, shard_bounds, hash_dest)
from otto_recommender_b200.synth import SynthSpec, generate

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_sessions = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
d = generate(SynthSpec(n_sessions=n_sessions, seed=5), torch.device("cuda", local))
lens = torch.bincount(d["session"].long(), minlength=n_sessions).cpu().numpy()
b = shard_bounds(lens, world)
m = (d["session"] >= int(b[rank])) & (d["session"] < int(b[rank + 1]))
eng = Engine(local)
eng.load_events(*[d[k][m].contiguous() for k in ("session", "aid", "ts", "type")])
ok = True
names = ("buy_to_buy", "cart_to_buy", "cart_to_cart", "click_to_cart_or_buy", "click_to_click")   # small -> large, so the
cases = [(name, flow, mc) for name in names                                                        # symmetric receive
         for flow, mc in (("scatter", 1), ("scatter", 3), ("reduce_first", 1), ("exchange_first", 3), ("push", 3))]  # buffer must grow
for name, flow, mc in cases:
    shard = (count_distributed(eng, name) if flow == "reduce_first" else
             count_exchange_first(eng, name, mc) if flow == "exchange_first" else
             count_exchange_scatter(eng, name, mc) if flow == "scatter" else count_exchange_push(eng, name, mc))
    a, bb, c = shard.fetch()
    assert np.all(hash_dest(a, world) == rank), "row on the wrong rank"
    got = gather_table(shard)
    if rank == 0:
        if n_sessions <= 60_000:                 # small enough for the plain-C oracle: the check is against the ORACLE
            from oracle import c_oracle
            hc = [d[k].cpu().numpy() for k in ("session", "aid", "ts", "type")]
            oa, ob, oc, _, _ = c_oracle.count_name(*hc, name)
            wa, wb, wc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
        else:                                    # else against the single-GPU sort + run-length path
            ref = Engine(local)
            ref.load_events(*[d[k] for k in ("session", "aid", "ts", "type")])
            wa, wb, wc = ref.count(name, min_count=mc, symmetric=False, hashed=False).fetch()
            ref.close()
        key = got[0].astype(np.int64) << 32 | got[1]
        o = np.argsort(key)
        same = np.array_equal(got[0][o], wa) and np.array_equal(got[1][o], wb) and np.array_equal(got[2][o], wc)
        print(json.dumps({"name": name, "flow": flow, "min_count": mc, "world": world, "rows": int(len(wa)),
                          "identical": bool(same)}))
        ok &= same
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
