"""Throughput of the popularity stage on the synthetic OTTO shape.

    python tools/bench_popularity.py [--sessions 12900000] [--steps 5] [--clusters 50]

Prints one JSON line: events/s through ottocov_count_popularity with the event columns resident in HBM
(general popularity = one cluster, and `--clusters` pseudo-clusters), the per-family kernel times, and the CPU
restatement (oracle/popularity_oracle.py, pandas) timed on a bounded sample of the same events."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from otto_recommender_b200 import Engine
from otto_recommender_b200.synth import SynthSpec, generate

ap = argparse.ArgumentParser()
ap.add_argument("--sessions", type=int, default=12_900_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--clusters", type=int, default=50)
ap.add_argument("--cpu-sample-events", type=int, default=20_000_000)
args = ap.parse_args()

dev = torch.device("cuda", 0)
d = generate(SynthSpec(n_sessions=args.sessions, n_aids=1_800_000, seed=42), dev)
s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
n = int(s.numel())
g = torch.Generator(device="cuda"); g.manual_seed(1)
cl_of_session = torch.randint(-1, args.clusters, (args.sessions,), generator=g, device=dev, dtype=torch.int32)
cols = {1: torch.zeros(n, dtype=torch.int32, device=dev), args.clusters: cl_of_session[s.long()].contiguous()}
ts_recent = int(t.max().item()) - 7 * 86400
eng = Engine(0)
out = {"metric": "events/s (popularity counts + 6 ordinal ranks per cluster, top-20 kept)", "unit": "events/s", "events": n,
       "data": "synthetic", "runs": {}}
for ncl, cl in cols.items():
    for _ in range(2):
        r = eng.count_popularity(cl, a, t, y, ts_recent=ts_recent, keep_top_k=20)
    eng.kernel_stats(reset=True)
    eng.set_profiling(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = eng.count_popularity(cl, a, t, y, ts_recent=ts_recent, keep_top_k=20)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    st = eng.kernel_stats(reset=True)
    eng.set_profiling(False)
    out["runs"][f"cl{ncl}"] = {"ms_per_call": ms, "events_per_s": n / (ms * 1e-3), "rows_kept": int(len(r["aid"])),
                               "kernels_ms_per_call": {k: v["ms"] / args.steps for k, v in st.items() if v["launches"]},
                               # pack reads 13 B + writes 8 B per event; each sort pass 16 B per event
                               "algo_GBps_sort_pass": (st["sort_pass"]["algo_bytes"] / (st["sort_pass"]["ms"] * 1e-3) / 1e9)
                               if st["sort_pass"]["ms"] > 0 else None}
# CPU restatement on a bounded sample
from oracle import popularity_oracle as po
m = min(n, args.cpu_sample_events)
hc = [x[:m].cpu().numpy() for x in (cols[args.clusters], a, t, y)]
t0 = time.perf_counter()
po.popularity_ranks_frame(*hc, keep_top_k=20, ts_recent=ts_recent)
dt = time.perf_counter() - t0
out["cpu_baseline"] = {"value": m / dt, "unit": "events/s", "cores": 1, "kind": "port",
                       "sample": f"first {m:,} events, {args.clusters} clusters, oracle/popularity_oracle.py (pandas), {dt:.1f} s"}
print(json.dumps(out))
