#!/usr/bin/env python
"""Wall-clock of the drop-in stage `python -m model.count_co_events` on a synthetic OTTO-shaped directory.

The only figures the reference publishes for this stage are the log strings "ETA 20min" (phase 1, count) and
"ETA 30min" (phase 2, merge) of model/count_co_events.py:202,210.  This tool writes the input layout the stage
expects ({DIR_DATA}/{alias}-parquet/{train,test}_sessions/*.parquet, 100k sessions per part,
etl/jsonl_to_parquet.py:23-29,81-84) and times the three phases through the module's own main().

    python tools/bench_cli.py [--train-parts 129] [--test-parts 17] [--dir /tmp/otto_cli]

One JSON line: seconds per phase, total, and the fused single-call alternative (count_population + threshold +
file-order fetch) on the same directory.  Host I/O (parquet read / write through pyarrow + pandas) is inside
every figure, exactly as it is for the reference.
"""
import argparse
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import torch

from otto_recommender_b200 import count_co_events as cce
from otto_recommender_b200.config import CoEventConfig
from otto_recommender_b200.synth import SynthSpec, generate

ap = argparse.ArgumentParser()
ap.add_argument("--train-parts", type=int, default=129)
ap.add_argument("--test-parts", type=int, default=17)
ap.add_argument("--part-sessions", type=int, default=100_000)
ap.add_argument("--dir", default="/tmp/otto_cli")
ap.add_argument("--keep", action="store_true")
args = ap.parse_args()

alias = "train-test"
root = args.dir
shutil.rmtree(root, ignore_errors=True)
t0 = time.time()
n_events = 0
first = 0
for pop, n_parts in (("train_sessions", args.train_parts), ("test_sessions", args.test_parts)):
    d = f"{root}/{alias}-parquet/{pop}"
    os.makedirs(d, exist_ok=True)
    for i in range(n_parts):
        ev = generate(SynthSpec(n_sessions=args.part_sessions, seed=1000 + first, first_session=first * args.part_sessions), "cuda")
        tab = pa.table({k: pa.array(ev[k].cpu().numpy()) for k in ("session", "aid", "ts", "type")})
        lo, hi = first * args.part_sessions, (first + 1) * args.part_sessions
        pq.write_table(tab, f"{d}/{lo:012d}_{hi:012d}.parquet")
        n_events += tab.num_rows
        first += 1
t_gen = time.time() - t0
cce.set_config(CoEventConfig(DIR_DATA=root))
phases = {}
for flags, label in (((1, 0, 0), "count"), ((0, 1, 0), "merge"), ((0, 0, 1), "merge_train_test")):
    t = time.time()
    cce.main(["--data_split_alias", alias, "--count", str(flags[0]), "--merge", str(flags[1]), "--merge_train_test", str(flags[2])])
    torch.cuda.synchronize()
    phases[label] = time.time() - t
stats = f"{root}/{alias}-counts-co-event"
rows = {n: pq.read_metadata(f"{stats}/{n}.parquet").num_rows for n in cce.config.CO_EVENTS_TO_COUNT}
# the fused alternative: one call per population, thresholds of the merge fused into the reduce
t = time.time()
eng = cce.get_engine()
fused_rows = {}
for pop in ("train_sessions", "test_sessions"):
    names = list(cce.config.CO_EVENTS_TO_COUNT)
    tabs = cce.count_population(f"{root}/{alias}-parquet/{pop}", names, [cce.config.MIN_COUNT_TO_SAVE[n] for n in names])
    for n, tb in tabs.items():
        a, b, c = tb.fetch(order="count_desc")
        fused_rows[(pop, n)] = len(a)
        tb.free()
t_fused = time.time() - t
print(json.dumps({"tool": "bench_cli", "train_parts": args.train_parts, "test_parts": args.test_parts, "events": n_events,
                  "generate_and_write_input_s": round(t_gen, 2), "phase_s": {k: round(v, 2) for k, v in phases.items()},
                  "total_s": round(sum(phases.values()), 2), "final_rows": rows,
                  "reference_published": "ETA 20min (count) + ETA 30min (merge), model/count_co_events.py:202,210",
                  "fused_count_population_both_populations_s": round(t_fused, 2),
                  "note": "wall clock incl. parquet I/O through pyarrow/pandas; the merge phases apply the reference's "
                          "row-count-triggered lossy steps exactly as configured"}))
if not args.keep:
    shutil.rmtree(root, ignore_errors=True)
