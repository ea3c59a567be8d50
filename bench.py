#!/usr/bin/env python
"""bench.py -- co-event pairs/s of the B200 co-visitation counting path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--sessions S]

Workloads (BASELINE.json configs; synthetic OTTO shape from otto_recommender_b200/synth.py):
    cooc      configs[1] (default): click_to_click 12 h, top-20, min_count 10 on 12.9 M sessions / 220 M events / 1.8 M aids
    all5      configs[2]: all five co-event kinds of config.CO_EVENTS_TO_COUNT with their MIN_COUNT_TO_SAVE thresholds
    longtail  configs[3]: click_to_cart_or_buy 24 h with one 465-click session forced in (README.md:18), integer counts
    decay     configs[3] with the float time-decay EXTENSION (w = max(0.10, 1 - |dt|/W) per pair; no reference counterpart), N = 1
    scale4    configs[4]: 4x scale, 51.6 M sessions / 880 M events / 7.2 M aids (46-bit keys), click_to_click
    popularity  SURVEY 8(f) rank 3 (count_popularity), one GPU, its own JSON line
One "step" = one pass of the hot path over that batch:

    raw event columns in HBM -> loader (order check / sort, dedup, split by type) -> per kind: window ranges ->
    pair expansion fused with the first bucket pass (keys go through a bijective mix and are scattered into one
    region per hash digit) -> the remaining distribution passes on the hash bits -> bucketed hash reduce in shared
    memory with the threshold and the symmetric mirror fused in -> key sort of the survivors -> segmented top-20

`value`  = emitted co-event pairs / device time, inputs resident in HBM (CUDA events, max over ranks).
`e2e`    = same metric through the public Python API with HOST (pinned) event columns in and the
           top-20 + thresholded pair table copied back to host inside the timed region.
N > 1    = strong scaling: the same events, sessions range-sharded over the ranks, keys re-sharded by hash(aid) over
           NVLink (otto_recommender_b200/dist.py).
`config.fingerprint` = order-independent fingerprint of the GLOBAL thresholded tables (rows, sum of counts, two 64-bit
           hash sums over (aid, aid_next, count)), all-reduced: the same at every N by construction of the path.
`config.parity_sample` (N = 1) = the GPU tables + top-20 on the CPU baseline's sample compared with the oracle's.
--impl reference = the reference's CPU algorithm for the SAME work the GPU arm does (pyarrow restatement of
           model/count_co_events.py -- polars itself is not installable here), all host threads, on a bounded
           sample: four fixed parts of the workload plus their merge, threshold and top-20 per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "pairs/s"
TOP_K = 20
FULL_SESSIONS = 12_900_000
MIN_COUNT_TO_SAVE = {"click_to_click": 10, "click_to_cart_or_buy": 5, "cart_to_cart": 2, "cart_to_buy": 2, "buy_to_buy": 2}
ALL_NAMES = ["click_to_click", "click_to_cart_or_buy", "cart_to_cart", "cart_to_buy", "buy_to_buy"]

WORKLOADS = {
    # name: (metric label, kinds, sessions, aids, synth extras, BASELINE config index)
    "cooc": ("co-event pairs/sec (click-to-click 12h, top-20)", ["click_to_click"], FULL_SESSIONS, 1_800_000, {}, 1),
    "all5": ("co-event pairs/sec (all five co-event kinds, top-20)", ALL_NAMES, FULL_SESSIONS, 1_800_000, {}, 2),
    "longtail": ("co-event pairs/sec (click-to-cart-or-buy 24h, long-tail sessions, top-20)", ["click_to_cart_or_buy"],
                 FULL_SESSIONS, 1_800_000, {"force_long_click_session": 465}, 3),
    "scale4": ("co-event pairs/sec (click-to-click 12h, 4x scale, top-20)", ["click_to_click"], 4 * FULL_SESSIONS, 7_200_000, {}, 4),
    # EXTENSION (the reference has no weighting): configs[3] with float time-decay scores, own f64 oracle
    "decay": ("co-event pairs/sec (time-decay weighted click-to-cart-or-buy 24h, long-tail sessions, top-20)",
              ["click_to_cart_or_buy"], FULL_SESSIONS, 1_800_000, {"force_long_click_session": 465}, 3),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlSampler:
    """SM clock + throttle reasons read in-process through NVML (what nvidia-smi itself calls), every 100 ms while
    the timed region runs.  (A separate `nvidia-smi -lms` process was measured in round 1 to stall kernel launches and
    stream synchronisations on some hosts; three light NVML calls per sample do not.)  Without pynvml the clocks
    are reported as unavailable -- there is no silent fallback to the perturbing sampler."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int = 0):
        self.gpu = gpu_index
        self.rows = []
        self.t0 = self.t1 = None
        self.ok = False
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            self._stop.wait(0.1)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "source": "nvml"}
        self._stop.set()
        self.thread.join(timeout=1)
        rows = [r for r in self.rows if self.t0 is None or (self.t0 - 0.05 <= r[0] <= (self.t1 or r[0]) + 0.05)]
        if not rows:
            rows = self.rows[-3:]
        sm = sorted(r[1] for r in rows)
        reasons = set()
        for r in rows:
            for bit, name in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml (pynvml, 100 ms)"}


def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# =====================================================================================================
# reference arm: the reference's CPU algorithm on the host cores, like for like with the GPU arm
# =====================================================================================================
def _ref_count_names(df_merged, names):
    """count_co_events (model/count_co_events.py:60-77) restricted to `names`: the same filter + group-by per kind."""
    import pyarrow as pa
    import pyarrow.compute as pc
    from oracle import ref_restatement as rr
    out = {}
    for name in names:
        type_this, types_next = rr.MAP_NAME_COUNT_TYPE[name]
        mask = pc.and_(pc.and_(pc.equal(df_merged["type"], type_this),
                               pc.is_in(df_merged["type_next"], value_set=pa.array(types_next, pa.int8()))),
                       pc.less_equal(pc.abs(df_merged["time_to_next"]), rr.MAP_MAX_TIME_TO_NEXT[name]))
        g = df_merged.filter(mask).group_by(["aid", "aid_next"], use_threads=True).aggregate([("aid_next", "count")])
        out[name] = pa.table({"aid": g["aid"], "aid_next": g["aid_next"], "count": pc.cast(g["aid_next_count"], pa.uint32())})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc
    from oracle import c_oracle
    from oracle import ref_restatement as rr
    from otto_recommender_b200.synth import SynthSpec, generate_numpy

    metric, names, sessions, n_aids, extra, cfg_idx = WORKLOADS[args.workload]
    cores = _host_threads()
    pa.set_cpu_count(cores)                       # torchrun exports OMP_NUM_THREADS=1, which pyarrow would obey
    pa.set_io_thread_count(max(2, min(cores, 8)))
    # The reference cuts its input into parts of 100k sessions (etl/jsonl_to_parquet.py:35-40) and self-joins each in
    # slices of 10k sessions (count_co_events.py:41-57).  The sample keeps the 10k-session slices and the multi-part
    # merge but uses smaller parts, so that K + W steps of ~10 s each end within a few minutes.
    part_sessions = args.ref_part_sessions
    n_parts = args.ref_parts
    parts = []
    for i in range(n_parts):                      # fixed parts: session blocks of the workload's generator, seeds 42 + i
        d = generate_numpy(SynthSpec(n_sessions=part_sessions, n_aids=n_aids, seed=42 + i, first_session=i * part_sessions,
                                     **({} if i else extra)))
        parts.append((d["session"], d["aid"], d["ts"], d["type"]))
    n_events = sum(len(p[0]) for p in parts)

    def step():
        # Exactly what one GPU step does, with the reference's own structure: per part read -> unique() -> self-join in
        # 10k-session slices -> +-24 h filter (count_co_events.py:91-93, 17-57), the filter + group-by of the kinds the
        # workload counts (:60-77), then the merge over the parts with the threshold (:103-181) and the per-aid top-20
        # (retrieve.py:41-51).
        per_name = {n: [] for n in names}
        pairs = 0
        for cols in parts:
            df = rr.unique_events(rr.events_table(*cols))
            res = _ref_count_names(rr.self_merge_big_df(df), names)
            for n in names:
                per_name[n].append(res[n])
                pairs += int(pc.sum(res[n]["count"]).as_py() or 0)
        for n in names:
            t = rr.merge_counts(n, per_name[n], exact=True, min_count_to_save=MIN_COUNT_TO_SAVE[n])
            rr.top_n_per_aid(t, TOP_K)
        return pairs

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(args.steps):
        pairs += step()
    dt = time.perf_counter() - t0
    value = pairs / dt

    # the plain-C oracle (one core) on one of the same parts, and its ideal multi-core extrapolation: the restated
    # dataframe pipeline materialises the n^2 join like the reference does, so a tight C loop on ONE core beats it
    t1 = time.perf_counter()
    c_pairs = 0
    for n in names:
        oa, ob, oc, emitted, _ = c_oracle.count_name(*parts[0], n)
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=MIN_COUNT_TO_SAVE[n])
        c_oracle.top_n(ka, kb, kc, TOP_K)
        c_pairs += emitted
    c_dt = time.perf_counter() - t1
    sample = (f"{n_parts} fixed {part_sessions // 1000}k-session parts ({n_events:,} events) of the workload per step: per part unique + self-join "
              f"(10k-session slices) + +-24 h filter + filter/group-by of {len(names)} kind(s), then merge + threshold + "
              f"top-{TOP_K} over the parts; pyarrow restatement of model/count_co_events.py (polars not installable), "
              f"{cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {', '.join(names)}; {sessions:,} sessions / {n_aids:,} aids "
                               f"(BASELINE configs[{cfg_idx}]); timed on a bounded sample",
                   "sample_sessions": n_parts * part_sessions, "sample_parts": n_parts,
                   "same_work_as_gpu_arm": "same kinds, thresholds, merge and top-20; sample of the sessions, not all of them",
                   "c_oracle_one_core_pairs_per_s": c_pairs / c_dt,
                   "c_oracle_ideal_all_cores_pairs_per_s": c_pairs / c_dt * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# =====================================================================================================
# our arm
# =====================================================================================================
def cpu_baseline_port(host_cols, n_sessions_sample, names):
    """Plain-C brute-force oracle (1 core) on the first n_sessions_sample sessions of the workload.
    Returns (cpu_baseline dict, sample columns, {name: (thresholded table, top-20)})."""
    import numpy as np
    from oracle import c_oracle
    s = host_cols[0]
    cut = int(np.searchsorted(s, s[0] + n_sessions_sample, "left"))
    cols = [c[:cut] for c in host_cols]
    t0 = time.perf_counter()
    emitted_all = 0
    want = {}
    for name in names:
        oa, ob, oc, emitted, nded = c_oracle.count_name(*cols, name)
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=MIN_COUNT_TO_SAVE[name])
        want[name] = ((ka, kb, kc), c_oracle.top_n(ka, kb, kc, TOP_K), emitted)
        emitted_all += emitted
    dt = time.perf_counter() - t0
    cpu = {"value": emitted_all / dt, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"first {n_sessions_sample:,} sessions ({cut:,} events, {emitted_all:,} pairs) of the workload, "
                     f"{', '.join(names)}: oracle/cov_oracle.c + numpy threshold/top-{TOP_K}, {dt:.1f} s"}
    return cpu, cols, want


def cpu_baseline_weighted(eng, host_cols, n_sessions_sample, names):
    """EXTENSION: float64 oracle (oracle/cov_oracle.c::cov_oracle_score) on a sample + the GPU's fixed-point scores on it."""
    import numpy as np
    from oracle import c_oracle
    s = host_cols[0]
    cut = int(np.searchsorted(s, s[0] + n_sessions_sample, "left"))
    cols = [c[:cut] for c in host_cols]
    t0 = time.perf_counter()
    want, pairs = {}, 0
    for name in names:
        want[name] = c_oracle.score_name(*cols, name)
        pairs += int(want[name][3].sum())
    dt = time.perf_counter() - t0
    eng.load_events(*cols)
    worst = 0.0
    verdict = None
    for name, (oa, ob, osc, oc) in want.items():
        mc = MIN_COUNT_TO_SAVE[name]
        ga, gb, gs, gc = eng.count_weighted(name, min_count=mc).fetch()
        keep = oc >= mc
        if not (np.array_equal(ga, oa[keep]) and np.array_equal(gb, ob[keep]) and np.array_equal(gc.astype(np.uint32), oc[keep])):
            verdict = f"MISMATCH: {name} keys / counts"
            break
        if len(gs):
            worst = max(worst, float((np.abs(gs - osc[keep]) / osc[keep]).max()))
    if verdict is None:
        verdict = f"keys and counts identical, scores within {worst:.1e} relative of the float64 oracle (tolerance 1e-5)" if worst <= 1e-5 \
            else f"MISMATCH: score relative error {worst:.2e} > 1e-5"
    cpu = {"value": pairs / dt, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"first {n_sessions_sample:,} sessions ({cut:,} events, {pairs:,} pairs), {', '.join(names)}: float64 weighted "
                     f"oracle oracle/cov_oracle.c::cov_oracle_score, {dt:.1f} s"}
    return cpu, verdict


def parity_on_sample(eng, cols, want):
    """The GPU path on the CPU baseline's sample, compared with the oracle's tables and top-20 (outside any timed region)."""
    import numpy as np
    from otto_recommender_b200.retrieve import topn_long
    eng.load_events(*cols)
    for name, ((ka, kb, kc), (ta, tb, tc, _), emitted) in want.items():
        f = eng.count(name, min_count=MIN_COUNT_TO_SAVE[name])
        if eng.count_info()["n_pairs"] != emitted:
            return f"MISMATCH: {name} emitted pairs"
        ga, gb, gc = f.fetch()
        if not (np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc)):
            return f"MISMATCH: {name} thresholded table"
        la, lb, lc, _ = topn_long(*eng.topk(f, TOP_K))
        if not (np.array_equal(la, ta) and np.array_equal(lb, tb) and np.array_equal(lc, tc)):
            return f"MISMATCH: {name} top-{TOP_K}"
        f.free()
    return "identical"


def table_fingerprint(table, dev):
    """Order-independent fingerprint of one table shard: [rows, sum of counts, two wrapping 64-bit hash sums]."""
    import torch
    a, b, c = table.fetch(device=True)
    if a.numel() == 0:
        return torch.zeros(4, dtype=torch.int64, device=dev)
    k = (a.to(torch.int64) << 32) | b.to(torch.int64)
    c64 = c.to(torch.int64)
    h1 = (k * -7046029254386353131 + c64 * 0x632BE59BD9B4E019)          # wraps mod 2^64
    h1 = h1 ^ (h1 >> 29)
    h2 = ((k ^ (k >> 31)) * 0x2545F4914F6CDD1D + (c64 << 17)) ^ (k >> 7)
    return torch.stack([torch.tensor(a.numel(), dtype=torch.int64, device=dev), c64.sum(), h1.sum(), h2.sum()])


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from otto_recommender_b200 import Engine
    from otto_recommender_b200 import dist as covdist
    from otto_recommender_b200.synth import SynthSpec, generate

    metric, names, default_sessions, n_aids, extra, cfg_idx = WORKLOADS[args.workload]
    sessions = args.sessions or default_sessions
    aid_bits = (n_aids - 1).bit_length()          # catalogue size is known: no all-reduce for the key width
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: the same global synthetic dataset on every rank, then this rank's session range ----
    spec = SynthSpec(n_sessions=sessions, n_aids=n_aids, seed=42, **extra)
    d = generate(spec, dev)
    if world > 1:
        lens = torch.bincount(d["session"].long(), minlength=sessions).cpu().numpy()
        b = covdist.shard_bounds(lens, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        m = (d["session"] >= lo) & (d["session"] < hi)
        d = {k: v[m].contiguous() for k, v in d.items()}
        del m
    cols = [d["session"], d["aid"], d["ts"], d["type"]]
    del d
    n_rows = int(cols[0].numel())
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    eng = Engine(device=local_rank)
    weighted = args.workload == "decay"
    if weighted and world > 1:
        raise SystemExit("--workload decay (float time-decay extension) is single-GPU")
    exchange = {"nccl": covdist.count_exchange_first, "push": covdist.count_exchange_push,
                "scatter": covdist.count_exchange_scatter}[args.exchange]
    state = {"tables": []}

    def count_all(budget=None):
        """the co-event kinds of the workload on the loaded events -> (pairs, per-kind info); tables kept in state"""
        for t in state["tables"]:
            t.free()
        state["tables"] = []
        pairs, infos = 0, {}
        for name in names:
            mc = MIN_COUNT_TO_SAVE[name]
            if weighted:
                f = eng.count_weighted(name, min_count=mc)
                ci = eng.count_info()
                pairs += ci["n_pairs"]; infos[name] = ci
                state["tables"].append(f)
                continue
            if world > 1:                                     # raw keys cross NVLink once, reduced where they land
                f = exchange(eng, name, mc, aid_bits=aid_bits)
            else:                                             # threshold fused into the reduce
                f = eng.count(name, min_count=mc, pair_budget=budget)
            ci = eng.count_info()
            pairs += ci["n_pairs"]
            infos[name] = ci
            state["tables"].append(f)
        return pairs, infos

    def step_device():
        eng.load_events(*cols)
        pairs, infos = count_all(args.pair_budget)
        for f in state["tables"]:
            if weighted:
                f.topk_rows(TOP_K)
            else:
                eng.topk(f, TOP_K, fetch=False)           # the per-aid top-20 stays in HBM (the e2e arm copies it out)
        return pairs, infos

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = NvmlSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    for _ in range(args.warmup):
        pairs_local, infos = step_device()
    barrier()
    eng.kernel_stats(reset=True)
    eng.set_profiling(True, families=["sort_pass"])       # the dominant kernel, timed live in the timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        pairs_local, infos = step_device()
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    stats = eng.kernel_stats(reset=True)
    eng.set_profiling(True)                                # extra, untimed steps with every family timed; the first
    step_device()                                          # one creates the CUDA events (host stalls between the
    barrier()                                              # bracketing records leak into its figures): report the
    eng.kernel_stats(reset=True)                           # second
    step_device()
    barrier()
    stats_all = eng.kernel_stats(reset=True)
    eng.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    mem = eng.memory_info()

    # ---- fingerprint of the global thresholded tables (outside the timed region) --------------------------
    fp = (torch.stack([table_fingerprint(t, dev) for t in state["tables"]]).sum(0) if not weighted
          else torch.tensor([sum(t.rows for t in state["tables"]), 0, 0, 0], dtype=torch.int64, device=dev))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    p_all = torch.tensor([pairs_local, n_rows], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(p_all, op=dist.ReduceOp.SUM)
        dist.all_reduce(fp, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    pairs_global, rows_global = (int(x) for x in p_all.tolist())
    fp = [int(x) for x in fp.tolist()]
    ms_per_step = ms / args.steps
    value = pairs_global / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) columns in, results back on the host, through the public API ------------
    host_cols = [c.cpu().pin_memory() for c in cols]

    host_parts = Engine.split_at_sessions(*host_cols, 64) if world == 1 else None

    def step_e2e():
        if weighted:
            eng.load_events(*host_cols)                   # H2D inside
            count_all()
            d2h = 0
            for f in state["tables"]:
                ta, tb, ts_, tr = f.topk(TOP_K)           # D2H inside
                fa, fb, fs, fc = f.fetch()
                d2h += ta.size * 20 + fa.size * 20
            return d2h
        if world == 1 and not args.no_streamed_e2e:
            # the public ingest call: the population as parts in host memory (what reading the ETL's parquet parts
            # gives), copied on a second stream and counted group by group behind the copies
            for t in state["tables"]:
                t.free()
            state["tables"] = eng.count_parts(host_parts, names, [MIN_COUNT_TO_SAVE[n] for n in names])
            state["h2d"] = eng.count_info()["h2d_bytes"]      # what the library really copied (session column as runs)
        else:
            eng.load_events(*host_cols)                   # H2D inside
            count_all(args.pair_budget)
        d2h = 0
        for f in state["tables"]:
            ax, nv, ay, ac = eng.topk(f, TOP_K, pinned=True)   # D2H inside
            fa, fb, fc = f.fetch(order="count_desc", pinned=True)
            d2h += (ax.size + nv.size + ay.size + ac.size + 3 * fa.size) * 4
        return d2h

    for _ in range(max(1, min(args.warmup, 2))):
        d2h = step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d2h = step_e2e()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)             # fetches block the host: both clocks agree
    t_ms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t_ms.item()) / args.steps
    e2e_value = pairs_global / (e2e_ms_per_step * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (radix distribution pass) -------------------------------------
    peak, peak_src = _peaks()
    # achieved = algorithmic bytes (16 B per key: read once, written once) of every distribution pass launched in
    # the timed region / their CUDA-event time, measured live on the launching stream.
    sp = stats["sort_pass"]
    achieved = sp["algo_bytes"] / (sp["ms"] * 1e-3) / 1e9 if sp["ms"] > 0 else 0.0
    first = infos[names[0]]
    keys_sorted = first["n_pairs"] // 2 if (names[0] in ("click_to_click", "cart_to_cart", "buy_to_buy") and not weighted) else first["n_pairs"]
    per_launch_bytes = 16.0 * keys_sorted       # the big launches: bucket passes over this rank's keys of the first kind
    per_launch_ms = per_launch_bytes / (achieved * 1e9) * 1e3 if achieved > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "sort_pass_traffic.json")
    if os.path.exists(tp) and world == 1 and args.workload == "cooc" and sessions == FULL_SESSIONS:   # captured on this launch shape
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    kernels = {k: {"launches_per_step": v["launches"], "ms_per_step": v["ms"],
                   "algo_GBps": (v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
               for k, v in stats_all.items() if v["launches"]}
    launches = sum(v["launches"] for v in stats.values())
    # the family that takes the most time (from the extra, fully profiled step): since the whole-bucket reduce took a
    # pass away, that is no longer the distribution pass the roofline object tracks
    largest = None
    timed = {k: v for k, v in kernels.items() if v.get("ms_per_step")}
    if timed:
        lk = max(timed, key=lambda k: timed[k]["ms_per_step"])
        lv = timed[lk]
        largest = {"family": lk, "ms_per_step": lv["ms_per_step"], "algo_GBps": lv["algo_GBps"],
                   "frac_of_hbm_peak": (lv["algo_GBps"] / peak) if (peak and lv["algo_GBps"]) else None}

    # ---- CPU baseline on a bounded sample of the same workload + parity of the GPU path on that sample (N = 1) ----
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        hc = [c.numpy() for c in host_cols]
        sample_sessions = min(args.cpu_sample_sessions if len(names) == 1 else args.cpu_sample_sessions // 3, sessions)
        if weighted:
            cpu, parity = cpu_baseline_weighted(eng, hc, min(300_000, sessions), names)
        else:
            cpu, sample_cols, want = cpu_baseline_port(hc, sample_sessions, names)
            parity = parity_on_sample(eng, sample_cols, want)

    out = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {', '.join(f'{n} (min_count {MIN_COUNT_TO_SAVE[n]})' for n in names)}, top-{TOP_K}: "
                        f"{sessions:,} sessions / {rows_global:,} event rows / {n_aids:,} aids (BASELINE configs[{cfg_idx}])",
            "pairs_per_step": pairs_global,
            "fingerprint": {"table_rows": fp[0], "sum_of_counts": fp[1], "hash_sum_1": fp[2], "hash_sum_2": fp[3],
                            "of": "global thresholded (aid, aid_next, count) tables of all kinds, all ranks"},
            "sort_passes": first["sort_passes"], "fused_first_pass": bool(first.get("fused", 0)), "chunks": first["n_chunks"],
            "peak_bytes_rank0": mem["peak_bytes"],
            "parallelism": (f"session-sharded x{world}, keys re-sharded by hash(aid): " +
                            {"scatter": "the expansion stores every key straight into its owner's HBM over NVLink (fused first "
                                        "bucket pass + exchange, peer stores)",
                             "push": "round-1 path: keys written locally, then one partition pass with peer stores",
                             "nccl": "NCCL all-to-all"}[args.exchange])
            if world > 1 else "single GPU",
            "l2": "inputs (event columns, pair keys) are far larger than the 126 MB L2; no flush needed",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step,
                "h2d_bytes_per_step": int(state.get("h2d", 13 * n_rows)), "d2h_bytes_per_step": int(d2h),
                "input": ("host columns session i32 / aid i32 / ts i32 / type i8 in pinned memory (13 B per event row); "
                          + ("count_parts ships the session column run-length encoded" if "h2d" in state else "copied as they are"))},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "rs_onesweep_kernel (radix distribution pass)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "algo_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                     "share_of_step": sp["ms"] / max(ms, 1e-9), "largest_family": largest},
        "kernels": kernels,
        "clocks": clocks,
    }
    if parity is not None:
        out["config"]["parity_sample"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and "MISMATCH" in parity:
        raise SystemExit(f"parity check on the CPU sample failed: {parity}")


# =====================================================================================================
# --workload popularity: the popularity stage, same conventions, its own JSON line
# =====================================================================================================
def run_popularity(args):
    """events/s through ottocov_count_popularity on the synthetic OTTO shape with the event columns resident in
    HBM: general popularity (one cluster) and `--clusters` pseudo-clusters; the CPU restatement
    (oracle/popularity_oracle.py, pandas, one core) timed on a bounded sample of the same events."""
    import torch
    from otto_recommender_b200 import Engine
    from otto_recommender_b200.synth import SynthSpec, generate

    n_aids = 1_800_000
    sessions = args.sessions or FULL_SESSIONS
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    d = generate(SynthSpec(n_sessions=sessions, n_aids=n_aids, seed=42), dev)
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    n = int(s.numel())
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    cl_of_session = torch.randint(-1, args.clusters, (sessions,), generator=g, device=dev, dtype=torch.int32)
    cols = {1: torch.zeros(n, dtype=torch.int32, device=dev), args.clusters: cl_of_session[s.long()].contiguous()}
    ts_recent = int(t.max().item()) - 7 * 86400
    eng = Engine(0)
    runs = {}
    for ncl, cl in cols.items():
        for _ in range(max(args.warmup, 1)):
            r = eng.count_popularity(cl, a, t, y, ts_recent=ts_recent, keep_top_k=TOP_K)
        eng.kernel_stats(reset=True)
        eng.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            r = eng.count_popularity(cl, a, t, y, ts_recent=ts_recent, keep_top_k=TOP_K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        st = eng.kernel_stats(reset=True)
        eng.set_profiling(False)
        sp = st["sort_pass"]
        runs[f"cl{ncl}"] = {"ms_per_step": ms, "events_per_s": n / (ms * 1e-3), "rows_kept": int(len(r["aid"])),
                            "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in st.items() if v["launches"]},
                            "sort_pass_algo_GBps": (sp["algo_bytes"] / (sp["ms"] * 1e-3) / 1e9) if sp["ms"] > 0 else None}
    out = {"metric": "events/s (popularity counts + 6 ordinal ranks per cluster, top-20 kept)", "unit": "events/s",
           "value": runs[f"cl{args.clusters}"]["events_per_s"], "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": runs[f"cl{args.clusters}"]["ms_per_step"], "higher_is_better": True, "dtype": "u64",
           "data": "synthetic", "vs_baseline": None,
           "config": {"workload": f"count_popularity: {sessions:,} sessions / {n:,} events / {n_aids:,} aids, "
                                  f"{args.clusters} pseudo-clusters (and one cluster), keep_top_k {TOP_K}"},
           "runs": runs}
    if not args.no_cpu_baseline:
        from oracle import popularity_oracle as po
        m = min(n, args.cpu_sample_events)
        hc = [x[:m].cpu().numpy() for x in (cols[args.clusters], a, t, y)]
        t0 = time.perf_counter()
        po.popularity_ranks_frame(*hc, keep_top_k=TOP_K, ts_recent=ts_recent)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m / dt, "unit": "events/s", "cores": 1, "kind": "port",
                               "sample": f"first {m:,} events, {args.clusters} clusters, oracle/popularity_oracle.py "
                                         f"(pandas), {dt:.1f} s"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cooc", choices=list(WORKLOADS) + ["popularity"],
                    help="cooc = the BASELINE.json metric on configs[1] (default); all5 / longtail / scale4 = configs[2..4]; "
                         "decay = configs[3] with the float time-decay extension (N=1); "
                         "popularity = the popularity stage, N=1, its own JSON line")
    ap.add_argument("--sessions", type=int, default=0, help="sessions of the synthetic workload (0 = the workload's own size)")
    ap.add_argument("--pair-budget", type=int, default=None, help="co-event keys expanded at once (HBM footprint); default: from free HBM")
    ap.add_argument("--cpu-sample-sessions", type=int, default=1_000_000)
    ap.add_argument("--ref-parts", type=int, default=4, help="--impl reference: fixed parts per step")
    ap.add_argument("--ref-part-sessions", type=int, default=50_000, help="--impl reference: sessions per part")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="scatter", choices=["scatter", "push", "nccl"],
                    help="N > 1: scatter = the expansion stores keys straight into their owners' HBM over NVLink (default); "
                         "push = round-1 path (local keys + one partition pass with peer stores); nccl = NCCL all-to-all")
    ap.add_argument("--no-clock-sampler", action="store_true", help="do not sample clocks during the timed region")
    ap.add_argument("--no-streamed-e2e", action="store_true",
                    help="N = 1 e2e through load_events + count instead of the streamed count_parts")
    ap.add_argument("--clusters", type=int, default=50, help="--workload popularity: pseudo-clusters")
    ap.add_argument("--cpu-sample-events", type=int, default=10_000_000, help="--workload popularity: events of the CPU sample")
    args = ap.parse_args()
    if args.workload == "popularity":
        run_popularity(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
