#!/usr/bin/env python
"""bench.py -- co-event pairs/s of the B200 co-visitation counting path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sessions S]

Workload (BASELINE.json configs[1]): click-to-click 12 h co-visitation top-20 on the full synthetic
OTTO shape -- 12.9 M sessions, ~220 M events, 1.8 M aids (generator: otto_recommender_b200/synth.py,
SURVEY.md App. C).  One "step" = one pass of the hot path over that batch:

    raw event columns in HBM -> loader (order check / sort, dedup, split by type) -> window ranges ->
    pair expansion (keys written through a bijective mix, bucket histograms fused in) -> 3 radix
    distribution passes on the top hash bits -> bucketed hash reduce in shared memory with the threshold
    (count >= 10) and the symmetric mirror fused in -> key sort of the survivors -> segmented top-20

`value`  = emitted co-event pairs / device time, inputs resident in HBM (CUDA events, max over ranks).
`e2e`    = same metric through the public Python API with HOST (pinned) event columns in and the
           top-20 + thresholded pair table copied back to host inside the timed region.
N > 1    = strong scaling: the same 220 M events, sessions range-sharded over the ranks, counts
           re-sharded by hash(aid) with an NCCL all-to-all (otto_recommender_b200/dist.py).
--workload popularity = the popularity stage instead, one GPU, its own JSON line.
--impl reference = the reference's CPU algorithm (pyarrow restatement of model/count_co_events.py,
           oracle/ref_restatement.py -- polars itself is not installable here) on all host cores, one
           100k-session part of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "co-event pairs/sec (click-to-click 12h, top-20)"
UNIT = "pairs/s"
NAME = "click_to_click"
FULL_SESSIONS = 12_900_000
N_AIDS = 1_800_000
MIN_COUNT = 10
AID_BITS = (N_AIDS - 1).bit_length()      # catalogue size is known: no all-reduce for the key width
TOP_K = 20


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu = gpu_index
        self.rows = []          # (host time the line arrived, line)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if self.t0 is None or (self.t0 - 0.05 <= t <= (self.t1 or t) + 0.25)]
        if not rows:
            rows = [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


class NvmlSampler:
    """The same clocks + throttle reasons read in-process through NVML (what nvidia-smi itself calls), every
    100 ms while the timed region runs.  A separate `nvidia-smi -lms` process queries a dozen fields per sample
    and was measured to stall kernel launches and stream synchronisations on some hosts (a 32 ms step became
    43-54 ms while it ran); three light NVML calls per sample do not."""
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


# =====================================================================================================
# reference arm: the reference's CPU algorithm on the host cores
# =====================================================================================================
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import pyarrow as pa
    from oracle import ref_restatement as rr
    from otto_recommender_b200.synth import SynthSpec, generate_numpy

    part_sessions = 100_000
    d = generate_numpy(SynthSpec(n_sessions=part_sessions, n_aids=N_AIDS, seed=42))
    cols = (d["session"], d["aid"], d["ts"], d["type"])

    def step():
        # phase 1 for one part, click_to_click only would under-count the reference's work (it builds
        # the +-24 h merged frame once and scans it five times); time the full per-part body and
        # report click-to-click pairs / that time, as BASELINE.md section 2 does.
        res = rr.count_part(rr.events_table(*cols))
        t = rr.merge_counts(NAME, [res[NAME]], exact=True, min_count_to_save=MIN_COUNT)
        rr.top_n_per_aid(t, TOP_K)
        return int(pa.compute.sum(res[NAME]["count"]).as_py())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(args.steps):
        pairs += step()
    dt = time.perf_counter() - t0
    value = pairs / dt
    cores = pa.cpu_count()
    sample = (f"one 100k-session part ({len(cols[0]):,} events) of the 12.9M-session workload per step, all five "
              f"count types + click_to_click merge/top-{TOP_K}; pyarrow restatement of model/count_co_events.py "
              f"(polars not installable), {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "click_to_click 12h top-20, 12.9M sessions / 220M events / 1.8M aids "
                               "(configs[1]); timed on a bounded sample", "sample_sessions": part_sessions},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# =====================================================================================================
# our arm
# =====================================================================================================
def cpu_baseline_port(host_cols, n_sessions_sample):
    """Plain-C brute-force oracle (1 core) on the first n_sessions_sample sessions of the workload."""
    import numpy as np
    from oracle import c_oracle
    s = host_cols[0]
    cut = int(np.searchsorted(s, s[0] + n_sessions_sample, "left"))
    cols = [c[:cut] for c in host_cols]
    t0 = time.perf_counter()
    oa, ob, oc, emitted, nded = c_oracle.count_name(*cols, NAME)
    ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=MIN_COUNT)
    c_oracle.top_n(ka, kb, kc, TOP_K)
    dt = time.perf_counter() - t0
    return {"value": emitted / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_sessions_sample:,} sessions ({cut:,} events, {emitted:,} pairs) of the workload, "
                      f"oracle/cov_oracle.c + numpy threshold/top-{TOP_K}, {dt:.1f} s"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from otto_recommender_b200 import Engine
    from otto_recommender_b200.dist import count_exchange_first, count_exchange_push, shard_bounds
    from otto_recommender_b200.synth import SynthSpec, generate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: the same global synthetic dataset on every rank, then this rank's session range ----
    spec = SynthSpec(n_sessions=args.sessions, n_aids=N_AIDS, seed=42)
    d = generate(spec, dev)
    if world > 1:
        lens = torch.bincount(d["session"].long(), minlength=args.sessions).cpu().numpy()
        b = shard_bounds(lens, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        m = (d["session"] >= lo) & (d["session"] < hi)
        d = {k: v[m].contiguous() for k, v in d.items()}
        del m
    cols = [d["session"], d["aid"], d["ts"], d["type"]]
    n_rows = int(cols[0].numel())
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    eng = Engine(device=local_rank)
    exchange = count_exchange_first if args.exchange == "nccl" else count_exchange_push

    def step_device():
        eng.load_events(*cols)
        if world > 1:                                     # raw keys cross NVLink once, reduce where they land
            f = exchange(eng, NAME, MIN_COUNT, aid_bits=AID_BITS)
            ci = eng.count_info()
        else:                                             # threshold fused into the run-length reduce
            f = eng.count(NAME, min_count=MIN_COUNT)
            ci = eng.count_info()
        eng.topk(f, TOP_K, device=True)
        rows_f = f.rows
        f.free()
        return ci, rows_f

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started BEFORE the warm-up (its NVML start-up stalls the driver for tens of ms) and
    # keeps sampling every 100 ms; only samples that arrive inside the timed region are reported.
    sampler = ClockSampler(local_rank) if args.clock_sampler == "smi" else NvmlSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
        if isinstance(sampler, NvmlSampler) and not sampler.ok:        # no pynvml: fall back to nvidia-smi
            sampler = ClockSampler(local_rank)
            sampler.start()
    for _ in range(args.warmup):
        ci, rows_f = step_device()
    barrier()
    if args.breakdown and rank == 0:
        def timed(label, fn):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
            print(f"[breakdown] {label:14s} {(time.perf_counter() - t0) * 1e3:9.3f} ms", file=sys.stderr)
            return r
        for _ in range(2):
            timed("load_events", lambda: eng.load_events(*cols))
            fl = timed("count+filter", lambda: eng.count(NAME, min_count=MIN_COUNT))
            timed("topk", lambda: eng.topk(fl, TOP_K, device=True))
            timed("free", lambda: fl.free())
            print(f"[breakdown] memory {eng.memory_info()}", file=sys.stderr)
    eng.kernel_stats(reset=True)
    eng.set_profiling(True, families=["sort_pass"])       # the dominant kernel, timed live in the timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        ci, rows_f = step_device()
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

    pairs_local = ci["n_pairs"]
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    p_all = torch.tensor([pairs_local, n_rows, ci["n_unique"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(p_all, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    pairs_global, rows_global, uniq_sum = (int(x) for x in p_all.tolist())
    ms_per_step = ms / args.steps
    value = pairs_global / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) columns in, results back on the host, through the public API ------------
    host_cols = [c.cpu().pin_memory() for c in cols]

    def step_e2e():
        eng.load_events(*host_cols)                       # H2D inside
        if world > 1:
            f = exchange(eng, NAME, MIN_COUNT, aid_bits=AID_BITS)
        else:
            f = eng.count(NAME, min_count=MIN_COUNT)
        ax, nv, ay, ac = eng.topk(f, TOP_K, pinned=True)   # D2H inside
        fa, fb, fc = f.fetch(order="count_desc", pinned=True)
        d2h = (ax.size + nv.size + ay.size + ac.size + 3 * fa.size) * 4
        f.free()
        return d2h

    for _ in range(max(1, args.warmup)):
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
    # achieved = algorithmic bytes of every distribution pass in the timed region / their CUDA-event time.
    # The per-launch figures are those of the dominant launches (the bucket passes over this rank's keys:
    # 16 B per key per pass), the same launch shape `traffic` was captured on with ncu.
    sp = stats["sort_pass"]
    achieved = sp["algo_bytes"] / (sp["ms"] * 1e-3) / 1e9 if sp["ms"] > 0 else 0.0
    keys_sorted = pairs_local // 2          # click_to_click is symmetric: canonical half pairs
    per_launch_bytes = 16.0 * keys_sorted
    per_launch_ms = per_launch_bytes / (achieved * 1e9) * 1e3 if achieved > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "sort_pass_traffic.json")
    if os.path.exists(tp) and world == 1 and args.sessions == FULL_SESSIONS:     # captured on exactly this launch shape
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    kernels = {k: {"launches_per_step": v["launches"], "ms_per_step": v["ms"],
                   "algo_GBps": (v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
               for k, v in stats_all.items() if v["launches"]}
    launches = sum(v["launches"] for v in stats.values())

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only) ---------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        hc = [c.numpy() for c in host_cols]
        cpu = cpu_baseline_port(hc, min(args.cpu_sample_sessions, args.sessions))

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": f"click_to_click 12h co-visitation top-{TOP_K}, min_count {MIN_COUNT}: {args.sessions:,} sessions / "
                        f"{rows_global:,} event rows / {N_AIDS:,} aids (BASELINE configs[1])",
            "pairs_per_step": pairs_global, "table_rows_sum_over_ranks": uniq_sum,
            "thresholded_rows_rank0": rows_f, "sort_passes": ci["sort_passes"], "chunks": ci["n_chunks"],
            "parallelism": (f"session-sharded x{world}, keys re-sharded by hash(aid): " +
                            ("fused partition + peer stores over NVLink" if args.exchange == "push" else "NCCL all-to-all"))
            if world > 1 else "single GPU",
            "l2": "inputs (event columns, pair keys) are far larger than the 126 MB L2; no flush needed",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step,
                "h2d_bytes_per_step": 13 * n_rows, "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "rs_onesweep_kernel (radix distribution pass)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "algo_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                     "share_of_step": sp["ms"] / max(ms, 1e-9)},
        "kernels": kernels,
        "clocks": clocks,
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================
# --workload popularity: the popularity stage, same conventions, its own JSON line
# =====================================================================================================
def run_popularity(args):
    """events/s through ottocov_count_popularity on the synthetic OTTO shape with the event columns resident in
    HBM: general popularity (one cluster) and `--clusters` pseudo-clusters; the CPU restatement
    (oracle/popularity_oracle.py, pandas, one core) timed on a bounded sample of the same events."""
    import numpy as np
    import torch
    from otto_recommender_b200 import Engine
    from otto_recommender_b200.synth import SynthSpec, generate

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    d = generate(SynthSpec(n_sessions=args.sessions, n_aids=N_AIDS, seed=42), dev)
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    n = int(s.numel())
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    cl_of_session = torch.randint(-1, args.clusters, (args.sessions,), generator=g, device=dev, dtype=torch.int32)
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
           "config": {"workload": f"count_popularity: {args.sessions:,} sessions / {n:,} events / {N_AIDS:,} aids, "
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
    ap.add_argument("--sessions", type=int, default=FULL_SESSIONS, help="sessions of the synthetic workload")
    ap.add_argument("--cpu-sample-sessions", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="per-API-call wall times on stderr")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="N > 1: fused partition + peer-store kernel over NVLink (push) or NCCL all-to-all (nccl)")
    ap.add_argument("--no-clock-sampler", action="store_true", help="do not sample clocks during the timed region")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi"],
                    help="clocks + throttle reasons during the timed region: in-process NVML (default) or an nvidia-smi -lms process")
    ap.add_argument("--workload", default="cooc", choices=["cooc", "popularity"],
                    help="cooc = the BASELINE.json metric (default); popularity = the popularity stage, N=1, its own JSON line")
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
