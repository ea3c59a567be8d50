"""BASELINE.json configs 3-5 as one-shot measurements (not the bench.py contract line): all five co-event
kinds on the full synthetic shape, the long-tail click_to_cart_or_buy case, and the 4x-scale footprint run.
Prints one JSON line per measurement; results are checked by size-independent properties only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from otto_recommender_b200 import Engine
from otto_recommender_b200.config import DEFAULT_CONFIG as CFG
from otto_recommender_b200.synth import SynthSpec, generate

which = sys.argv[1] if len(sys.argv) > 1 else "all5"
dev = torch.device("cuda", 0)
eng = Engine(0)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


if which == "all5":
    d = generate(SynthSpec(n_sessions=12_900_000, seed=42), dev)
    cols = [d[k] for k in ("session", "aid", "ts", "type")]
    ms_load, _ = timed(lambda: eng.load_events(*cols))
    total_ms, total_pairs = ms_load, 0
    for name in CFG.CO_EVENTS_TO_COUNT:
        def run():
            t = eng.count(name, min_count=CFG.MIN_COUNT_TO_SAVE[name])
            ci = eng.count_info()
            eng.topk(t, 20, device=True)
            rows = t.rows
            t.free()
            return ci, rows
        ms, (ci, rows) = timed(run)
        total_ms += ms; total_pairs += ci["n_pairs"]
        print(json.dumps({"config": "all five kinds, full shape, 1xB200", "name": name, "ms": ms, "pairs": ci["n_pairs"],
                          "rows_after_threshold": rows, "G_pairs_per_s": ci["n_pairs"] / ms / 1e6}))
    print(json.dumps({"config": "all five kinds, full shape, 1xB200", "name": "TOTAL (load once + 5 kinds)", "ms": total_ms,
                      "pairs": total_pairs, "G_pairs_per_s": total_pairs / total_ms / 1e6, "memory": eng.memory_info()}))
elif which == "longtail":
    d = generate(SynthSpec(n_sessions=12_900_000, seed=42, force_long_click_session=465), dev)
    cols = [d[k] for k in ("session", "aid", "ts", "type")]
    eng.load_events(*cols)
    name = "click_to_cart_or_buy"
    def run():
        t = eng.count(name, min_count=CFG.MIN_COUNT_TO_SAVE[name]); ci = eng.count_info(); eng.topk(t, 20, device=True)
        r = t.rows; t.free(); return ci, r
    ms, (ci, rows) = timed(run)
    full = eng.count(name)
    c1 = eng.count(type_this=0, next_types=[1], window=86400); c2 = eng.count(type_this=0, next_types=[2], window=86400)
    m = eng.merge([c1, c2])
    same = all(torch.equal(x, y) for x, y in zip(full.fetch(device=True), m.fetch(device=True)))
    print(json.dumps({"config": "click_to_cart_or_buy 24h, long-tail (one 465-click session), 1xB200", "ms": ms,
                      "pairs": ci["n_pairs"], "rows_after_threshold": rows, "G_pairs_per_s": ci["n_pairs"] / ms / 1e6,
                      "cart_or_buy == cart + buy": bool(same), "sum_of_counts == pairs": full.total() == ci["n_pairs"]}))
elif which == "scale4":
    d = generate(SynthSpec(n_sessions=51_600_000, n_aids=7_200_000, seed=42), dev)
    cols = [d[k] for k in ("session", "aid", "ts", "type")]
    n_rows = cols[0].numel()
    torch.cuda.empty_cache()
    info = eng.load_events(*cols)
    for budget in (0, 2_000_000_000, 1_000_000_000):
        t0 = time.perf_counter()
        t = eng.count("click_to_click", min_count=10, pair_budget=budget)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        ci = eng.count_info()
        print(json.dumps({"config": "4x scale: 51.6M sessions / 7.2M aids, click_to_click, 1xB200", "event_rows": n_rows,
                          "aid_bits": info["aid_bits"], "pair_budget": budget, "chunks": ci["n_chunks"], "ms": ms,
                          "pairs": ci["n_pairs"], "rows_after_threshold": t.rows, "G_pairs_per_s": ci["n_pairs"] / ms / 1e6,
                          "memory": eng.memory_info()}))
        t.free(); eng.trim()
