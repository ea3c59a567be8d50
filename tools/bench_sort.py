"""Micro-benchmark of the radix distribution pass: N random pair-shaped keys (two 21-bit fields),
6 passes, CUDA-event time per pass from the library's own profiling counters."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from otto_recommender_b200 import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
eng = Engine(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
lo = torch.randint(0, 1_800_000, (n,), generator=g, device="cuda", dtype=torch.int64)
hi = torch.randint(0, 1_800_000, (n,), generator=g, device="cuda", dtype=torch.int64)
src = (hi << 32) | lo
del lo, hi
for rep in range(4):
    keys = src.clone()
    if rep == 1:
        eng.kernel_stats(reset=True); eng.set_profiling(True)
    eng.sort_u64(keys.data_ptr(), None, n, 0, 21)
    eng.sort_u64(keys.data_ptr(), None, n, 32, 53)
torch.cuda.synchronize()
st = eng.kernel_stats()
sp = st["sort_pass"]
ok = bool((keys[1:] >= keys[:-1]).all())
print(json.dumps({"n": n, "algo": os.environ.get("OTTOCOV_RS_ALGO", "1"), "sorted": ok,
                  "pass_ms": sp["ms"] / sp["launches"], "pass_GBps": sp["algo_bytes"] / sp["ms"] / 1e6,
                  "hist_GBps": st["histogram"]["algo_bytes"] / st["histogram"]["ms"] / 1e6}))
