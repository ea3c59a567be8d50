"""Print the headline fields of a bench.py JSON line: python tools/show_bench.py <log>"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(f"N={d['n_gpus']} ms/step {d['ms_per_step']:.2f}  {d['value'] / 1e9:.2f} G pairs/s   e2e {d['e2e']['ms_per_step']:.2f} ms "
      f"({d['e2e']['value'] / 1e9:.2f} G/s)  launches {d['gpu_launches']}")
r = d["roofline"]
print(f"roofline {r['kernel']}: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of {r['peak']:.0f}; share of step {r['share_of_step']:.2f}")
for k, v in d["kernels"].items():
    g = v["algo_GBps"]
    print(f"  {k:10s} x{v['launches_per_step']:<3d} {v['ms_per_step']:8.3f} ms  {g if g is None else round(g)} GB/s")
print("clocks", d.get("clocks"))
