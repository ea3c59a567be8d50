mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_full.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "G pairs/s", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], "roof", round(d["roofline"]["achieved"]), d["roofline"]["frac"])
print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
