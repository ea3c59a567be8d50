mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in 3 4; do OTTOCOV_RLE_MINB=$v timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_rle$v.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rle$v.log").read().strip().splitlines()[-1])
print("RLE_MINB=$v ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
done
