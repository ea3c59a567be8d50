# round 2: whole-bucket reduce with balanced warp queues; one vs two probe chains per lane
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fused_first_pass_overflow" > gpurun_out/r2b20_pytest.log 2>&1; tail -3 gpurun_out/r2b20_pytest.log
for C in 2 1; do
OTTOCOV_HRB_CHAINS=$C timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streamed-e2e > gpurun_out/r2b20_bench_c$C.log 2> gpurun_out/r2b20_bench_c$C.err; tail -2 gpurun_out/r2b20_bench_c$C.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b20_bench_c$C.log").read().strip().splitlines()[-1])
f=d["config"]["fingerprint"]
print("chains $C: step", round(d["ms_per_step"],2), "reduce", round(d["kernels"]["reduce"]["ms_per_step"],2), "pass", round(d["kernels"]["sort_pass"]["ms_per_step"],2), "fp", f["table_rows"], f["sum_of_counts"], f["hash_sum_1"], "passes", d["config"]["sort_passes"])
PY
done
