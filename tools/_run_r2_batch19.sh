# round 2: two-phase emit + CAS-first insert (both reduce kernels): parity, then the bench with the tile kernel and with whole buckets
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fused or hash_reduce or reduce_pairs or count_parts" > gpurun_out/r2b19_pytest.log 2>&1; tail -3 gpurun_out/r2b19_pytest.log
for H in 0 1; do
OTTOCOV_HRB=$H timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streamed-e2e > gpurun_out/r2b19_bench_h$H.log 2> gpurun_out/r2b19_bench_h$H.err; tail -2 gpurun_out/r2b19_bench_h$H.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b19_bench_h$H.log").read().strip().splitlines()[-1])
f=d["config"]["fingerprint"]
print("hrb $H: step", round(d["ms_per_step"],2), "reduce", round(d["kernels"]["reduce"]["ms_per_step"],2), "pass", round(d["kernels"]["sort_pass"]["ms_per_step"],2), "fp", f["table_rows"], f["sum_of_counts"], f["hash_sum_1"], "passes", d["config"]["sort_passes"])
PY
done
