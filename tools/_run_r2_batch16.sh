# round 2: whole-bucket hash reduce (one distribution pass fewer): parity children first, then the bench per round count
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fused or hash_reduce" > gpurun_out/r2b16_pytest.log 2>&1; tail -5 gpurun_out/r2b16_pytest.log
for RB in 1 2 0; do
OTTOCOV_HRB_ROUND_BITS=$RB timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streamed-e2e > gpurun_out/r2b16_bench_rb$RB.log 2> gpurun_out/r2b16_bench_rb$RB.err; tail -2 gpurun_out/r2b16_bench_rb$RB.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2b16_bench_rb$RB.log 2>/dev/null | head -10
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b16_bench_rb$RB.log").read().strip().splitlines()[-1])
print("fingerprint", d["config"]["fingerprint"]["table_rows"], d["config"]["fingerprint"]["sum_of_counts"], d["config"]["fingerprint"]["hash_sum_1"], d["config"]["fingerprint"]["hash_sum_2"], "passes", d["config"]["sort_passes"])
PY
done
