# round 2: whole-bucket reduce, tightened insert loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fused_first_pass_overflow" > gpurun_out/r2b22_pytest.log 2>&1; tail -3 gpurun_out/r2b22_pytest.log
for C in 1; do
OTTOCOV_HRB_CHAINS=$C timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streamed-e2e > gpurun_out/r2b22_bench_c$C.log 2> gpurun_out/r2b22_bench_c$C.err; tail -2 gpurun_out/r2b22_bench_c$C.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b22_bench_c$C.log").read().strip().splitlines()[-1])
f=d["config"]["fingerprint"]
print("chains $C: step", round(d["ms_per_step"],2), "reduce", round(d["kernels"]["reduce"]["ms_per_step"],2), "pass", round(d["kernels"]["sort_pass"]["ms_per_step"],2), "fp", f["table_rows"], f["sum_of_counts"], f["hash_sum_1"], "passes", d["config"]["sort_passes"])
PY
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler --no-streamed-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hash_reduce_buckets_kernel' -c 1 -o gpurun_out/r02_hrb4 $CMD > gpurun_out/r2b22_ncu.log 2>&1
tail -2 gpurun_out/r2b22_ncu.log
ncu -i gpurun_out/r02_hrb4.ncu-rep --page details > gpurun_out/r02_hrb4_ncu_details.txt 2>&1
