# round 2, batch 1: fused first pass + SPLIT ranking chains -- tests, then A/B benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2b1_pytest.log 2>&1; tail -5 gpurun_out/r2b1_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2b1_bench_fused_s2.log 2> gpurun_out/r2b1_bench_fused_s2.err; python tools/show_bench.py gpurun_out/r2b1_bench_fused_s2.log
OTTOCOV_NO_FUSED_PASS=1 timeout 300 $B > gpurun_out/r2b1_bench_unfused_s2.log 2>&1; python tools/show_bench.py gpurun_out/r2b1_bench_unfused_s2.log
for v in s1 s4; do
  OTTOCOV_SO_NAME=libottocov_$v.so timeout 300 $B > gpurun_out/r2b1_bench_fused_$v.log 2>&1; python tools/show_bench.py gpurun_out/r2b1_bench_fused_$v.log
done
for v in "" _s1 _s4; do
  OTTOCOV_SO_NAME=libottocov$v.so timeout 120 python tools/bench_sort.py 3e8 2>&1 | tail -1
done
