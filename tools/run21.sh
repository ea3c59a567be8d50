mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -k "golden or random_vs_oracle or edge or fused or symmetric or exchange_first or long_runs or merge_filter" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|exit" gpurun_out/sanitizer_memcheck.log | tail -6
for c in all5 longtail; do timeout 600 python tools/bench_configs.py $c 2>&1 | tee -a gpurun_out/bench_configs.log | cut -c1-330; done
