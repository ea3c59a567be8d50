mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --breakdown > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
cat gpurun_out/bench_full.log; tail -30 gpurun_out/bench_full.err
