timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_big.log 2> gpurun_out/bench_big.err; python tools/show_bench.py gpurun_out/bench_big.log; tail -3 gpurun_out/bench_big.err
OTTOCOV_HR_MODE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; python tools/show_bench.py gpurun_out/bench_small.log | grep -E "ms/step|sort_pass|reduce"
