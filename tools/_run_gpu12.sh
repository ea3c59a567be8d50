timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipe2.log 2> gpurun_out/bench_pipe2.err; python tools/show_bench.py gpurun_out/bench_pipe2.log | grep -E "ms/step|load|roofline"; tail -2 gpurun_out/bench_pipe2.err
