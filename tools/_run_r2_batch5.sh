mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --durations=8 > gpurun_out/r2b5_pytest.log 2>&1; tail -22 gpurun_out/r2b5_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b5_bench.log 2> gpurun_out/r2b5_bench.err; tail -3 gpurun_out/r2b5_bench.err; python tools/show_bench.py gpurun_out/r2b5_bench.log
timeout 400 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-streamed-e2e > gpurun_out/r2b5_bench_nostream.log 2>&1; python tools/show_bench.py gpurun_out/r2b5_bench_nostream.log | head -1
