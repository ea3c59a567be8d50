mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 600 -k "count_parts or fused_population or cli_three" > gpurun_out/r2b7_pytest.log 2>&1; tail -15 gpurun_out/r2b7_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b7_bench.log 2> gpurun_out/r2b7_bench.err; tail -3 gpurun_out/r2b7_bench.err; python tools/show_bench.py gpurun_out/r2b7_bench.log | head -3
OTTOCOV_NO_SESSION_RLE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b7_bench_norle.log 2>&1; python tools/show_bench.py gpurun_out/r2b7_bench_norle.log | head -1
