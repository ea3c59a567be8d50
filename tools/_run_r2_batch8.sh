mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "count_parts" > gpurun_out/r2b8_pytest.log 2>&1; tail -3 gpurun_out/r2b8_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b8_bench.log 2> gpurun_out/r2b8_bench.err; tail -3 gpurun_out/r2b8_bench.err; python tools/show_bench.py gpurun_out/r2b8_bench.log > gpurun_out/r2b8_show.txt; head -12 gpurun_out/r2b8_show.txt
OTTOCOV_SO_NAME=libottocov_hr256.so timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b8_bench_hr256.log 2>&1; python tools/show_bench.py gpurun_out/r2b8_bench_hr256.log > gpurun_out/r2b8_show2.txt; grep -E "ms/step|reduce" gpurun_out/r2b8_show2.txt
OTTOCOV_SO_NAME=libottocov_hr256.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hash_reduce_vs_oracle or hash_reduce_many or fused_first_pass_vs" 2>&1 | tail -2
