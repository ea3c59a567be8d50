set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hash3.log 2> gpurun_out/bench_hash3.err; python tools/show_bench.py gpurun_out/bench_hash3.log
OTTOCOV_SO_NAME=libottocov_pre12.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pre12.log 2> gpurun_out/bench_pre12.err; python tools/show_bench.py gpurun_out/bench_pre12.log | grep -E "ms/step|rle"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/plain_ll.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_hash.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -3 gpurun_out/ncu_ll.log
