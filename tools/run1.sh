mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --sessions 1000000 --steps 2 --warmup 1 --cpu-sample-sessions 100000 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/bench_small.log
