mkdir -p gpurun_out
for v in 3 4; do OTTOCOV_RS_MINB=$v timeout 300 python tools/bench_sort.py 268435456; done 2>&1 | tee gpurun_out/bench_sort.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
