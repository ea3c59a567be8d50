mkdir -p gpurun_out
for v in 0 1 2; do OTTOCOV_RS_ALGO=$v timeout 300 python tools/bench_sort.py 268435456; done 2>&1 | tee gpurun_out/bench_sort.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "sort or golden or random" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_sort.py 67108864 > gpurun_out/plain_sort.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 8 -c 1 -o gpurun_out/sort_pass_v4 -f \
    python tools/bench_sort.py 67108864 > gpurun_out/ncu_v4.log 2>&1
echo "ncu exit $?"
