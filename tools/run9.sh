mkdir -p gpurun_out
timeout 120 python tools/bench_sort.py 1000000 2>&1 | tail -3
timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 -k "sort" 2>&1 | tail -5
timeout 300 python tools/bench_sort.py 268435456 2>&1 | tee gpurun_out/bench_sort.log | tail -3
