mkdir -p gpurun_out
for v in 0 1 2; do OTTOCOV_RS_ALGO=$v timeout 300 python tools/bench_sort.py 268435456; done 2>&1 | tee gpurun_out/bench_sort.log
for v in 0 2; do OTTOCOV_RS_ALGO=$v timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "sort" 2>&1 | tail -2; done
