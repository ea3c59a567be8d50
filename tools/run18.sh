mkdir -p gpurun_out
for a in 1 2; do echo "algo=$a"; OTTOCOV_RS_ALGO=$a timeout 300 python tools/bench_sort.py 268435456; done 2>&1 | tee gpurun_out/bench_sort_ptx.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "sort or golden or random or chunk" 2>&1 | tail -3
