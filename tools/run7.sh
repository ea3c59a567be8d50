mkdir -p gpurun_out
for d in 0 1 2 3 4 8 9 15; do echo "dbg=$d"; OTTOCOV_RS_DEBUG=$d timeout 300 python tools/bench_sort.py 268435456; done 2>&1 | tee gpurun_out/bench_sort_dbg.log
