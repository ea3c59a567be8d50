mkdir -p gpurun_out
for a in 1 2; do for d in 0 16; do echo "algo=$a dbg=$d"; OTTOCOV_RS_ALGO=$a OTTOCOV_RS_DEBUG=$d timeout 300 python tools/bench_sort.py 268435456; done; done 2>&1 | tee gpurun_out/bench_sort_direct.log
