mkdir -p gpurun_out
OTTOCOV_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err
grep trace gpurun_out/bench_trace.err | tail -22
