timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k40.log 2> gpurun_out/bench_k40.err; echo "== k40 (default)"; python tools/show_bench.py gpurun_out/bench_k40.log; tail -2 gpurun_out/bench_k40.err
for v in k36 k44 k48; do
OTTOCOV_SO_NAME=libottocov_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "== $v"; python tools/show_bench.py gpurun_out/bench_$v.log | grep -E "ms/step|sort_pass"; tail -2 gpurun_out/bench_$v.err
done
