timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hash4.log 2> gpurun_out/bench_hash4.err; python tools/show_bench.py gpurun_out/bench_hash4.log
for v in scan8_3 scan4_3 scan4_4; do
OTTOCOV_SO_NAME=libottocov_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo $v; python tools/show_bench.py gpurun_out/bench_$v.log | grep -E "ms/step|load|window|topk"
done
