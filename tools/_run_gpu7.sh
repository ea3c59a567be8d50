for v in ipt8_5 ipt8_6 ipt12_4; do
OTTOCOV_SO_NAME=libottocov_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "== $v"; python tools/show_bench.py gpurun_out/bench_$v.log | grep -E "ms/step|sort_pass|roofline"; tail -2 gpurun_out/bench_$v.err
done
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "sort or hash_reduce_many or config1" > gpurun_out/pytest_sub.log 2>&1; tail -3 gpurun_out/pytest_sub.log
