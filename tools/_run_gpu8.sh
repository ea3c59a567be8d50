for v in t28_2_256 t32_3_256 t24_2_512 t32_1_512 t40_2_256; do
OTTOCOV_SO_NAME=libottocov_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "== $v"; python tools/show_bench.py gpurun_out/bench_$v.log | grep -E "ms/step|sort_pass"; tail -2 gpurun_out/bench_$v.err
done
