for v in libottocov_old.so libottocov.so libottocov_old.so libottocov.so; do
OTTOCOV_SO_NAME=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; echo "== $v"; python tools/show_bench.py gpurun_out/bench_x.log | grep -E "ms/step|load|window"
done
