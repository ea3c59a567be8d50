mkdir -p gpurun_out
for v in "" _a1 _a2 _a3 _m5 _m6; do
  echo "== variant '$v'"
  OTTOCOV_SO_NAME=libottocov$v.so timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2b4$v.log 2>&1; python tools/show_bench.py gpurun_out/r2b4$v.log | grep -E "ms/step|expand|sort_pass"
done
