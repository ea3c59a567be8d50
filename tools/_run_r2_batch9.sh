mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2b9_pytest.log 2>&1; tail -4 gpurun_out/r2b9_pytest.log
for W in cooc all5 longtail; do
  echo "== workload $W"
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $W > gpurun_out/r2b9_bench_$W.log 2> gpurun_out/r2b9_bench_$W.err; tail -2 gpurun_out/r2b9_bench_$W.err | cut -c1-400; python tools/show_bench.py gpurun_out/r2b9_bench_$W.log > gpurun_out/r2b9_show_$W.txt; head -11 gpurun_out/r2b9_show_$W.txt
done
