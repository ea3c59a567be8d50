mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2b11_pytest.log 2>&1; tail -12 gpurun_out/r2b11_pytest.log
for W in cooc all5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/r2b11_bench_$W.log 2> gpurun_out/r2b11_bench_$W.err; tail -2 gpurun_out/r2b11_bench_$W.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2b11_bench_$W.log > gpurun_out/r2b11_show_$W.txt; head -2 gpurun_out/r2b11_show_$W.txt
done
