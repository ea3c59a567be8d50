mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 600 -k "count_parts or fused_population or cli_three" > gpurun_out/r2b14_pytest.log 2>&1; tail -3 gpurun_out/r2b14_pytest.log
for W in cooc all5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/r2b14_bench_$W.log 2> gpurun_out/r2b14_bench_$W.err; tail -2 gpurun_out/r2b14_bench_$W.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2b14_bench_$W.log > gpurun_out/r2b14_show_$W.txt; head -1 gpurun_out/r2b14_show_$W.txt
done
bash tools/_run_r2_batch13.sh 2>&1 | tail -16
