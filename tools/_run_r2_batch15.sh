# round 2: loader statistics fused into the split scan + implicit window starts for symmetric kinds
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2b15_pytest.log 2>&1; tail -4 gpurun_out/r2b15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b15_bench_cooc_n1.log 2> gpurun_out/r2b15_bench_cooc_n1.err; echo "bench exit $?"; tail -2 gpurun_out/r2b15_bench_cooc_n1.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2b15_bench_cooc_n1.log
OTTOCOV_NO_FUSED_LOADER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b15_bench_nofl.log 2> gpurun_out/r2b15_bench_nofl.err; python tools/show_bench.py gpurun_out/r2b15_bench_nofl.log | head -8
