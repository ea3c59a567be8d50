# round 2, N=1 evidence run: GPU tests, smoke, default bench (driver-style steps), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_cooc_n1.log 2> gpurun_out/r2f_bench_cooc_n1.err; echo "bench exit $?"; tail -2 gpurun_out/r2f_bench_cooc_n1.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2f_bench_cooc_n1.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.log 2>&1; tail -c 1500 gpurun_out/r2f_bench_reference.log
