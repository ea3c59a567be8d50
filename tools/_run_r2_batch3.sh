mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --durations=15 > gpurun_out/r2b3_pytest.log 2>&1; tail -25 gpurun_out/r2b3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b3_bench.log 2> gpurun_out/r2b3_bench.err; tail -3 gpurun_out/r2b3_bench.err; python tools/show_bench.py gpurun_out/r2b3_bench.log
