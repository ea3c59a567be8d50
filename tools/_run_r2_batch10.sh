mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2b10_pytest.log 2>&1; tail -12 gpurun_out/r2b10_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b10_bench.log 2> gpurun_out/r2b10_bench.err; tail -2 gpurun_out/r2b10_bench.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2b10_bench.log | head -3
