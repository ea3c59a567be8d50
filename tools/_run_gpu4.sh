timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_popularity.py > gpurun_out/bench_popularity.log 2> gpurun_out/bench_popularity.err; echo "pop exit $?"; tail -c 1800 gpurun_out/bench_popularity.log; tail -3 gpurun_out/bench_popularity.err
