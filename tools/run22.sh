mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py scale4 2>&1 | tee gpurun_out/bench_scale4.log | cut -c1-420
