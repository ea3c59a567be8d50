export OTTOCOV_HR_MODE=1
for v in "--clock-sampler nvml" "--clock-sampler smi" "--no-clock-sampler"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > gpurun_out/bench_s.log 2> gpurun_out/bench_s.err; echo "== $v"; python tools/show_bench.py gpurun_out/bench_s.log | grep -E "ms/step|clocks"; tail -2 gpurun_out/bench_s.err
done
