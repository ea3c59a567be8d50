# usage: bash tools/_run_r2_dist.sh N   -- multi-GPU parity (vs the C oracle at 50k sessions, vs single GPU at 300k) + bench
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dist_check.py 50000 > gpurun_out/r2_dist_check_${N}gpu_oracle.log 2>&1; grep -c '"identical": true' gpurun_out/r2_dist_check_${N}gpu_oracle.log; tail -2 gpurun_out/r2_dist_check_${N}gpu_oracle.log
timeout 300 $TR tools/dist_check.py 300000 > gpurun_out/r2_dist_check_${N}gpu.log 2>&1; grep -c '"identical": true' gpurun_out/r2_dist_check_${N}gpu.log; tail -2 gpurun_out/r2_dist_check_${N}gpu.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.log 2> gpurun_out/r2_bench_n${N}.err; tail -3 gpurun_out/r2_bench_n${N}.err; python tools/show_bench.py gpurun_out/r2_bench_n${N}.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange push > gpurun_out/r2_bench_n${N}_oldpush.log 2>&1; python tools/show_bench.py gpurun_out/r2_bench_n${N}_oldpush.log
