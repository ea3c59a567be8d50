# round 2, last build: 8-GPU parity vs the C oracle, then the bench at N = 8, 4, 2 on the same box
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/dist_check.py 50000 > gpurun_out/r2g_dist_check_8gpu_oracle.log 2>&1; grep -c '"identical": true' gpurun_out/r2g_dist_check_8gpu_oracle.log; tail -1 gpurun_out/r2g_dist_check_8gpu_oracle.log
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"
  timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_cooc_n${N}.log 2> gpurun_out/r2g_bench_cooc_n${N}.err; tail -2 gpurun_out/r2g_bench_cooc_n${N}.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2g_bench_cooc_n${N}.log | head -12
done
