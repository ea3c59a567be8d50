# usage: bash tools/_run_r2_dist_bench.sh N [workload]   -- bench only on N GPUs
N=$1; W=${2:-cooc}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload $W > gpurun_out/r2_bench_${W}_n${N}.log 2> gpurun_out/r2_bench_${W}_n${N}.err; tail -3 gpurun_out/r2_bench_${W}_n${N}.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2_bench_${W}_n${N}.log
