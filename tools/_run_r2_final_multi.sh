# usage: bash tools/_run_r2_final_multi.sh "N1 N2 ..." "workload ..."   -- bench lines at several N on one multi-GPU box
mkdir -p gpurun_out
for N in $1; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"
  for W in $2; do
    echo "== N=$N workload $W"
    timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload $W > gpurun_out/r2f_bench_${W}_n${N}.log 2> gpurun_out/r2f_bench_${W}_n${N}.err; tail -2 gpurun_out/r2f_bench_${W}_n${N}.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2f_bench_${W}_n${N}.log
  done
done
