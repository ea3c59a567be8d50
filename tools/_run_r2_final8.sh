mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/dist_check.py 50000 > gpurun_out/r2f_dist_check_8gpu_oracle.log 2>&1; grep -c '"identical": true' gpurun_out/r2f_dist_check_8gpu_oracle.log; tail -1 gpurun_out/r2f_dist_check_8gpu_oracle.log
bash tools/_run_r2_final_multi.sh "8" "cooc all5 scale4"
