mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 300000 > gpurun_out/dist_check2.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check2.log
grep -E "identical|DIST_CHECK|exit|rror" gpurun_out/dist_check2.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench2 exit $?" >> gpurun_out/bench_n2.log
tail -3 gpurun_out/bench_n2.log | cut -c1-1500
