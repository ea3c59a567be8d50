mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 400000 > gpurun_out/dist_check$N.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check$N.log
grep -E "DIST_CHECK|exit|rror" gpurun_out/dist_check$N.log | tail -5; grep -c '"identical": true' gpurun_out/dist_check$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_n$N.log
tail -2 gpurun_out/bench_n$N.log | cut -c1-700
