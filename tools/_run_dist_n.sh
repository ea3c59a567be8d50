# usage: bash tools/_run_dist_n.sh N   -- multi-GPU parity check, then the bench, on N GPUs of one box
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 300000 > gpurun_out/dist_check$N.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check$N.log
grep -E "DIST_CHECK|exit|rror" gpurun_out/dist_check$N.log | tail -6; grep -c '"identical": true' gpurun_out/dist_check$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench_n$N.log
