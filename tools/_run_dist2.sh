mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 300000 > gpurun_out/dist_check2.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check2.log
grep -E "DIST_CHECK|exit|rror" gpurun_out/dist_check2.log | tail -6; grep -c '"identical": true' gpurun_out/dist_check2.log
