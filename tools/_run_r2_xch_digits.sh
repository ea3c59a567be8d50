N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
for D in 128 64 32 16; do
  echo "== digits $D"
  OTTOCOV_XCH_DIGITS=$D timeout 300 $TR bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_xch_d${D}_n${N}.log 2>&1; python tools/show_bench.py gpurun_out/r2_xch_d${D}_n${N}.log | grep -E "ms/step|expand|sort_pass|reduce"
done
