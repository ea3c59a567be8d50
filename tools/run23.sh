mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "push_keys or exchange_first or partition" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 300000 > gpurun_out/dist_check2.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check2.log
grep -E "DIST_CHECK|exit|rror|Error" gpurun_out/dist_check2.log | tail -8; grep -c '"identical": true' gpurun_out/dist_check2.log
for ex in nccl push; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --exchange $ex > gpurun_out/bench_n2_$ex.log 2>&1; echo "bench $ex exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n2_$ex.log") if l.startswith("{")][-1])
    print("$ex", "ms/step", d["ms_per_step"], "G pairs/s", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("parse fail", e)
PY
done
