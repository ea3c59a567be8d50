N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_n$N.log 2>&1; echo "bench exit $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.log") if l.startswith("{")][-1])
print("N=$N ms/step", d["ms_per_step"], "G pairs/s", d["value"]/1e9, "e2e ms", d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
