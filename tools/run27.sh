mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b$i.log 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/b$i.log').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kernel sum', sum(v['ms_per_step'] for v in d['kernels'].values()))"; done
nproc; cat /proc/cpuinfo | grep "model name" | head -1; uptime
