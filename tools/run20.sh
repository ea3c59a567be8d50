mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-clock-sampler > gpurun_out/b_nosamp.log 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/b_nosamp.log').read().strip().splitlines()[-1]); print('no sampler: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_samp.log 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/b_samp.log').read().strip().splitlines()[-1]); print('sampler   : ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
done
