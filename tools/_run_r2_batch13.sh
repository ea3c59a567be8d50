OTTOCOV_TRACE=1 timeout 300 python - <<'PY' 2>&1 | grep -E "trace|e2e" | tail -45
import sys, time, torch
sys.path.insert(0, '.')
from otto_recommender_b200 import Engine
from otto_recommender_b200.synth import SynthSpec, generate
d = generate(SynthSpec(n_sessions=12_900_000, n_aids=1_800_000, seed=42), 'cuda')
host = [d[k].cpu().pin_memory() for k in ('session','aid','ts','type')]
del d; torch.cuda.empty_cache()
eng = Engine(0)
parts = Engine.split_at_sessions(*host, 64)
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    tabs = eng.count_parts(parts, ['click_to_click'], [10])
    torch.cuda.synchronize(); print('e2e count_parts ms', (time.perf_counter()-t0)*1e3, eng.count_info(), file=sys.stderr)
    for t in tabs: t.free()
PY
