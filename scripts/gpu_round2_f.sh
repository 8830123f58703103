#!/bin/bash
# GPU call F: full suite after reverting the slab GroupNorm / pair kernel, new colsum, split wide GroupNorm; bench c1 (+c3 leg), c5
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_main.log 2>&1
echo "== full suite rc=$?"; tail -15 gpurun_out/pytest_main.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_c1.err
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "== bench c5 rc=$?"; tail -c 3500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
