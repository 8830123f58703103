#!/bin/bash
# GPU call R: full suite + bench c1 (all legs) + c5 after a GroupNorm / attention change
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_main.log 2>&1
echo "== full suite rc=$?"; grep -n "^E  .*Error\|^E   *assert\|^FAILED\|passed\|failed" gpurun_out/pytest_main.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; tail -2 gpurun_out/bench_c1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['gpu_launches'])
print({k: v for k, v in d['roofline']['other_launches_ms'].items() if v > 0.05})
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'), {a: b for a, b in d[k].get('roofline',{}).get('other_launches_ms', {}).items() if b > 0.3})
PY
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "== bench c5 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().split('\n')[-1])
print('c5', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d.get('sampling', {}).get('unet_forward_ms'))
print({k: v for k, v in d['roofline']['other_launches_ms'].items() if v > 0.3})
PY
