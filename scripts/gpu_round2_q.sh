#!/bin/bash
# GPU call Q: GroupNorm launch diet (finalize folded into the apply kernels, dgamma / dbeta on the side stream) + fused q/k/v projection:
# full suite, bench c1 with all legs, then the same C1 pass with the q/k/v fusion off (A/B inside one box)
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
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
# (the A/B switch DPB200_FUSE_QKV existed only at the commit this call ran on)
DPB200_FUSE_QKV=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-finetune --no-c3 > gpurun_out/bench_c1_noqkv.json 2> gpurun_out/bench_c1_noqkv.err
echo "== bench c1 (q/k/v unfused) rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1_noqkv.json').read().strip().split('\n')[-1])
print('c1 unfused qkv', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['gpu_launches'])
PY
