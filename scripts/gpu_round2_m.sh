#!/bin/bash
# GPU call M: persistent fprop / dgrad kernel with the A operand through TENSOR memory (TS mode, single accumulator set): tests + bench
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_main.log 2>&1
echo "== full suite rc=$?"; grep -n "^E  .*Error\|^E   *assert\|^FAILED\|passed\|failed" gpurun_out/pytest_main.log | head -30
timeout 300 python scripts/time_conv_shapes.py > gpurun_out/conv_shapes.txt 2>&1; tail -11 gpurun_out/conv_shapes.txt | cut -c1-150
DPB200_LAYERS_OUT=gpurun_out/layers_c1.json timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['gpu_launches'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
for k,v in list(d['roofline']['top_layers_ms'].items())[:10]: print(k, v)
PY
tail -3 gpurun_out/bench_c1.err
