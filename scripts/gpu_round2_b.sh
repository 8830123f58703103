#!/bin/bash
# GPU call B: the two fixed tests, the bf16 op tests, sampling (graphed DDIM), cleaned conv_tc.cu through the whole suite, bench with layer dump
set -u
mkdir -p gpurun_out
timeout -k 10 500 python -m pytest tests/test_conv_bf16_gpu.py -q -m gpu --timeout=300 -x > gpurun_out/pytest_bf16.log 2>&1
echo "== bf16 suite rc=$?"; tail -25 gpurun_out/pytest_bf16.log
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 --ignore=tests/test_conv_bf16_gpu.py > gpurun_out/pytest_main.log 2>&1
echo "== main suite rc=$?"; tail -15 gpurun_out/pytest_main.log
rm -f gpurun_out/layers.jsonl
DPB200_LAYERS_OUT=gpurun_out/layers.jsonl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_c1.err
