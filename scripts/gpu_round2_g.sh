#!/bin/bash
# GPU call G: split-K for the small-M fprop / dgrad launches: op tests, UNet parity tests, bench c1 (+c3 leg), c5, layer tables
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x --timeout=600 > gpurun_out/pytest_tc.log 2>&1
echo "== conv tc rc=$?"; tail -12 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py tests/test_conv_bf16_gpu.py -q --timeout=900 > gpurun_out/pytest_unet.log 2>&1
echo "== unet rc=$?"; tail -12 gpurun_out/pytest_unet.log
DPB200_LAYERS_OUT=gpurun_out/layers_c1.json timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_c1.err
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "== bench c5 rc=$?"; tail -c 2500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
