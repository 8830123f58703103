#!/bin/bash
# GPU call H: fprop / dgrad / attention GEMMs on the 3 x fp16 split (kind::f16), wgrad still 3xTF32: op tests, UNet parity, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q --timeout=600 > gpurun_out/pytest_tc.log 2>&1
echo "== conv tc rc=$?"; tail -25 gpurun_out/pytest_tc.log
timeout 300 python scripts/time_conv_shapes.py > gpurun_out/conv_shapes.txt 2>&1; tail -12 gpurun_out/conv_shapes.txt
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -q --timeout=900 > gpurun_out/pytest_unet.log 2>&1
echo "== unet rc=$?"; tail -25 gpurun_out/pytest_unet.log
DPB200_LAYERS_OUT=gpurun_out/layers_c1.json timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_c1.err
