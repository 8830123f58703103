#!/bin/bash
# GPU call D: the cta_group::2 pair kernel — op tests under a tight timeout first, then the UNet-level suite and the bench
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout=120 -x > gpurun_out/pytest_tc.log 2>&1
rc=$?; echo "== conv_tc suite rc=$rc"; tail -25 gpurun_out/pytest_tc.log
if [ $rc -ne 0 ]; then
  nvidia-smi --query-gpu=name,memory.used --format=csv
  timeout -k 10 240 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu -x --timeout=200 -k "32-128-32-32-128-3" > gpurun_out/sanitizer.log 2>&1
  grep -E "Invalid|Error|at 0x|by thread|kernel|=========" gpurun_out/sanitizer.log | head -30
  exit 0
fi
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 --ignore=tests/test_conv_tc_gpu.py > gpurun_out/pytest_main.log 2>&1
echo "== main suite rc=$?"; tail -15 gpurun_out/pytest_main.log
rm -f gpurun_out/layers.jsonl
DPB200_LAYERS_OUT=gpurun_out/layers.jsonl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'])
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_c1.err
