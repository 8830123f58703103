#!/bin/bash
# GPU call U: A/B of GroupNorm builds through DPB200_LIB (default | lab_gnA = 4 resident blocks per SM | lab_gnB = A + explicit load
# groups in the backward kernels): stand-alone GroupNorm timings and the bench line with all legs for each; full suite on lab_gnB
set -u
mkdir -p gpurun_out
for v in default gnA gnB; do
  if [ $v != default ]; then export DPB200_LIB=$PWD/diff-pruning_b200/lab_$v.so; fi
  echo "######## library: $v"
  timeout 300 python scripts/gpu_prof_gn.py 2>&1 | tail -5
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_$v.json 2> gpurun_out/bench_c1_$v.err
  echo "== bench c1 rc=$?"; tail -2 gpurun_out/bench_c1_$v.err; V=$v python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/bench_c1_%s.json' % os.environ['V']).read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['gpu_launches'])
print({k: v for k, v in d['roofline']['other_launches_ms'].items() if v > 0.3})
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'])
PY
done
timeout 1500 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_gnB.log 2>&1
echo "== full suite (lab_gnB) rc=$?"; grep -n "^E  .*Error\|^E   *assert\|^FAILED\|passed\|failed" gpurun_out/pytest_gnB.log | head -30
