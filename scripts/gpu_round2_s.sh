#!/bin/bash
# GPU call S: parallel GroupNorm finalize (default library) and, as an A/B through DPB200_LIB, the SFU SiLU + 16-bit dropout build
# (diff-pruning_b200/lab_fastgn.so): full suite, GroupNorm stand-alone timings, bench c1 with all legs — for both libraries
set -u
mkdir -p gpurun_out
for v in default fastgn; do
  if [ $v = fastgn ]; then export DPB200_LIB=$PWD/diff-pruning_b200/lab_fastgn.so; fi
  echo "######## library: $v"
  timeout 1500 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_$v.log 2>&1
  echo "== full suite rc=$?"; grep -n "^E  .*Error\|^E   *assert\|^FAILED\|passed\|failed" gpurun_out/pytest_$v.log | head -30
  timeout 300 python scripts/gpu_prof_gn.py 2>&1 | tail -6
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_$v.json 2> gpurun_out/bench_c1_$v.err
  echo "== bench c1 rc=$?"; tail -2 gpurun_out/bench_c1_$v.err; V=$v python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/bench_c1_%s.json' % os.environ['V']).read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['gpu_launches'])
print({k: v for k, v in d['roofline']['other_launches_ms'].items() if v > 0.05})
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('breakdown_ms'), d[k].get('roofline',{}).get('frac'), {a: b for a, b in d[k].get('roofline',{}).get('other_launches_ms', {}).items() if b > 0.3})
PY
done
