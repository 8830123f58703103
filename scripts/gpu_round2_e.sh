#!/bin/bash
# GPU call E: LDM test, GroupNorm op tests, ncu launch lists (C1 pass + bf16 finetune step) with the slab GroupNorm kernels, bench c1 + c5
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -q -m gpu --timeout=400 -k "ldm or groupnorm" > gpurun_out/pytest_sel.log 2>&1
echo "== ldm + groupnorm rc=$?"; tail -25 gpurun_out/pytest_sel.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c1_pass.csv python bench.py --profile-pass > gpurun_out/ncu_launch.log 2>&1
echo "launch list c1 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_ft_bf16.csv python scripts/gpu_prof_finetune.py c1 bf16 > gpurun_out/ncu_launch_bf16.log 2>&1
echo "launch list bf16 rc=$?"
for f in launches_c1_pass launches_ft_bf16; do echo "## $f"; python tools/launch_summary.py gpurun_out/$f.csv | head -16; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
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
