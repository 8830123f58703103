#!/bin/bash
# GPU call T: final evidence of round 2 (second session): full GPU suite, smoke(), the default bench line exactly as the driver runs it,
# the reference arm, the c3 / c5 lines, the ncu launch list of one C1 pass (the `ncu --set full` GroupNorm capture of profiles/r02_ncu_groupnorm.md was the last step of its first run)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_main.log 2>&1
echo "== full suite rc=$?"; grep -n "^E  .*Error\|^E   *assert\|^FAILED\|passed\|failed" gpurun_out/pytest_main.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== default bench rc=$? wall: $(( $(date +%s) - S0 )) s"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
print('c1', d['value'], d['ms_per_step'], 'e2e', d['e2e'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic','breakdown_ms','frac_of_tier_ceiling')})
print('cpu', d.get('cpu_baseline')); print('eager', d.get('gpu_eager_baseline'))
for k in ('finetune','finetune_bf16','config3'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c1.json 2> gpurun_out/bench_ref_c1.err; echo "ref arm rc=$?"; cut -c1-300 gpurun_out/bench_ref_c1.json
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "== bench c5 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().split('\n')[-1])
print('c5', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d.get('sampling', {}).get('unet_forward_ms'))
PY
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "== bench c3 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().split('\n')[-1])
print('c3', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c1_pass.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; python tools/launch_summary.py gpurun_out/launches_c1_pass.csv > gpurun_out/launches_c1_pass.md; head -24 gpurun_out/launches_c1_pass.md
