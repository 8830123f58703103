#!/bin/bash
# GPU call N: evidence refresh for the final round-2 build: smoke(), reference arm, c5 / c3 bench lines, ncu launch list + conv traffic +
# full-set captures of the tensor kernels
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c1.json 2> gpurun_out/bench_ref_c1.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/bench_ref_c1.json
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "== bench c5 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().split('\n')[-1])
print('c5', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d.get('sampling'), d.get('gpu_eager_baseline'))
PY
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "== bench c3 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().split('\n')[-1])
print('c3', d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d.get('gpu_eager_baseline'), d.get('cpu_baseline'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c1_pass.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; python tools/launch_summary.py gpurun_out/launches_c1_pass.csv > gpurun_out/launches_c1_pass.md; head -24 gpurun_out/launches_c1_pass.md
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -k regex:'conv_tc|wgrad_tc|splitk' --csv --log-file gpurun_out/conv_traffic.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"; python tools/conv_traffic.py gpurun_out/conv_traffic.csv gpurun_out/r02_conv_traffic.json \
  "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc|wgrad_tc|splitk, one eager C1 batch-128 Taylor pass (scripts/gpu_round2_n.sh), round-2 final build"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_ps" -s 8 -c 6 \
   -o gpurun_out/prof_ps -f python bench.py --profile-pass --batch 128 > gpurun_out/ncu_full_ps.log 2>&1
echo "full ps rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"wgrad_tc" -s 4 -c 4 \
   -o gpurun_out/prof_wg -f python bench.py --profile-pass --batch 128 > gpurun_out/ncu_full_wg.log 2>&1
echo "full wg rc=$?"; ls -la gpurun_out/*.ncu-rep
