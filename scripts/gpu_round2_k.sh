#!/bin/bash
# GPU call K (profiling): launch list of one eager C1 pass, DRAM traffic of its conv launches, full-set captures of the two tensor kernels
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c1_pass.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; python tools/launch_summary.py gpurun_out/launches_c1_pass.csv > gpurun_out/launches_c1_pass.md; head -30 gpurun_out/launches_c1_pass.md
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -k regex:'conv_tc|wgrad_tc|splitk' --csv --log-file gpurun_out/conv_traffic.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"; python tools/conv_traffic.py gpurun_out/conv_traffic.csv gpurun_out/r02_conv_traffic.json \
  "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc|wgrad_tc|splitk, one eager C1 batch-128 Taylor pass (scripts/gpu_round2_k.sh), round-2 3 x fp16 split build"
# full-set capture: skip the first launches (small layers), take a few of each kernel from the 32x32 / 16x16 levels
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_ps" -s 8 -c 4 \
   -o gpurun_out/prof_ps -f python bench.py --profile-pass --batch 128 > gpurun_out/ncu_full_ps.log 2>&1
echo "full ps rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"wgrad_tc" -s 4 -c 4 \
   -o gpurun_out/prof_wg -f python bench.py --profile-pass --batch 128 > gpurun_out/ncu_full_wg.log 2>&1
echo "full wg rc=$?"; ls -la gpurun_out/*.ncu-rep
