#!/bin/bash
# final profiling pass: (1) launch list of one eager pass, (2) DRAM traffic of every tensor-core conv launch of that pass,
# (3) one full-set capture of the persistent conv kernel on big 3x3 layers
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -k regex:'conv_tc|wgrad_tc' --csv --log-file gpurun_out/conv_traffic.csv python bench.py --profile-pass --batch 128 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_ps|wgrad_tc" -s 6 -c 8 \
   -o gpurun_out/prof_ps -f python bench.py --profile-pass --batch 128 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
