#!/bin/bash
# bench line + ncu launch list of one eager pass (B200_PROFILING.md recipe)
set -u
mkdir -p gpurun_out
STEPS=${STEPS:-10}
python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-pass --batch ${PBATCH:-128} > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
