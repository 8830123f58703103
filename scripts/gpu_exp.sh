#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_finetune.csv python scripts/gpu_prof_finetune.py > gpurun_out/ncu_ft.log 2>&1
echo rc=$?; python tools/launch_summary.py gpurun_out/launches_finetune.csv | head -40
