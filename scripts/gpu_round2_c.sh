#!/bin/bash
# GPU call C: fixed tests, ncu launch lists (C1 scoring pass, bf16 + fp32 finetune steps), full-set capture of the bf16 kernels
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_conv_bf16_gpu.py -q -m gpu --timeout=400 > gpurun_out/pytest_bf16.log 2>&1
echo "== bf16 suite rc=$?"; tail -12 gpurun_out/pytest_bf16.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --timeout=600 -k "pruned_c1 or compat_surface or threshold or ddim" > gpurun_out/pytest_sel.log 2>&1
echo "== selected rc=$?"; tail -12 gpurun_out/pytest_sel.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c1_pass.csv python bench.py --profile-pass > gpurun_out/ncu_launch.log 2>&1
echo "launch list c1 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_ft_bf16.csv python scripts/gpu_prof_finetune.py c1 bf16 > gpurun_out/ncu_launch_bf16.log 2>&1
echo "launch list bf16 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_ft_fp32.csv python scripts/gpu_prof_finetune.py c1 fp32 > gpurun_out/ncu_launch_fp32.log 2>&1
echo "launch list fp32 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_bf16|wgrad_bf16" -s 20 -c 10 \
   -o gpurun_out/prof_bf16 -f python scripts/gpu_prof_finetune.py c1 bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
for f in launches_c1_pass launches_ft_bf16 launches_ft_fp32; do echo "## $f"; python tools/launch_summary.py gpurun_out/$f.csv | head -30; done
