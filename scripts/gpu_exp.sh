#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -5 gpurun_out/test_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-finetune > gpurun_out/bench_x.json 2>gpurun_out/bench.err
python -c "
import json,re;d=json.load(open('gpurun_out/bench_x.json'));print('ms/pass',round(d['ms_per_step'],2),d['roofline']['breakdown_ms'], d['gpu_launches'])"
