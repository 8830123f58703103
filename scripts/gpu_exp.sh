#!/bin/bash
mkdir -p gpurun_out
SANITIZE=0 timeout 900 bash scripts/gpu_check.sh tests/test_conv_tc_gpu.py tests/test_unet_gpu.py
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g2.json 2>gpurun_out/bench.err
python -c "
import json,re;d=json.load(open('gpurun_out/bench_g2.json'));m=re.search(r'device time ([0-9.]+) ms',d['roofline']['note']);print('ms/pass',round(d['ms_per_step'],2),'conv ms',m.group(1),'finetune',d['finetune']['value'])"
