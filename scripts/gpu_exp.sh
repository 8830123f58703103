#!/bin/bash
mkdir -p gpurun_out
for ce in 16384 8192 4096 2048; do
DPB200_GN_CHUNK_ELEMS=$ce timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-finetune > gpurun_out/bench_x.json 2>gpurun_out/bench.err
python -c "
import json,re;d=json.load(open('gpurun_out/bench_x.json'));print('gn chunk elems $ce','ms/pass',round(d['ms_per_step'],3))"
done
