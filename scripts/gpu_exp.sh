#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_x.json 2>gpurun_out/bench.err
python -c "
import json,re;d=json.load(open('gpurun_out/bench_x.json'));f=d.get('finetune');print('ms/pass',round(d['ms_per_step'],2),'finetune img/s',round(f['value'],1),'ms',round(f['ms_per_step'],2))"
