#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 -" "32 -" "4 trim" "32 trim"; do set -- $cfg
if [ "$2" != "-" ]; then export DPB200_LIB=$PWD/diff-pruning_b200/variants/libdpb200_$2.so; else unset DPB200_LIB; fi
DPB200_PITCH=$1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_x.json 2>gpurun_out/bench.err
python -c "
import json,re;d=json.load(open('gpurun_out/bench_x.json'));f=d.get('finetune');print('pitch $1 lib $2','ms/pass',round(d['ms_per_step'],2),'finetune img/s',round(f['value'],1),'ms',round(f['ms_per_step'],2))"
done
