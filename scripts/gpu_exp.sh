#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; cut -c1-330 gpurun_out/bench.json
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-200 gpurun_out/bench_ref.json
