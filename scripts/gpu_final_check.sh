#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== skip pytest (already green)"
echo "== skip smoke"
echo "== bench"; SECONDS=0; python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$? wall=${SECONDS}s; cut -c1-400 gpurun_out/bench.json
echo "== reference arm"; SECONDS=0; python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$? wall=${SECONDS}s; cut -c1-700 gpurun_out/bench_ref.json
