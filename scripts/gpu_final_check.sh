#!/bin/bash
# round-end validation on one B200: GPU test suite, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? wall=${SECONDS}s; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; SECONDS=0; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo rc=$? wall=${SECONDS}s; tail -2 gpurun_out/smoke.log
echo "== bench"; SECONDS=0; python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$? wall=${SECONDS}s; cut -c1-400 gpurun_out/bench.json
echo "== reference arm"; SECONDS=0; python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$? wall=${SECONDS}s; cut -c1-700 gpurun_out/bench_ref.json
