#!/bin/bash
# GPU call A of round 2: full parity suite (bf16 file last, under its own tight timeout), then the two bench configs and the reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 --ignore=tests/test_conv_bf16_gpu.py > gpurun_out/pytest_main.log 2>&1
echo "== main suite rc=$?"; tail -30 gpurun_out/pytest_main.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "== bench c1 rc=$?"; tail -c 3000 gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "== bench c3 rc=$?"; tail -c 3000 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c1.json 2> gpurun_out/bench_ref_c1.err
echo "== ref c1 rc=$?"; cat gpurun_out/bench_ref_c1.json
timeout -k 10 400 python -m pytest tests/test_conv_bf16_gpu.py -q -m gpu --timeout=300 > gpurun_out/pytest_bf16.log 2>&1
echo "== bf16 suite rc=$?"; tail -40 gpurun_out/pytest_bf16.log
