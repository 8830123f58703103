#!/bin/bash
# 2-GPU check: multi-GPU parity tests + the C1 bench line as the driver launches it (main leg only)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-finetune --no-c3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_2gpu.json
