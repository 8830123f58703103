#!/bin/bash
# Runs on the GPU box (via gpurun): builds nothing (the .so travels), runs the gpu tests file by file so one
# sticky CUDA error cannot hide the rest, keeps full logs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch;print(torch.cuda.get_device_name(0))" >> gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -m gpu --maxfail=6 -x --timeout=600 > gpurun_out/${name}.log 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -25 gpurun_out/${name}.log
  if [ $rc -ne 0 ] && [ "${SANITIZE:-1}" = "1" ]; then
    timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "$f" -q -m gpu -x --timeout=500 > gpurun_out/${name}.sanitizer.log 2>&1
    echo "== sanitizer tail"; grep -E "Invalid|Error|at 0x|by thread|dp_|kernel" gpurun_out/${name}.sanitizer.log | head -30
    break
  fi
done
