#!/bin/bash
mkdir -p gpurun_out
for sk in 0 4 6; do
echo "pt skip $sk: $(ONE_SHAPE=1 DPB200_TC_PERSISTENT=4 DPB200_TC_DEBUG_SKIP=$sk timeout 60 python scripts/time_conv_shapes.py 2>&1 | cut -c1-44)"
done
