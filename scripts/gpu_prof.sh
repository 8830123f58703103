#!/bin/bash
# one `ncu --set full` capture of a named kernel inside one eager bench pass
set -u
mkdir -p gpurun_out
KREGEX=${KREGEX:-conv_tc_kernel}
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$KREGEX -s ${SKIP:-10} -c ${COUNT:-3} \
   -o gpurun_out/prof_${TAG:-tc} -f python bench.py --profile-pass --batch ${PBATCH:-128} > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
