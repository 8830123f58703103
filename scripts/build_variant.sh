#!/bin/bash
# build_variant.sh NAME -DFLAG...  ->  diff-pruning_b200/variants/libdpb200_NAME.so (conv_tc.cu rebuilt with the extra flags,
# other objects reused from the normal in-tree build).  Select at run time with DPB200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python diff-pruning_b200/build.py > /dev/null
name=$1; shift
mkdir -p diff-pruning_b200/variants
nvcc -DDPB200_HAVE_TC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include -I diff-pruning_b200/csrc "$@" \
  -c diff-pruning_b200/csrc/conv_tc.cu -o diff-pruning_b200/variants/conv_tc_$name.o
objs=$(ls diff-pruning_b200/csrc/*.o | grep -v conv_tc.o)
nvcc -shared -o diff-pruning_b200/variants/libdpb200_$name.so $objs diff-pruning_b200/variants/conv_tc_$name.o -lcudart
echo diff-pruning_b200/variants/libdpb200_$name.so
