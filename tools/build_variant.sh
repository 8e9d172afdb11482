#!/bin/bash
# build_variant.sh NAME "<nvcc -D flags>": a side-by-side engine build that differs only in the QuadLoss / G<=8 sweep
# kernels (the ones config 2 runs) -> lowrankmodels.jl_b200/csrc/variants/libglrm_b200_NAME.so  (tools/tune.py LIB=...)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../lowrankmodels.jl_b200/csrc"
mkdir -p variants build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC $FLAGS \
  -DGLRM_INST_LOSS=1 -DGLRM_INST_WIDE=0 -DGLRM_INST_NAME=launch_quad_narrow -c -o variants/sweep_quad_narrow_$NAME.o sweep_inst.cu
OBJS=$(ls build/*.o | grep -v sweep_quad_narrow.o)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o variants/libglrm_b200_$NAME.so $OBJS variants/sweep_quad_narrow_$NAME.o -ldl
rm -f variants/sweep_quad_narrow_$NAME.o
echo built variants/libglrm_b200_$NAME.so
