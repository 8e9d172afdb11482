#!/bin/bash
# Builds the CUDA engine (sm_100a only) and the host-side synthetic-data helper, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -ccbin /usr/bin/g++ -Xcompiler -fPIC -shared ${GLRM_NVCC_EXTRA} \
      -o libglrm_b200.so glrm_engine.cu -ldl
/usr/bin/gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o libglrm_synth.so synth_pattern.c
