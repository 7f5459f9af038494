#!/bin/sh
# Builds libgrimb200.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
#   grimb200.cu    CUDA kernels + C ABI (FP64 operation order: -fmad=false)
#   grimb_text.cpp host text pipeline (no FP contraction either: prior matrices must be bit-exact)
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
     --extended-lambda -Xcompiler -fPIC,-ffp-contract=off,-pthread -shared \
     -o libgrimb200.so grimb200.cu grimb_text.cpp "$@"
