#!/bin/sh
# Builds libgrimb200.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
     --extended-lambda -Xcompiler -fPIC -shared -o libgrimb200.so grimb200.cu "$@"
