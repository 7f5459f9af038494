#!/bin/sh
# Builds the C-ABI libraries for sm_100a, in-tree (the .so files travel to the GPU box with the snapshot).
#   grimb200.cu    CUDA kernels + C ABI (FP64 operation order: -fmad=false)
#   grimb_text.cpp host text pipeline (no FP contraction either: prior matrices must be bit-exact)
# libgrimb200.so : 64-bit packed haplotype keys (<= 63 key bits: every 5/6-locus table, small 9-locus ones)
# libgrimb200w.so: 128-bit packed keys (wide 9-locus tables); same ABI with GRIMB_KEY_WORDS = 2
# A failed compilation removes the stale library, so that an old build can never pass for the new source.
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false --extended-lambda \
 -Xcompiler -fPIC,-ffp-contract=off,-pthread -shared"
( nvcc $FLAGS -DGRIMB_KW=1 -DGRIMB_KEY_WORDS=1 -o libgrimb200.so.tmp grimb200.cu grimb_text.cpp "$@" \
    && mv libgrimb200.so.tmp libgrimb200.so || { rm -f libgrimb200.so libgrimb200.so.tmp; exit 1; } ) &
P1=$!
( nvcc $FLAGS -DGRIMB_KW=2 -DGRIMB_KEY_WORDS=2 -o libgrimb200w.so.tmp grimb200.cu grimb_text.cpp "$@" \
    && mv libgrimb200w.so.tmp libgrimb200w.so || { rm -f libgrimb200w.so libgrimb200w.so.tmp; exit 1; } ) &
P2=$!
R=0
wait $P1 || R=1
wait $P2 || R=1
if [ $R -ne 0 ]; then echo "build.sh: compilation FAILED" >&2; exit 1; fi
test -f libgrimb200.so && test -f libgrimb200w.so
