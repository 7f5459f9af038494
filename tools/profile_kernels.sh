#!/bin/sh
# ncu --set full captures of the kernels behind bench.py's sub-records, ONE launch each (run under gpurun after
# `python bench.py` has exited 0 without ncu).  Usage: sh tools/profile_kernels.sh "k_fast_score:4 k_impute_typed:1 k_impute:3"
# (kernel-name regex : launches of that kernel to skip).  Summaries land in gpurun_out/ncu_summary_<kernel>.txt.
set -e
mkdir -p gpurun_out
for spec in ${1:-"k_fast_score:4 k_impute_typed:1 k_impute:3"}; do
  K="${spec%%:*}"; SK="${spec##*:}"
  ncu --set full --clock-control none --import-source on -k "regex:^${K}\$|^${K}[(<]" -s "$SK" -c 1 -f -o gpurun_out/${K} \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${K}.log 2>&1 || { tail -5 gpurun_out/ncu_full_${K}.log; continue; }
  python tools/ncu_summary.py gpurun_out/${K}.ncu-rep 40 > gpurun_out/ncu_summary_${K}.txt
  head -24 gpurun_out/ncu_summary_${K}.txt
done
