#!/bin/sh
# The round's standard profiling pass for bench.py on ONE B200 (run under gpurun, after `python bench.py`
# has exited 0 without ncu).  Writes into gpurun_out/; copy what should be judged into profiles/.
#   gpurun --timeout 600 -- 'sh tools/profile_bench.sh [kernel-regex]'
# 1. launch list (gpu__time_duration.sum per launch of our kernels; cold-cache, serialised: shares only)
# 2. one --set full capture of the dominant kernel (default k_fast_probe), with source correlation
# 3. the text summary tools/ncu_summary.py makes of it (headline metrics, stall samples per source line)
set -e
K="${1:-k_fast_probe}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(fast|impute|classify)" -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^${K}" -s 4 -c 1 -f -o gpurun_out/${K} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${K}.ncu-rep 40 > gpurun_out/ncu_summary_${K}.txt
head -25 gpurun_out/ncu_summary_${K}.txt
