#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one bench run + one `--set full` capture of the hot kernels.
# usage: scripts/gpu_profile.sh <tag>      outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-prof}
KRE='regex:seg_|ranks_kernel|quantile_kernel|scale_events|banded_dp|backtrace|theil_sen|compact_align|align_kernel'
mkdir -p gpurun_out
# (1) every launch of OUR kernels in one short bench run (cold-cache, serialised: compare shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --reads 2000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# (2) full capture of the second pipeline run of a 2000 x 10 kb batch (the first run is the cold one)
ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k "$KRE" \
    -s 11 -c 11 -o gpurun_out/${TAG}_full python scripts/quick_perf.py 500 10000 4 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
echo done
