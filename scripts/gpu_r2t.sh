#!/bin/bash
# round 2: second occupancy sweep (1-warp CTAs are the shipped default now) + one ncu capture of the shipped band fill
bash scripts/gpu_variants.sh r2t 30000 w1b28 w1b32 w1b24x2 w1b24x1 w1b28x2 s12 s14
COMMON="--set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on"
BARGS="--steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --bin-samples 1e12"
DNB_SPLIT_ALIGN=1 timeout 600 ncu $COMMON -k "regex:align_kernel" -c 2 -o gpurun_out/r2t_split \
    python bench.py --reads 2500 --analogue-reads 0 $BARGS > gpurun_out/r2t_split.log 2>&1; echo "ncu split rc=$?"
ncu -i gpurun_out/r2t_split.ncu-rep --page raw --csv > gpurun_out/r2t_split_raw.csv 2>/dev/null
