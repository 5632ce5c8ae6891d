#!/bin/bash
# round 2: ncu --set full captures of the shipped kernels on a bench-scale bin (2.5*10^9 samples, the first bin of an
# 8000-read workload of the C2 length law), fused and split align launches, and the analogue forward kernel
set -u
TAG=${1:-r2i}
mkdir -p gpurun_out
COMMON="--set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on"
DNB_SPLIT_ALIGN=1 timeout 900 ncu $COMMON \
    -k "regex:seg_scan_kernel|seg_tile_kernel|align_kernel|theil_sen_kernel|quantile_kernel" -c 6 -o gpurun_out/${TAG}_split \
    python bench.py --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --analogue-reads 0 > gpurun_out/${TAG}_split.log 2>&1
echo "ncu split rc=$?"
timeout 900 ncu $COMMON -k "regex:align_kernel" -c 1 -o gpurun_out/${TAG}_fused \
    python bench.py --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --analogue-reads 0 > gpurun_out/${TAG}_fused.log 2>&1
echo "ncu fused rc=$?"
timeout 900 ncu $COMMON -k "regex:llr_forward_kernel|llr_sites_kernel" -c 2 -o gpurun_out/${TAG}_llr \
    python bench.py --reads 1000 --steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --analogue-reads 1000 > gpurun_out/${TAG}_llr.log 2>&1
echo "ncu llr rc=$?"
for f in split fused llr; do
    ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
done
grep -h '"counts_per_step"' gpurun_out/${TAG}_split.log | tail -c 1200
ls -la gpurun_out/${TAG}_*
echo done
