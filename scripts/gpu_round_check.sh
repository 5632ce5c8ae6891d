#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, the default bench (both arms), and the launch list of a short bench.
# usage: scripts/gpu_round_check.sh <tag>      outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-chk}
KRE='regex:seg_|ranks_kernel|quantile_kernel|scale_events|theil_sen|compact_align|align_kernel|hmm_'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; free -g >> gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --reads 2000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
echo done
