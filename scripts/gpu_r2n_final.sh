#!/bin/bash
# round 2, evidence run on one B200: GPU suite, f1 statistical parity, lean ncu captures of the shipped kernels (one
# 6000-read device bin so that a launch == the step and the one-warp-per-CTA launch has 1.7 waves of 3552 warp slots),
# launch list of a short bench, the default bench and the reference arm
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python scripts/ea_statistical_parity.py 2000 80000 > gpurun_out/${TAG}_ea_statistical_parity.json 2> gpurun_out/${TAG}_ea_statistical_parity.err; echo "ea parity rc=$?"
cat gpurun_out/${TAG}_ea_statistical_parity.json | cut -c1-900
COMMON="--set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on"
BARGS="--steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --bin-samples 1e12"
timeout 900 ncu $COMMON -k "regex:seg_scan_kernel|seg_tile_kernel|align_kernel|theil_sen_kernel|quantile_kernel" -c 5 -o gpurun_out/${TAG}_fused \
    python bench.py --reads 6000 --analogue-reads 0 $BARGS > gpurun_out/${TAG}_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu $COMMON -k "regex:llr_forward_kernel|llr_sites_kernel" -c 2 -o gpurun_out/${TAG}_llr \
    python bench.py --reads 500 --analogue-reads 500 $BARGS > gpurun_out/${TAG}_llr.log 2>&1; echo "ncu llr rc=$?"
for f in fused llr; do ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --reads 2000 --steps 1 --warmup 1 --no-cpu-baseline --parity-reads 0 > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "value ms", round(d["ms_per_step"]), "e2e ms", round(d["e2e"]["ms_per_step"]))
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
    print("roofline frac", d["roofline"]["frac"], "seg", d["roofline_segmentation"]["frac"], d["roofline_segmentation"]["issue_view"])
    print("chain", d["chain"] and round(d["chain"]["value"]), "analogue", d["analogue"] and round(d["analogue"]["value"]), "ultra", d["ultra_long"] and round(d["ultra_long"]["value"]))
    print("parity", d["parity_check"])
except Exception as ex:
    print("no bench json", ex)
PY
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
echo done
