#!/bin/bash
# round 2, last evidence run: ncu capture of the final build's kernels (constants for the roofline) + the default bench
set -u
TAG=${1:-r2zz}
mkdir -p gpurun_out
COMMON="--set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on"
BARGS="--steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 --ultra-reads 0 --bin-samples 1e12"
timeout 600 ncu $COMMON -k "regex:seg_scan_kernel|seg_tile_kernel|align_kernel" -c 3 -o gpurun_out/${TAG}_fused \
    python bench.py --reads 6000 --analogue-reads 0 $BARGS > gpurun_out/${TAG}_fused.log 2>&1; echo "ncu fused rc=$?"
ncu -i gpurun_out/${TAG}_fused.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_raw.csv 2>/dev/null
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "parity", d["parity_check"]["mismatches"],
      "chain", round(d["chain"]["value"]), "analogue", round(d["analogue"]["value"]), "ultra", round(d["ultra_long"]["value"]))
print({k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
PY
