#!/bin/bash
# round 2: default vs block-map scan at bench scale, e2e host phases in steady state, ncu of the shipped kernels
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"]), "value ms", round(d["ms_per_step"]))
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
    hp = d["e2e"]["host_phases"]
    print("host", {k: v for k, v in hp["per_step_s"].items() if v > 0.005}, hp["counters_timed_region"])
    if d.get("chain"): print("chain", round(d["chain"]["value"]), d["chain"]["eventalign_kernel_ms"], d["chain"]["ms_per_pass"])
except Exception as ex:
    print("no bench json", ex)
PY
}
timeout 900 python bench.py --reads 30000 --steps 4 --warmup 3 --no-cpu-baseline --parity-reads 0 > gpurun_out/${TAG}_bench30k.json 2> gpurun_out/${TAG}_bench30k.err; echo "bench rc=$?"
show gpurun_out/${TAG}_bench30k.json
DNB_SEG_PARITY_SCAN=1 timeout 900 python bench.py --reads 30000 --steps 4 --warmup 3 --no-cpu-baseline --chain-reads 0 --parity-reads 0 > gpurun_out/${TAG}_bench30k_scan.json 2> gpurun_out/${TAG}_bench30k_scan.err; echo "bench(scan) rc=$?"
tail -c 400 gpurun_out/${TAG}_bench30k_scan.err
show gpurun_out/${TAG}_bench30k_scan.json
timeout 900 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on \
    -k "regex:align_kernel|seg_tile_kernel|seg_checkpoint_kernel|theil_sen|quantile_kernel" -s 5 -c 5 -o gpurun_out/${TAG}_full \
    python bench.py --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_full* | head
echo done
