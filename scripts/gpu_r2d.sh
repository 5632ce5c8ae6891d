#!/bin/bash
# round 2: acceptance of the streaming scan (default) + arena allocation; bench at 30k reads in both segmentation modes
set -u
TAG=${1:-r2d}
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
    if d.get("parity_check"): print("parity", d["parity_check"]["reads"], d["parity_check"]["mismatches"])
except Exception as ex:
    print("no bench json", ex)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --reads 30000 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench30k.json 2> gpurun_out/${TAG}_bench30k.err; echo "bench rc=$?"
tail -c 500 gpurun_out/${TAG}_bench30k.err
show gpurun_out/${TAG}_bench30k.json
DNB_SEG_PARITY_SCAN=0 timeout 900 python bench.py --reads 30000 --steps 4 --warmup 3 --no-cpu-baseline --chain-reads 0 --parity-reads 0 > gpurun_out/${TAG}_bench30k_chain.json 2> gpurun_out/${TAG}_bench30k_chain.err; echo "bench(checkpoint chain) rc=$?"
show gpurun_out/${TAG}_bench30k_chain.json
echo done
