#!/bin/bash
# round 2: uniform-branch slot updates in the band fill, radix ranking in Theil-Sen, out-of-line exp/log in the forward kernel
set -u
TAG=${1:-r2k}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "value ms", round(d["ms_per_step"]), "e2e ms", round(d["e2e"]["ms_per_step"]))
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
    a = d.get("analogue")
    if a: print("analogue", round(a["value"]), a["ms_per_pass"], a["forward_kernel_ms"], a["sites_kernel_ms"])
    print("parity", d.get("parity_check") and d["parity_check"]["mismatches"])
except Exception as ex:
    print("no bench json", ex)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --reads 30000 --steps 4 --warmup 3 --no-cpu-baseline --chain-reads 0 --parity-reads 32 --ultra-reads 0 > gpurun_out/${TAG}_bench30k.json 2> gpurun_out/${TAG}_bench30k.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${TAG}_bench30k.err
show gpurun_out/${TAG}_bench30k.json
echo done
