#!/bin/bash
# round 2: GPU suite with the analogue stage + 1 Mb read, and the bench with every leg at 30k reads
set -u
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
timeout 1500 python bench.py --reads 30000 --steps 4 --warmup 3 > gpurun_out/${TAG}_bench30k.json 2> gpurun_out/${TAG}_bench30k.err; echo "bench rc=$?"
tail -c 800 gpurun_out/${TAG}_bench30k.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench30k.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"]), "value ms", round(d["ms_per_step"]))
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
    hp = d["e2e"]["host_phases"]
    print("host", {k: v for k, v in hp["per_step_s"].items() if v > 0.005}, hp["counters_timed_region"])
    print("chain", d["chain"] and (round(d["chain"]["value"]), d["chain"]["eventalign_kernel_ms"], d["chain"]["ms_per_pass"], d["chain"]["cpu_reference"]))
    print("analogue", d["analogue"])
    u = d["ultra_long"]
    print("ultra", u and (round(u["value"]), u["ms_per_pass"], u["stage_ms_per_pass"], u["bins"]))
    print("parity", d["parity_check"])
    print("cpu", d["cpu_baseline"])
except Exception as ex:
    print("no bench json", ex)
PY
echo done
