#!/bin/bash
# round 2: GPU suite on the new host path (direct DMA, q2r runs, compact results, wp default) + a reduced bench in both
# segmentation modes + the host-phase trace of the e2e leg
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --reads 30000 --steps 3 --warmup 2 > gpurun_out/${TAG}_bench30k.json 2> gpurun_out/${TAG}_bench30k.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench30k.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench30k.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "h2d", d["e2e"]["h2d_bytes_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"], d["e2e"]["host_register_s"])
    print("stage", d["config"]["stage_ms_per_step"])
    print("parity", d["parity_check"])
    print("chain", d["chain"]["value"], d["chain"]["eventalign_kernel_ms"], d["chain"]["ms_per_pass"])
except Exception as ex:
    print("no bench json", ex)
PY
DNB_SEG_PARITY_SCAN=1 timeout 900 python bench.py --reads 30000 --steps 3 --warmup 2 --no-cpu-baseline --chain-reads 0 --parity-reads 0 > gpurun_out/${TAG}_bench30k_scan.json 2> gpurun_out/${TAG}_bench30k_scan.err; echo "bench(scan) rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench30k_scan.json"))
    print("SCAN value", d["value"], "e2e", d["e2e"]["value"])
    print("stage", d["config"]["stage_ms_per_step"])
except Exception as ex:
    print("no bench json", ex)
PY
DNB_TRACE_HOST=1 timeout 600 python bench.py --reads 10000 --steps 1 --warmup 1 --no-cpu-baseline --chain-reads 0 --parity-reads 0 > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err; echo "trace rc=$?"
grep -c "dnb host" gpurun_out/${TAG}_trace.err
echo done
