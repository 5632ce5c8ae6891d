#!/bin/bash
# round 2, first GPU call: acceptance run of the two default-off experiments + box facts
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,memory.total --format=csv; nproc; free -g; nvidia-smi topo -m; lscpu | head -25; } > gpurun_out/${TAG}_box.txt 2>&1
DNB_SEG_PARITY_SCAN=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_parity_scan.log 2>&1; echo "pytest (DNB_SEG_PARITY_SCAN=1) rc=$?" | tee -a gpurun_out/${TAG}_pytest_parity_scan.log
tail -8 gpurun_out/${TAG}_pytest_parity_scan.log
DNB_EA_WINDOW_PARALLEL=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_wp.log 2>&1; echo "pytest (wp on) rc=$?" | tee -a gpurun_out/${TAG}_pytest_wp.log
tail -8 gpurun_out/${TAG}_pytest_wp.log
DNB_EA_WINDOW_PARALLEL=1 timeout 120 python scripts/wp_check.py 2>&1 | tail -3
timeout 300 python scripts/quick_perf.py 2000 30000 1 2>&1 | grep -E "run 2|->" | tail -2
DNB_SEG_PARITY_SCAN=1 timeout 300 python scripts/quick_perf.py 2000 30000 1 2>&1 | grep -E "run 2|->|status" | tail -3
timeout 400 python tests/helpers/ea_perf.py 1000 10000 0 8 > gpurun_out/${TAG}_ea_perf_serial.json 2> gpurun_out/${TAG}_ea_perf_serial.err; echo "serial rc=$?"
DNB_EA_WINDOW_PARALLEL=1 timeout 400 python tests/helpers/ea_perf.py 1000 10000 0 8 > gpurun_out/${TAG}_ea_perf_wp.json 2> gpurun_out/${TAG}_ea_perf_wp.err; echo "wp rc=$?"
python - <<PY
import json
for m in ("serial", "wp"):
    try:
        d = json.load(open("gpurun_out/${TAG}_ea_perf_%s.json" % m))
        print(m, {k: d[k] for k in ("eventalign", "eventalign_saturated") if k in d}, d["resident_chain"]["stage2"])
    except Exception as ex:
        print(m, "no result:", ex)
PY
echo done
