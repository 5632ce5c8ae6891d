#!/bin/bash
# GPU box: everything the two default-off experiments (DNB_EA_WINDOW_PARALLEL=1, DNB_SEG_PARITY_SCAN=1) still have to pass
# before either becomes the default: the whole GPU suite with the switch on, the f1/f2 measurement in both modes, one ncu capture of its kernels.
set -u
TAG=${1:-wp}
mkdir -p gpurun_out
DNB_EA_WINDOW_PARALLEL=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest (switch on) rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
DNB_EA_WINDOW_PARALLEL=1 timeout 120 python scripts/wp_check.py 2>&1 | tail -3
timeout 600 python tests/helpers/ea_perf.py 1000 10000 0 8 > gpurun_out/${TAG}_ea_perf_serial.json 2> gpurun_out/${TAG}_ea_perf_serial.err; echo "serial rc=$?"
DNB_EA_WINDOW_PARALLEL=1 timeout 600 python tests/helpers/ea_perf.py 1000 10000 0 8 > gpurun_out/${TAG}_ea_perf_wp.json 2> gpurun_out/${TAG}_ea_perf_wp.err; echo "wp rc=$?"
python - <<PY
import json
for m in ("serial", "wp"):
    try:
        d = json.load(open("gpurun_out/${TAG}_ea_perf_%s.json" % m))
        print(m, {k: d[k] for k in ("eventalign", "eventalign_saturated") if k in d}, d["resident_chain"]["stage2"])
    except Exception as ex:
        print(m, "no result:", ex)
PY
DNB_EA_WINDOW_PARALLEL=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wp_walk_kernel|wp_window_kernel" -s 4 -c 4 \
    -o gpurun_out/${TAG}_full python tests/helpers/ea_perf.py 300 10000 0 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
# ---- the other default-off experiment: block-map scan instead of the per-sample checkpoint chain (seg_scan.cu) ----
DNB_SEG_PARITY_SCAN=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_parity_scan.log 2>&1; echo "pytest (DNB_SEG_PARITY_SCAN=1) rc=$?" | tee -a gpurun_out/${TAG}_pytest_parity_scan.log
tail -8 gpurun_out/${TAG}_pytest_parity_scan.log
timeout 300 python scripts/quick_perf.py 2000 30000 1 2>&1 | grep -E "run 2|->" | tail -2
DNB_SEG_PARITY_SCAN=1 timeout 300 python scripts/quick_perf.py 2000 30000 1 2>&1 | grep -E "run 2|->" | tail -2
echo done
