#!/bin/bash
# GPU box: rows f1/f2 -- parity tests (features, eventalign, shim), throughput of the stage, one full ncu capture
# usage: scripts/gpu_f2.sh <tag> [all]     ("all" runs the whole GPU suite instead of the f1/f2 files)
set -u
TAG=${1:-f2}
mkdir -p gpurun_out
if [ "${2:-}" = "all" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
else
  timeout 900 python -m pytest tests/test_features_gpu.py tests/test_eventalign_gpu.py tests/test_shim_gpu.py -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
fi
tail -40 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python tests/helpers/ea_perf.py 1000 10000 16 8 > gpurun_out/${TAG}_ea_perf.json 2> gpurun_out/${TAG}_ea_perf.err; echo "ea_perf rc=$?"
tail -c 2500 gpurun_out/${TAG}_ea_perf.json; tail -5 gpurun_out/${TAG}_ea_perf.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:eventalign_kernel|features_kernel" -s 3 -c 2 \
    -o gpurun_out/${TAG}_full python tests/helpers/ea_perf.py 300 10000 0 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
echo done
