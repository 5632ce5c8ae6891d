#!/bin/bash
# GPU box: parity tests + default bench (no CPU baseline) -> gpurun_out/<tag>_*
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,1) for k,v in d['config']['stage_ms_per_step'].items()}, d['config']['counts_per_step']['seg_serial_reads'], d['clocks'])
PY
