#!/bin/bash
# GPU box: e2e-leg sweep on a 30k-read workload (bin size / batches in flight), one JSON line per config
mkdir -p gpurun_out
: > gpurun_out/e2e_sweep.jsonl
IFS=";" read -ra CFGS <<< "${SWEEP_CFGS:-4.0e8 4}"
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  timeout 400 python bench.py --reads ${SWEEP_READS:-30000} --steps 2 --warmup 1 --no-cpu-baseline --e2e-bin-samples $1 --e2e-inflight $2 \
     2>> gpurun_out/e2e_sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'bin': '$1', 'inflight': $2, 'value': d['value'], 'e2e': d['e2e']['value'], 'e2e_ms': d['e2e']['ms_per_step'], 'ms': d['ms_per_step']}))" >> gpurun_out/e2e_sweep.jsonl
done
cat gpurun_out/e2e_sweep.jsonl
