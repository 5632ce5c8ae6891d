#!/bin/bash
# round 2: device bin size of the value leg at the full 100k-read workload (wave quantisation of the one-warp-per-read launch)
for bs in 2.5e9 3.75e9 5e9; do
  timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-reads 0 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 --bin-samples $bs \
      > gpurun_out/r2x_bin_$bs.json 2> gpurun_out/r2x_bin_$bs.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2x_bin_$bs.json"))
    print("$bs", "bins", d["config"]["bins"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
except Exception as ex:
    print("$bs FAILED", ex); print(open("gpurun_out/r2x_bin_$bs.err").read()[-600:])
PY
done
