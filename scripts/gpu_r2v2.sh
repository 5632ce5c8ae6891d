#!/bin/bash
# round 2: value leg with 2 / 3 resident bins running at a time (full 100k-read workload)
for c in 2 3; do
  timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-reads 0 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 --value-inflight $c \
      > gpurun_out/r2v2_inflight$c.json 2> gpurun_out/r2v2_inflight$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2v2_inflight$c.json"))
    print("inflight $c", "bins", d["config"]["bins"], "value", round(d["value"]), "sequential", d["config"]["sequential_pass"] and round(d["config"]["sequential_pass"]["value"]), "e2e", round(d["e2e"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
except Exception as ex:
    print("$c FAILED", ex); print(open("gpurun_out/r2v2_inflight$c.err").read()[-800:])
PY
done
