#!/bin/bash
# round 2: e2e pipeline settings at the full 100k-read workload (compute slots x submission size x order), one generation
timeout 1400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-reads 0 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 \
   --e2e-sweep "2:8e8:8,3:8e8:8,4:8e8:8,3:8e8:8:interleave,2:8e8:8:interleave,3:1.2e9:8,4:5e8:12,3:5e8:12:interleave" \
   > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
grep "e2e sweep" gpurun_out/r2y_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
