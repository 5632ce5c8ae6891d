#!/bin/bash
# round 2: NaN-slope path of Theil-Sen (fixtures + 48 synthetic cases through dnb_theil_sen_batch), then the short bench
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "nan or theil or golden or random" 2>&1 | tail -5
timeout 600 python bench.py --reads 30000 --steps 2 --warmup 2 --no-cpu-baseline --parity-reads 16 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 > gpurun_out/r2w_bench30k.json 2> gpurun_out/r2w_bench30k.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2w_bench30k.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()}, "parity", d["parity_check"]["mismatches"])
PY
