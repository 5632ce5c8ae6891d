#!/bin/bash
# GPU box: time every prebuilt library variant (scripts/build_variants.sh) on the same short bench; the shipped library is
# restored afterwards.  usage: scripts/gpu_variants.sh <out-tag> <reads> <tag> [<tag> ...]
out=$1; reads=$2; shift 2
mkdir -p gpurun_out
cp dnascent_b200/lib/libdnascent_b200.so /tmp/shipped.so
for tag in shipped "$@"; do
  [ $tag = shipped ] && cp /tmp/shipped.so dnascent_b200/lib/libdnascent_b200.so || cp dnascent_b200/lib/variants/libdnascent_b200_${tag}.so dnascent_b200/lib/libdnascent_b200.so
  timeout 600 python bench.py --reads $reads --steps 2 --warmup 2 --no-cpu-baseline --parity-reads 16 --chain-reads 0 --analogue-reads 0 --ultra-reads ${ULTRA:-0} \
      > gpurun_out/${out}_${tag}.json 2> gpurun_out/${out}_${tag}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${out}_${tag}.json"))
    s = d["config"]["stage_ms_per_step"]
    print("${tag}", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 1) for k, v in s.items() if k not in ("host_step", "wall")},
          "parity", d["parity_check"]["mismatches"], "ultra", d.get("ultra_long") and round(d["ultra_long"]["value"]))
except Exception as ex:
    print("${tag}", "FAILED", ex)
PY
done
cp /tmp/shipped.so dnascent_b200/lib/libdnascent_b200.so
