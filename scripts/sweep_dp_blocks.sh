#!/bin/bash
# GPU box: rebuild the alignment kernel with different occupancy targets and time the pipeline
for mb in 4 5 6; do
  touch dnascent_b200/csrc/banded_dp.cu
  make -s -C dnascent_b200/csrc EXTRA_NVFLAGS="-DDP_MIN_BLOCKS=$mb" > /dev/null 2>&1
  echo "== DP_MIN_BLOCKS=$mb: $(grep -A2 'align_kernelILi0' dnascent_b200/lib/obj/banded_dp.ptxas.log | grep Used)"
  python scripts/quick_perf.py 500 10000 4 2>&1 | grep -E "run 2" | sed 's/.*banded_dp/banded_dp/' | cut -c1-80
  python bench.py --reads 20000 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['stage_ms_per_step'])"
done
