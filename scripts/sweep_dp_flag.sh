#!/bin/bash
# GPU box: rebuild the alignment kernel with -D$1=<each remaining arg> and time the pipeline
name=$1; shift
for val in "$@"; do
  touch dnascent_b200/csrc/banded_dp.cu
  make -s -C dnascent_b200/csrc EXTRA_NVFLAGS="-D$name=$val" > /dev/null 2>&1
  echo "== $name=$val: $(grep -A2 'align_kernelILi0' dnascent_b200/lib/obj/banded_dp.ptxas.log | grep Used)"
  python scripts/quick_perf.py 500 10000 4 2>&1 | grep -E "run 2" | sed 's/.*banded_dp/banded_dp/' | cut -c1-60
  python bench.py --reads 20000 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['stage_ms_per_step']['banded_dp'])"
done
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
