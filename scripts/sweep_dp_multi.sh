#!/bin/bash
# GPU box: rebuild the alignment kernel with each "-DA=x -DB=y" argument string and time the band fill at 20k reads
for flags in "$@"; do
  touch dnascent_b200/csrc/banded_dp.cu
  make -s -C dnascent_b200/csrc EXTRA_NVFLAGS="$flags" > /dev/null 2>&1
  echo "== $flags: $(grep -A2 'align_kernelILi0' dnascent_b200/lib/obj/banded_dp.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
  python bench.py --reads 20000 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['config']['stage_ms_per_step']['banded_dp'],1), round(d['config']['stage_ms_per_step']['backtrace'],1))"
done
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
