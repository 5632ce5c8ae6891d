#!/bin/bash
# round 2, 2 GPUs of one box: the multi-device tests (skipped on a 1-GPU box) and the bench exactly as the driver launches it at N=2
set -u
TAG=${1:-r2n2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or two_devices" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --reads ${READS:-40000} --steps 2 --warmup 1 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "reads/gpu", d["config"]["reads_per_gpu"], "parity", d["parity_check"])
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
except Exception as ex:
    print("no bench json", ex)
PY
