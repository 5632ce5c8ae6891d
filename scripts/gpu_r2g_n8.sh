#!/bin/bash
# round 2, 8 GPUs of one box: multi-device tests, then the literal configs[4] batch (10^6 reads = 125 000 per GPU)
# through bench.py exactly as the driver launches it, with the e2e host-phase table in the JSON line
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=index,name,memory.total --format=csv; nproc; free -g; } > gpurun_out/${TAG}_box.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or two_devices" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -4 gpurun_out/${TAG}_pytest_multi.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --reads 125000 --bin-samples 2.0e9 --steps 2 --warmup 1 --chain-reads 0 --analogue-reads 0 --ultra-reads 0 \
    > gpurun_out/${TAG}_bench_n8_1Mreads.json 2> gpurun_out/${TAG}_bench_n8_1Mreads.err; echo "bench n8 rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_n8_1Mreads.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n8_1Mreads.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "reads/gpu", d["config"]["reads_per_gpu"], "reduced", d["config"]["reads_per_gpu_reduced_for_host_ram"])
    print("e2e", {k: v for k, v in d["e2e"].items() if k != "host_phases"})
    hp = d["e2e"]["host_phases"]
    print("host", {k: v for k, v in hp["per_step_s"].items() if v > 0.01}, hp["counters_timed_region"])
    print("stage", {k: round(v, 1) for k, v in d["config"]["stage_ms_per_step"].items()})
except Exception as ex:
    print("no bench json", ex)
PY
free -g | head -2
echo done
