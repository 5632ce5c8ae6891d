#!/bin/bash
# GPU box: eventalign + shim parity tests, then (optionally) the rest of the GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eventalign_gpu.py tests/test_shim_gpu.py -x -q > gpurun_out/ea_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/ea_pytest.log
