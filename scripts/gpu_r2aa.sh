#!/bin/bash
# round 2: lean band bookkeeping (shipped) against the literal forms (lean0), then the parity tests on the shipped build
bash scripts/gpu_variants.sh r2aa 30000 lean0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wire_formats_gpu.py -x -q 2>&1 | tail -3
