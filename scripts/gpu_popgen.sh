#!/bin/bash
TAG=${1:-pg}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_popgen_device.py -q -m gpu --timeout 600 > gpurun_out/pytest_popgen_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E '^(FAILED|ERROR)|passed|failed|Error|assert' gpurun_out/pytest_popgen_$TAG.log | head -30
