#!/bin/bash
# One gpurun call: the GPU parity suite only (fast feedback).  usage: gpurun --timeout 1200 -- 'bash scripts/gpu_tests.sh <tag> [pytest args]'
TAG=${1:-r02}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt; nproc >> gpurun_out/gpus_$TAG.txt
timeout 1100 python -m pytest tests -q -m gpu --timeout 600 "$@" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
grep -E '^(FAILED|ERROR)|passed|failed' gpurun_out/pytest_$TAG.log | tail -60
