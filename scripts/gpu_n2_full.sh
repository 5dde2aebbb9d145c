#!/bin/bash
# 2 GPUs: parity suite (incl. the NCCL two-process test), timeline, bench line
TAG=${1:-n2}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E '^(FAILED|ERROR)|passed|failed' gpurun_out/pytest_$TAG.log | tail -5
bash scripts/gpu_ktrace_n.sh kt_$TAG $N > gpurun_out/kt_$TAG.txt 2>&1
grep -A9 "step-to-step [34][0-9][0-9][0-9][0-9]," gpurun_out/kt_$TAG.txt | tail -22 | cut -c1-170
