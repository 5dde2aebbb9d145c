#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_exp4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp4.log
for B in 0 8 3; do
  echo "ESIM_PT_BLOCKS=$B"
  ESIM_PT_BLOCKS=$B ESIM_KTRACE=1 timeout 120 python scripts/profile_steps.py --steps 960 --skip 24 2>&1 | tail -6
done
