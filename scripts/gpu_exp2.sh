#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_exp2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp2.log
ESIM_STEP_V=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_exp2_v1.log 2>&1; echo "pytest V=1 rc=$?"; tail -2 gpurun_out/pytest_exp2_v1.log
timeout 400 python scripts/kstep_ab.py --steps 240 "ESIM_TAIL_FLAGWAIT=0" "ESIM_TAIL_FLAGWAIT=1" "ESIM_TAIL_FLAGWAIT=1 ESIM_STEP_V=1" "ESIM_TAIL_FLAGWAIT=0 ESIM_STEP_V=1" > gpurun_out/ab2.log 2>&1
timeout 300 python scripts/kstep_ab.py --steps 480 --exposure-chance 0.004 "ESIM_TAIL_FLAGWAIT=0" "ESIM_TAIL_FLAGWAIT=1" > gpurun_out/ab2_fast.log 2>&1
ESIM_KTRACE=1 timeout 120 python scripts/profile_steps.py --steps 960 --skip 24 > gpurun_out/ktrace_exp2.log 2>&1
cat gpurun_out/ab2.log gpurun_out/ab2_fast.log gpurun_out/ktrace_exp2.log
