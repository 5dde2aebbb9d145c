#!/bin/bash
mkdir -p gpurun_out
./scripts/microbench_cold_read.bin 3452978 > gpurun_out/microbench_3p5M.log 2>&1; echo "microbench rc=$?"
./scripts/microbench_cold_read.bin 8400000 > gpurun_out/microbench_8p4M.log 2>&1
python scripts/kstep_ab.py --steps 240 "ESIM_STEP_V=1" "ESIM_STEP_V=2" "ESIM_STEP_V=3" "ESIM_STEP_V=4" "ESIM_STEP_V=2 ESIM_STEP_BLOCKS=8" "ESIM_STEP_V=1 ESIM_STEP_PF=0" "ESIM_STEP_V=2 ESIM_STEP_PF=0" > gpurun_out/ab1.log 2>&1
python scripts/kstep_ab.py --steps 480 --exposure-chance 0.004 "ESIM_STEP_V=1" "ESIM_STEP_V=2" "ESIM_STEP_V=3" "ESIM_STEP_V=4" > gpurun_out/ab1_fast.log 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_exp1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp1.log
cat gpurun_out/microbench_3p5M.log gpurun_out/ab1.log gpurun_out/ab1_fast.log
