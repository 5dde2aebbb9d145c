#!/bin/bash
# GPU parity suite, then A/B of compile-time variants (epidemicsimulator_b200.build.build_variant).
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E '^(FAILED|ERROR)|passed|failed' gpurun_out/pytest_$TAG.log | tail -20
P=epidemicsimulator_b200
V="ESIM_B200_LIB=$P/libesim_b200.so ESIM_B200_LIB=$P/libesim_b200_pt6.so ESIM_B200_LIB=$P/libesim_b200_pt10.so"
timeout 500 python scripts/kstep_ab.py --steps 240 $V > gpurun_out/ab_$TAG.log 2>&1
timeout 500 python scripts/kstep_ab.py --steps 120 --areas 27500 --cross 0.9 $V > gpurun_out/ab_8p4M_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log gpurun_out/ab_8p4M_$TAG.log
