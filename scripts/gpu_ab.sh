#!/bin/bash
# GPU parity suite, then A/B of compile-time variants (epidemicsimulator_b200.build.build_variant).
TAG=${1:-r02}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E '^(FAILED|ERROR)|passed|failed' gpurun_out/pytest_$TAG.log | tail -20
P=epidemicsimulator_b200
V=""
for name in "$@"; do V="$V ESIM_B200_LIB=$P/libesim_b200$name.so"; done
timeout 500 python scripts/kstep_ab.py --steps 240 $V > gpurun_out/ab_$TAG.log 2>&1
timeout 500 python scripts/kstep_ab.py --steps 120 --areas 27500 --cross 0.9 $V > gpurun_out/ab_8p4M_$TAG.log 2>&1
timeout 500 python scripts/kstep_ab.py --steps 120 --peak-mix $V > gpurun_out/ab_peak_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log gpurun_out/ab_8p4M_$TAG.log gpurun_out/ab_peak_$TAG.log
