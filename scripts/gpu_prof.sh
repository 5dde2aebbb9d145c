#!/bin/bash
# ncu launch list + one full capture of the fused step kernels on the BASELINE workload (graph replay)
TAG=${1:-fused}
mkdir -p gpurun_out
python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 60 -c 120 --csv --log-file gpurun_out/launches_$TAG.csv \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_tail_fused' -s 40 -c 4 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/source_$TAG.csv 2>/dev/null
cat gpurun_out/plain_$TAG.log
