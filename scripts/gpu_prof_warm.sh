#!/bin/bash
# ncu full capture of the step kernels WITHOUT cache flushes between replays (closer to the L2-warm graph replay)
TAG=${1:-warm}
mkdir -p gpurun_out
python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --cache-control none --clock-control none --import-source on -k regex:'k_step|k_tail_fused' -s 40 -c 4 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/source_$TAG.csv 2>/dev/null
