#!/bin/bash
# One gpurun call (1 GPU): GPU parity tests, the default bench line, the reference arm, the device timeline, the ncu launch list
# and one full capture.   usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r02}
mkdir -p gpurun_out
nproc > gpurun_out/nproc_$TAG.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
grep -E '^(FAILED|ERROR)|passed|failed' gpurun_out/pytest_$TAG.log | tail -8
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20_$TAG.json 2> gpurun_out/bench20_$TAG.err; echo "bench (driver's flags) rc=$?"
timeout 600 python bench.py --impl reference --steps 5000 --warmup 24 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_$TAG.json | cut -c1-1200
ESIM_KTRACE=1 python scripts/profile_steps.py --steps 960 --skip 24 > gpurun_out/ktrace_$TAG.log 2>&1
python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 60 -c 150 --csv --log-file gpurun_out/launches_$TAG.csv \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_tail_fused|k_pt' -s 30 -c 8 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ls -la gpurun_out | tail -12
