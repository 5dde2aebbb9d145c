#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the reference arm, the ncu launch list and one full capture.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r01}
mkdir -p gpurun_out
nproc > gpurun_out/nproc_$TAG.txt
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
ESIM_STEP_V=1 ESIM_TAIL_FLAGWAIT=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_v1_$TAG.log 2>&1; echo "pytest (first fused build, grid-dependency tail) rc=$?" | tee -a gpurun_out/pytest_v1_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5000 --warmup 24 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_$TAG.json gpurun_out/bench_ref_$TAG.json
ESIM_KTRACE=1 python scripts/profile_steps.py --steps 960 --skip 24 > gpurun_out/ktrace_$TAG.log 2>&1
python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 60 -c 150 --csv --log-file gpurun_out/launches_$TAG.csv \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_tail_fused|k_pt' -s 30 -c 8 -f -o gpurun_out/prof_$TAG \
    python scripts/profile_steps.py --steps 48 --skip 24 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ls -la gpurun_out | tail -20
