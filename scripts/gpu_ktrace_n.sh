#!/bin/bash
# device-side timeline of the step kernels on N ranks (ESIM_KTRACE=1 prints it when a handle is destroyed)
TAG=${1:-kt}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
ESIM_KTRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 480 --warmup 24 --no-extra-legs "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/ktrace_$TAG.log; echo "rc=$?"
grep -E "^\[esim\]" gpurun_out/ktrace_$TAG.log | head -120
