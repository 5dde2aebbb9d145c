#!/bin/bash
# 2 (or N) GPUs: real ranks against the oracle, then the multi-GPU bench line (BASELINE configs[4] by default).
# usage: gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_n2.sh <tag> <N> [bench args]'
TAG=${1:-n2}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py --comm p2p --areas 400 --cross 0.6 --steps 900 2>&1 | grep -E "sharded_check|MISMATCH|Error|error" | tee gpurun_out/sharded_check_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/sharded_check.py --comm nccl --areas 200 --cross 0.6 --steps 600 2>&1 | grep -E "sharded_check|MISMATCH|Error|error" | tee -a gpurun_out/sharded_check_$TAG.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$TAG.err | cut -c1-300; cat gpurun_out/bench_$TAG.json | cut -c1-3000
