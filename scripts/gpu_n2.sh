#!/bin/bash
# 2-GPU validation in one gpurun --gpus 2 call: GPU tests that need two devices, the torchrun oracle check in both exchange
# modes, and the bench line at N=2 (weak scaling: the BASELINE per-GPU workload on every rank).
TAG=${1:-n2}
STEPS=${STEPS:-960}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
for COMM in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py --comm $COMM > gpurun_out/sharded_${COMM}_$TAG.log 2>&1
  echo "sharded_check $COMM rc=$?"; grep sharded_check gpurun_out/sharded_${COMM}_$TAG.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps $STEPS --warmup 24 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench N=2 rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
