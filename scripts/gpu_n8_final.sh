#!/bin/bash
# 8 GPUs: BASELINE configs[4] (default) and configs[3] (england56), with the driver's flags and with 480 steps
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n8.txt
bash scripts/gpu_bench_n.sh n8_uk67_480 8 --steps 480 --warmup 24
bash scripts/gpu_bench_n.sh n8_uk67_driverflags 8 --steps 20 --warmup 5
bash scripts/gpu_bench_n.sh n8_england56_480 8 --config england56 --steps 480 --warmup 24
