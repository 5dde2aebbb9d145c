#!/bin/bash
# BASELINE configs[3] (England-scale 56 M citizens, x = 0.6), strong scaling on N GPUs
N=${1:-2}
bash scripts/gpu_bench_n.sh n${N}_england56_480 $N --config england56 --steps 480 --warmup 24
