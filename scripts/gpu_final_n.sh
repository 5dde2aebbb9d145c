#!/bin/bash
# final multi-GPU lines of the round on N GPUs: BASELINE configs[4] (default) and configs[3]
N=${1:-2}
bash scripts/gpu_bench_n.sh n${N}_uk67_480 $N --steps 480 --warmup 24
bash scripts/gpu_bench_n.sh n${N}_england56_480 $N --config england56 --steps 480 --warmup 24
