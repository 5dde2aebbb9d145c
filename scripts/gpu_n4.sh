#!/bin/bash
bash scripts/gpu_bench_n.sh n4_uk67_480 4 --steps 480 --warmup 24
bash scripts/gpu_bench_n.sh n4_england56_480 4 --config england56 --steps 480 --warmup 24
