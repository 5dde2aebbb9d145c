#!/bin/bash
# A/B of library variants on N GPUs: untraced bench lines (no extra legs)
N=${1:-2}; shift
mkdir -p gpurun_out
for V in "$@"; do
  LIB=epidemicsimulator_b200/libesim_b200$V.so
  ESIM_B200_LIB=$LIB timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 480 --warmup 24 --no-extra-legs > gpurun_out/ab_n_tmp.json 2> gpurun_out/ab_n_tmp.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/ab_n_tmp.json') if l.startswith('{')][-1])
    st=d['config']['steps_executed']
    print('variant "$V" N', d['n_gpus'], 'flushed', round(d['ms_per_step']*1e3,2), 'replay', round(d['graph_replay_ms_per_step']*1e3,2), 'tail', round(d['kernel_seconds']['k_tail_fused']/st*1e6,2), 'parity', d['parity_checked'])
except Exception as e:
    print('variant "$V" FAILED', e); print(open('gpurun_out/ab_n_tmp.err').read()[-800:])
PY
done
