#!/bin/bash
# 2-GPU A/B of the peer-to-peer tail's fences: oracle check (fused peer-to-peer pipeline incl. vaccination), then the bench line
mkdir -p gpurun_out
for F in 0 1; do
  ESIM_TAIL_FENCE=$F timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$F scripts/sharded_check.py --comm p2p --steps 900 2>&1 | grep sharded_check
  ESIM_TAIL_FENCE=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$F bench.py --gpus 2 --steps 960 --warmup 24 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_n2_fence$F.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_fence$F.json').read())
print('ESIM_TAIL_FENCE=$F value %.4e (%.2f us/step) replay %.4e (%.2f us/step) tail %.2f us' % (d['value'], d['ms_per_step']*1e3, d['value_graph_replay'], d['graph_replay_ms_per_step']*1e3, d['kernel_seconds']['k_tail_fused']/d['config']['steps_executed']*1e6))
PY
done
