#!/bin/bash
# quick weak-scaling look: graph replay (warm) per-step time at the given rank counts, BASELINE per-GPU workload
STEPS=${STEPS:-480}
for N in "$@"; do
  if [ "$N" = 1 ]; then python bench.py --gpus 1 --steps $STEPS --warmup 24 --no-cpu-baseline > gpurun_out/q_$N.json 2> gpurun_out/q_$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps $STEPS --warmup 24 > gpurun_out/q_$N.json 2> gpurun_out/q_$N.err; fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/q_$N.json').read().strip().splitlines()[-1])
    print($N, 'value %.3e (%.1f us/step) graph %.3e (%.1f us/step) e2e %.3e'%(d['value'], d['ms_per_step']*1e3, d['value_graph_replay'], d['graph_replay_ms_per_step']*1e3, d['e2e']['value']), {k: round(v/d['config']['steps_executed']*1e6,1) for k,v in d['kernel_seconds'].items()})
except Exception as e:
    print($N, 'failed', e); print(open('gpurun_out/q_$N.err').read()[-2000:])
PY
done
