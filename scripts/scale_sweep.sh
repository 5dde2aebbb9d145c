#!/bin/bash
# weak-scaling sweep on one box: N = 1 2 4 8 ranks, one JSON line each into gpurun_out/scale_N.json
STEPS=${STEPS:-960}
EXTRA="$@"
python bench.py --gpus 1 --steps $STEPS --warmup 24 --no-cpu-baseline $EXTRA > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps $STEPS --warmup 24 $EXTRA > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/scale_%d.json'%n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n,'failed',e); continue
    if n==1: base=d
    print(n, 'value %.3e graph %.3e (%.1f us/step) e2e %.3e'%(d['value'], d['value_graph_replay'], d['graph_replay_ms_per_step']*1e3, d['e2e']['value']), 'eff_graph %.2f'%(d['value_graph_replay']/(n*base['value_graph_replay'])) if base else '', d['kernel_seconds'])
PY
