#!/bin/bash
# plain multi-GPU bench line (no tracing).  usage: gpurun --gpus N -- 'bash scripts/gpu_bench_n.sh <tag> <N> [bench args]'
TAG=${1:-bn}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$TAG.json') if l.startswith('{')][-1])
st=d['config']['steps_executed']
print('N', d['n_gpus'], 'flushed us/step', round(d['ms_per_step']*1e3,2), 'replay us/step', round(d['graph_replay_ms_per_step']*1e3,2), 'value', '%.4e'%d['value'], 'replay', '%.4e'%d['value_graph_replay'], 'parity', d['parity_checked'])
print({k: round(v/st*1e6,2) for k,v in d['kernel_seconds'].items()})
w=d.get('weak_scaling_reference') or d.get('strong_scaling_reference')
if w: print('1-GPU reference (%s): flushed' % ('weak: per-GPU workload' if 'weak_scaling_reference' in d else 'strong: whole population'), round(w['ms_per_step']*1e3,2), 'replay', round(w['graph_replay_ms_per_step']*1e3,2), 'eff flushed %.3f replay %.3f' % (d['value']/(d['n_gpus']*w['value']), d['value_graph_replay']/(d['n_gpus']*w['value_graph_replay'])))
print('e2e', d['e2e']['seconds'], d['e2e']['setup_seconds'], 'popgen', d['config']['population_seconds'])
PY
