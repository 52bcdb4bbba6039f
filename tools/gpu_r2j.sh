#!/bin/bash
# N-GPU bench only (argument: N)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1]
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print('ragged_dp', json.dumps(d.get('ragged_dp'))[:700])
PY
