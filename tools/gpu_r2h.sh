#!/bin/bash
# sweep of the side-stream GEMM CTA limit (weight gradients under the backward recurrence)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "44 0" "44 44" "64 0" "64 64" "56 0"; do
  set -- $c
  ASRB_WGRAD_CTAS=$1 ASRB_WGRAD_CTAS_LAST=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_c$1_$2.json 2> gpurun_out/bench_c$1_$2.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c$1_$2.json'))
print('ctas', $1, 'last', $2, 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))
PY
done
