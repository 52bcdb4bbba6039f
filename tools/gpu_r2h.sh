#!/bin/bash
# sweep of the side-stream GEMM CTA limit (weight gradients under the backward recurrence)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in ${@:-44 56 64 67 76}; do
  ASRB_WGRAD_CTAS=$c timeout 600 python bench.py --steps 8 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c$c.json'))
print('ctas', $c, 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))
PY
done
