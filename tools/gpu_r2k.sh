#!/bin/bash
# conv2 forward / data gradient: output rows per work item (1 = one-row kernel, 2, 4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1 2 4; do
  ASRB_CONV_ROWS=$r timeout 600 python bench.py --steps 6 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_rows$r.json 2> gpurun_out/bench_rows$r.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_rows$r.json'))
k=d['kernel_ms_per_step']
print('rows', $r, 'ms/step', round(d['ms_per_step'],2), 'conv32 fwd', k.get('asrb_conv32_fwd'), k.get('asrb_conv32_fwd_rows'), 'dgrad', k.get('asrb_conv32_bwd_data'), k.get('asrb_conv32_bwd_data_rows'), 'loss', d['loss'])
PY
done
