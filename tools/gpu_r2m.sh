#!/bin/bash
# input projection overlapped with the forward recurrence: blocks / pairs on the main stream / CTA cap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - "$*" <<PY
import json,sys
d=json.load(open('gpurun_out/bench_q.json'))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'loss', d['loss'])
PY
}
run ASRB_INPROJ_OVERLAP=0
run ASRB_INPROJ_BLOCKS=6 ASRB_INPROJ_MAIN_PAIRS=2
run ASRB_INPROJ_BLOCKS=6 ASRB_INPROJ_MAIN_PAIRS=1
run ASRB_INPROJ_BLOCKS=8 ASRB_INPROJ_MAIN_PAIRS=3
run ASRB_INPROJ_BLOCKS=8 ASRB_INPROJ_MAIN_PAIRS=2
run ASRB_INPROJ_BLOCKS=10 ASRB_INPROJ_MAIN_PAIRS=4
