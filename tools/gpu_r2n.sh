#!/bin/bash
# verified hand-over of the recurrent kernels: protocol test, A/B whole-step timing, stress, then the whole GPU suite + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "rnn" > gpurun_out/n_rnn_tests.log 2>&1; echo "rnn tests rc=$?"; tail -3 gpurun_out/n_rnn_tests.log
quick() {
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-isolation --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - "$*" <<PY
import json,sys
d=json.load(open('gpurun_out/bench_q.json'))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'loss', d['loss'], 'launches', d.get('gpu_launches'))
PY
}
quick ASRB_RNN_DBG=0
quick ASRB_RNN_DBG=4096
timeout 300 python tools/stress_fullsize.py 0 80 2>&1 | tail -4
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/n_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/n_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2h.json
timeout 200 python tools/trace_rnn.py > gpurun_out/n_trace.txt 2>&1; grep "^##" gpurun_out/n_trace.txt
