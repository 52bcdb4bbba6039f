#!/bin/bash
# full GPU test-suite + bench (new legs) + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'loss', d['loss'], 'launches', d['gpu_launches'])
print('kernel ms', d['kernel_ms_per_step'])
print('torch_cudnn', d.get('torch_cudnn'))
print('cpu', d.get('cpu_baseline'))
for k in ('asrb_ctc_cfg5_fwd','asrb_ctc_cfg5_bwd','asrb_ctc_cfg5_fwd_bwd','asrb_spectrogram','asrb_rnn_fwd','asrb_rnn_bwd'):
    print(k, d['rooflines'].get(k))
PY
