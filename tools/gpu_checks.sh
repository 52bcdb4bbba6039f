#!/bin/bash
# Runs the GPU test groups as separate, time-limited processes (a hung kernel in one group must not take the
# others down) and collects logs under gpurun_out/.  Usage: tools/gpu_checks.sh [group ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PYT="python -m pytest -q -rA -p no:cacheprovider --timeout=900"
run() { # name, limit, args...
  local name=$1 limit=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $limit $PYT "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? $(tail -1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
groups=${@:-"gemm conv rows rnn_simt rnn_tc ctc stft lookahead model"}
: > gpurun_out/summary.txt
for g in $groups; do
  case $g in
    gemm)     run gemm 300 tests/test_gpu_kernels.py -k "gemm" ;;
    conv)     run conv 300 tests/test_gpu_kernels.py -k "conv2d or bn_act or layout" ;;
    conv32)   run conv32 300 tests/test_gpu_kernels.py -k "conv32" ;;
    conv1)    run conv1 300 tests/test_gpu_kernels.py -k "conv1_tensor" ;;
    rows)     run rows 300 tests/test_gpu_kernels.py -k "bn_rows or log_softmax" ;;
    rnn_simt) run rnn_simt 600 tests/test_gpu_kernels.py -k "rnn and simt_debug" ;;
    rnn_tc)   run rnn_tc 300 tests/test_gpu_kernels.py -k "rnn and (tf32 or bf16)" ;;
    ctc)      run ctc 300 tests/test_gpu_kernels.py -k "ctc" ;;
    lookahead) run lookahead 300 tests/test_gpu_kernels.py -k "lookahead" ;;
    stft)     run stft 300 tests/test_gpu_kernels.py -k "spectrogram" ;;
    model)    run model 900 tests/test_gpu_model.py ;;
    smoke)    echo "=== smoke" | tee -a gpurun_out/summary.txt; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt ;;
    bench)    echo "=== bench" | tee -a gpurun_out/summary.txt; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$? $(tail -c 600 gpurun_out/bench.log)" | tee -a gpurun_out/summary.txt ;;
  esac
done
cat gpurun_out/summary.txt
