#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/trace_rnn.py > gpurun_out/trace2.txt 2>&1; echo "trace exit=$?"; grep "##\|cta 0" gpurun_out/trace2.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "rnn" -p no:cacheprovider 2>&1 | tail -3
