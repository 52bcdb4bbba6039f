#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/gpu_checks.sh rnn_tc
timeout 300 python tools/trace_rnn.py > gpurun_out/trace2.txt 2>&1; echo "trace exit=$?"; grep -v "cta \(1\|25\):" gpurun_out/trace2.txt
