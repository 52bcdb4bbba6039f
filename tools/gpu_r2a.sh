#!/bin/bash
# round-2 first visit: layout probe, recurrent kernel parity + traces, model tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out

tools/gpu_checks.sh rnn_tc
timeout 300 python tools/trace_rnn.py > gpurun_out/trace2.txt 2>&1; echo "trace exit=$?"; cat gpurun_out/trace2.txt
timeout 300 python tools/trace_rnn.py 751 2 lstm 1024 128 > gpurun_out/trace2_lstm.txt 2>&1; echo "trace lstm exit=$?"; grep "##" gpurun_out/trace2_lstm.txt
tools/gpu_checks.sh rnn_simt model smoke
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['loss']); print(d['kernel_ms_per_step'])"
