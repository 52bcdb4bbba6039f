#!/bin/bash
# round-2 closing profiles: launch list of one training step, full ncu captures of the two recurrent kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ASRB_WGRAD_OVERLAP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1; echo "launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'rnn_rec3' -s 4 -c 2 -f -o gpurun_out/prof_rnn3 python tools/profile_step.py > gpurun_out/full_rnn3.log 2>&1; echo "full rnn3 exit=$?"
ls -la gpurun_out/prof_rnn3.ncu-rep gpurun_out/launches.csv
