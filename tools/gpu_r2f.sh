#!/bin/bash
# ncu evidence for round 2: launch list of one step (single stream) + full captures of the rnn3 kernels, a conv2 kernel each, CTC cfg5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ASRB_WGRAD_OVERLAP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1; echo "launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'rnn_rec3' -s 4 -c 2 -f -o gpurun_out/prof_rnn3 python tools/profile_step.py > gpurun_out/full_rnn3.log 2>&1; echo "full rnn3 exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'conv_row_tc_kernel|conv_wgrad_tc_kernel' -c 3 -f -o gpurun_out/prof_conv python tools/profile_step.py > gpurun_out/full_conv.log 2>&1; echo "full conv exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_alpha_warp|ctc_bwd_kernel' -s 2 -c 2 -f \
   -o gpurun_out/prof_ctc5 python tools/bench_kernels.py ctc > gpurun_out/full_ctc5.log 2>&1; echo "ctc5 exit=$?"
ls -la gpurun_out/*.ncu-rep
