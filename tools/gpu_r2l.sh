#!/bin/bash
# full ncu captures of the conv2 kernels (forward, data gradient, weight gradient) in one training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'conv_rows_tc_kernel|conv_wgrad_cls_kernel|conv1_' -c 5 -f -o gpurun_out/prof_conv python tools/profile_step.py > gpurun_out/full_conv.log 2>&1; echo "full conv exit=$?"
ls -la gpurun_out/prof_conv.ncu-rep
