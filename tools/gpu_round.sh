#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, isolation benchmarks, ncu launch list and full captures.
# Usage (under gpurun): tools/gpu_round.sh [tests] [bench] [kernels] [launches] [full] [overlap] [ctc5]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${@:-"tests bench kernels launches full"}
for w in $what; do
  case $w in
    tests)    tools/gpu_checks.sh gemm conv conv32 conv1 rows rnn_simt rnn_tc ctc stft lookahead model smoke ;;
    bench)    timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json ;;
    kernels)  timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; echo "kernels exit=$?"; cat gpurun_out/kernels.json ;;
    launches)
              # single-stream form (ncu serialises the launches anyway): the same pass bench.py times its rooflines in
              ASRB_WGRAD_OVERLAP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
                --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1; echo "launches exit=$?" ;;
    full)
              # one fwd + one bwd recurrent launch, two GEMMs (tf32 in-proj, bf16 backward), the CTC kernels: full shape
              timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'rnn_rec_kernel' -s 4 -c 2 -f -o gpurun_out/prof_rnn python tools/profile_step.py > gpurun_out/full_rnn.log 2>&1; echo "full rnn exit=$?"
              timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'gemm_tn_tf32_kernel' -s 1 -c 1 -f -o gpurun_out/prof_gemm python tools/profile_step.py > gpurun_out/full_gemm.log 2>&1; echo "full gemm exit=$?"
              timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'gemm_tn_tf32_kernel' -s 30 -c 1 -f -o gpurun_out/prof_gemm_bf16 python tools/profile_step.py > gpurun_out/full_gemm2.log 2>&1; echo "full gemm bf16 exit=$?"
              timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'ctc_' -c 4 -f -o gpurun_out/prof_ctc python tools/profile_step.py > gpurun_out/full_ctc.log 2>&1; echo "full ctc exit=$?" ;;
    overlap)  timeout 300 python tools/bench_overlap.py > gpurun_out/overlap.txt 2>&1; echo "overlap exit=$?"; cat gpurun_out/overlap.txt ;;
    ctc5)     timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_alpha_warp|ctc_bwd_kernel' -s 2 -c 2 -f \
                -o gpurun_out/prof_ctc5 python tools/bench_kernels.py ctc > gpurun_out/full_ctc5.log 2>&1; echo "ctc5 exit=$?" ;;
  esac
done
