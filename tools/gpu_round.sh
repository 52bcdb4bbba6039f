#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, isolation benchmarks, ncu launch list and full captures.
# Usage (under gpurun): tools/gpu_round.sh [tests] [bench] [kernels] [launches] [full]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${@:-"tests bench kernels launches full"}
for w in $what; do
  case $w in
    tests)    tools/gpu_checks.sh gemm conv conv32 conv1 rows rnn_simt rnn_tc ctc stft model smoke ;;
    bench)    timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json ;;
    kernels)  timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; echo "kernels exit=$?"; cat gpurun_out/kernels.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
                --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1; echo "launches exit=$?" ;;
    full)     timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'rnn_rec_kernel|ctc_' -c 6 -f -o gpurun_out/prof_rnn python tools/profile_step.py --batch 64 --frames 301 > gpurun_out/full.log 2>&1; echo "full exit=$?" ;;
  esac
done
