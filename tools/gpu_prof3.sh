#!/bin/bash
# ncu full captures of selected launches of one decode.  Usage: gpu_prof3.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
python tools/profile_step.py --iters 1 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
for spec in "conv_umma 14 conv_s0k11" "conv_umma 59 conv_s3k3c1" "conv_umma 97 conv_s5k3c1" "conv_umma 98 conv_s5k3c2" "act1d 90 act_s5"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_${3}_$TAG \
      python tools/profile_step.py --iters 1 > gpurun_out/ncu_${3}_$TAG.log 2>&1
  echo "$3 rc=$?"
done
