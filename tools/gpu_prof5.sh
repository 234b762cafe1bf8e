#!/bin/bash
# ncu full capture of one launch of a kernel.  Usage: gpu_prof5.sh <tag> <kernel-regex> <skip> <name>   (env passes through)
TAG=$1; KR=$2; SKIP=$3; NAME=$4
mkdir -p gpurun_out
python tools/profile_step.py --iters 1 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c 1 -o gpurun_out/prof_${NAME}_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_${NAME}_$TAG.log 2>&1
echo "$NAME rc=$?"
