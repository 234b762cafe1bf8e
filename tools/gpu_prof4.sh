#!/bin/bash
# ncu full capture of one conv launch.  Usage: gpu_prof4.sh <tag> <skip> [name]
TAG=${1:-x}; SKIP=${2:-109}; NAME=${3:-conv_s5k11c1}
mkdir -p gpurun_out
export BVG_FUSE_ACT=${BVG_FUSE_ACT:-0}
python tools/profile_step.py --iters 1 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s $SKIP -c 1 -o gpurun_out/prof_${NAME}_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_${NAME}_$TAG.log 2>&1
echo "$NAME rc=$?"
