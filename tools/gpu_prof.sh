#!/bin/bash
# ncu pass: launch list of one decode + full capture of the two hot kernels.  Usage: gpu_prof.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python tools/profile_step.py --iters 2 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/profile_step.py --iters 2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/profile_step.py --iters 1 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_umma|act1d" -s 40 -c 6 -o gpurun_out/prof_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -8
