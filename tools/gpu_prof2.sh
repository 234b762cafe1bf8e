#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
python tools/profile_step.py --iters 2 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/profile_step.py --iters 2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/profile_step.py --iters 1 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 3 -c 1 -o gpurun_out/prof_conv_s0_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 62 -c 1 -o gpurun_out/prof_conv_s3_$TAG \
    python tools/profile_step.py --iters 1 >> gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 100 -c 1 -o gpurun_out/prof_conv_s5_$TAG \
    python tools/profile_step.py --iters 1 >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
