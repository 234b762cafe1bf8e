#!/bin/bash
# Final pass of a round: release pass + per-stage sweep + side measurements, all from one box.
TAG=${1:-final}
bash tools/gpu_release.sh $TAG
bash tools/gpu_sweep2.sh $TAG "BVG_X=0" > gpurun_out/stage_sweep_$TAG.txt 2>&1; tail -9 gpurun_out/stage_sweep_$TAG.txt
bash tools/gpu_extras.sh $TAG > gpurun_out/extras_$TAG.log 2>&1; tail -4 gpurun_out/extras_$TAG.log
for f in 24 118; do python tools/small_step.py --frames $f; done > gpurun_out/small_$TAG.txt 2>&1; cat gpurun_out/small_$TAG.txt
