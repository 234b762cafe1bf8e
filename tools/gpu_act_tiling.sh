#!/bin/bash
# Sweep the activation kernel's tile length / tiles per warp on the stage 0-3 shapes of cfg2 (tools/act_exp.cu).
mkdir -p gpurun_out
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DBVG_ACT_EXP=0 -o /tmp/act_exp_0 tools/act_exp.cu || exit 1
for shape in "768 940" "384 3760" "192 15040" "96 60160"; do
  for gt in 1 2; do for tw in 64 80 96 112 128 144 160 176 192 208 224 240; do
    echo -n "GT=$gt TW=$tw  "; BVG_ACT_GT=$gt BVG_ACT_TW=$tw /tmp/act_exp_0 $shape
  done; done
  echo -n "auto        "; /tmp/act_exp_0 $shape
done 2>&1 | tee gpurun_out/act_tiling.log
