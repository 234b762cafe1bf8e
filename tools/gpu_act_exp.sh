#!/bin/bash
# Builds tools/act_exp.cu once per experiment mask and times the stage-5 / stage-3 / stage-0 activation launches of cfg2.
mkdir -p gpurun_out
for m in ${@:-0 1 2 4 8 16 3 12 31}; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DBVG_ACT_EXP=$m -o /tmp/act_exp_$m tools/act_exp.cu || exit 1
  /tmp/act_exp_$m 24 240640; /tmp/act_exp_$m 96 60160; /tmp/act_exp_$m 768 940
done 2>&1 | tee gpurun_out/act_exp.log
