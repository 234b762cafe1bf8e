#!/bin/bash
# ncu --set full captures of the fp32 tensor-core mode: the stage-0 k=11 convolution (F32IO kernel) and a stage-5 activation
# (fp32 kernel, split output).  Sequential launch order in this mode: conv launch 14, activation launch 90.
mkdir -p gpurun_out
python tools/profile_step.py --precision fp32tc --iters 1 > gpurun_out/plain_tc32.log 2>&1 || { tail -5 gpurun_out/plain_tc32.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 14 -c 1 -o gpurun_out/prof_tc32_conv_s0k11 \
    python tools/profile_step.py --precision fp32tc --iters 1 > gpurun_out/ncu_tc32_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:act1d_c8_v3 -s 90 -c 1 -o gpurun_out/prof_tc32_act_s5 \
    python tools/profile_step.py --precision fp32tc --iters 1 > gpurun_out/ncu_tc32_act.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 109 -c 1 -o gpurun_out/prof_tc32_conv_s5k11 \
    python tools/profile_step.py --precision fp32tc --iters 1 > gpurun_out/ncu_tc32_conv5.log 2>&1
ls -la gpurun_out/prof_tc32_*.ncu-rep
