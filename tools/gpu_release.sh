#!/bin/bash
# Release-candidate pass: full -m gpu suite, smoke, bench (both arms), ncu launch list + full captures.
TAG=${1:-rc}
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -m gpu -p no:cacheprovider -s tests > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$? $(grep -E 'passed|failed' gpurun_out/tests_$TAG.log | tail -1)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
python tools/profile_step.py --iters 2 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/profile_step.py --iters 2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/profile_step.py --iters 1 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 4 -c 1 -o gpurun_out/prof_conv_s0k11_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_full_conv_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:act1d_c8 -s 30 -c 1 -o gpurun_out/prof_act_s5_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_full_act_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 99 -c 1 -o gpurun_out/prof_conv_s5k11_$TAG \
    python tools/profile_step.py --iters 1 > gpurun_out/ncu_full_conv5_$TAG.log 2>&1
echo "ncu full rc=$?"
# (lockstep launch order: conv launch 4 = stage 0, k = 11, m = 0 first convolution; 99 = the same of stage 5; activation launch 30 =
#  the first grouped launch of stage 5, three blocks in one grid)
timeout 300 python bench.py --precision fp32tc --steps 5 --warmup 3 --no-cpu-baseline --no-srt > gpurun_out/bench_fp32tc_$TAG.json 2> gpurun_out/bench_fp32tc_$TAG.err; echo "bench fp32tc rc=$?"
head -c 1500 gpurun_out/bench_$TAG.json
