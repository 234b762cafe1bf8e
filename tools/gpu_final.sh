#!/bin/bash
# full -m gpu suite with default settings and with the MMA activation forced for every length, then a bench
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 --timeout-method=thread"
timeout 1200 $PYT -s tests > gpurun_out/final_tests.log 2>&1; echo "tests(default) rc=$? $(grep -E 'passed|failed' gpurun_out/final_tests.log | tail -1)"
grep -E "SNR dB" gpurun_out/final_tests.log | head -3
BVG_ACT_MMA_MINLEN=1 timeout 1200 $PYT tests > gpurun_out/final_tests_mma.log 2>&1; echo "tests(mma everywhere) rc=$? $(grep -E 'passed|failed' gpurun_out/final_tests_mma.log | tail -1)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_final.json
