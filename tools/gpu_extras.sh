#!/bin/bash
# Round-2 side measurements: cfg5 sweep (bf16 + fp16), reference-kernel comparison, small-decode latency, ECAPA timing.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python tools/cfg5_sweep.py --precision bf16 > gpurun_out/cfg5_sweep_bf16_$TAG.md 2> gpurun_out/cfg5_sweep_$TAG.err; echo "cfg5 bf16 rc=$?"
timeout 900 python tools/cfg5_sweep.py --precision fp16 --steps 3 > gpurun_out/cfg5_sweep_fp16_$TAG.md 2>> gpurun_out/cfg5_sweep_$TAG.err; echo "cfg5 fp16 rc=$?"
timeout 600 python tools/ref_kernel_bench.py > gpurun_out/ref_kernel_bench_$TAG.md 2> gpurun_out/ref_kernel_bench_$TAG.err; echo "ref kernel rc=$?"
timeout 300 python tools/launch_bench.py > gpurun_out/launch_bench_$TAG.txt 2>&1; echo "launch bench rc=$?"
timeout 300 python tools/ecapa_bench.py > gpurun_out/ecapa_bench_$TAG.txt 2>&1; echo "ecapa rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp16 --no-srt > gpurun_out/bench_fp16_$TAG.json 2> /dev/null; echo "bench fp16 rc=$?"
tail -3 gpurun_out/ref_kernel_bench_$TAG.md; tail -8 gpurun_out/cfg5_sweep_bf16_$TAG.md
