#!/bin/bash
# First GPU pass: op tests (CUDA-core), fp32 forward parity, tcgen05 probe + tests, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_ops.py -k "not tcgen05" > gpurun_out/t_ops_cc.log 2>&1; echo "ops_cc rc=$?"
timeout 300 python tools/umma_probe.py > gpurun_out/umma_probe.log 2>&1; PROBE=$?; echo "probe rc=$PROBE"
timeout 1500 $PYT -s tests/test_gpu_forward.py -k "fp32 and not bf16 or dropin or speaker or oracle or checkpoint" > gpurun_out/t_fwd_fp32.log 2>&1; echo "fwd_fp32 rc=$?"
if [ "$PROBE" = "0" ]; then
  timeout 900 $PYT tests/test_gpu_ops.py -k "tcgen05" > gpurun_out/t_ops_tc.log 2>&1; echo "ops_tc rc=$?"
  timeout 1500 $PYT -s tests/test_gpu_forward.py -k "bf16" > gpurun_out/t_fwd_bf16.log 2>&1; echo "fwd_bf16 rc=$?"
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
fi
timeout 900 python bench.py --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 rc=$?"
tail -n 5 gpurun_out/*.log
