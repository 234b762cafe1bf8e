#!/bin/bash
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
export BVG_ACT_MMA=1
timeout 600 $PYT tests/test_gpu_ops.py -k "packed_kernel" 2>&1 | tail -15
timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/mma_fwd.log 2>&1; echo "fwd rc=$? $(grep -E 'passed|failed' gpurun_out/mma_fwd.log | tail -1)"
grep -E "SNR|FAILED|Error" gpurun_out/mma_fwd.log | head -8
bash tools/gpu_sweep2.sh mma "BVG_ACT_MMA=1" "BVG_ACT_MMA=0" 2>&1 | grep -E "===|stage [0-9]|steps|FAILED"
