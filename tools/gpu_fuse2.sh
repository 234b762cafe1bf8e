#!/bin/bash
# residual-via-identity-MMA + fused activation: op tests, forward tests, then per-stage sweep.
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
timeout 300 python tools/fuse_debug.py 2>&1 | grep case
timeout 900 $PYT tests/test_gpu_ops.py > gpurun_out/fuse2_ops.log 2>&1; echo "ops rc=$? $(grep -E 'passed|failed' gpurun_out/fuse2_ops.log | tail -1)"
timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/fuse2_fwd.log 2>&1; echo "fwd rc=$? $(grep -E 'passed|failed' gpurun_out/fuse2_fwd.log | tail -1)"
grep -E "SNR|FAILED|Error" gpurun_out/fuse2_fwd.log | head -8
BVG_FUSE_ACT=0 timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/fuse2_fwd0.log 2>&1; echo "fwd(unfused) rc=$? $(grep -E 'passed|failed' gpurun_out/fuse2_fwd0.log | tail -1)"
grep -E "SNR|FAILED|Error" gpurun_out/fuse2_fwd0.log | head -8
bash tools/gpu_sweep2.sh fuse2 "BVG_FUSE_ACT=1" "BVG_FUSE_ACT=0" "BVG_FUSE_ACT=0 BVG_RES_MMA_MAXC=0" "BVG_FUSE_ACT=1 BVG_RES_MMA_MAXC=96" 2>&1 | grep -E "===|stage [0-9]|pre|steps|FAILED"
