#!/bin/bash
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
export BVG_FUSE_ACT=${BVG_FUSE_ACT:-0}
timeout 900 $PYT tests/test_gpu_ops.py tests/test_gpu_forward.py > gpurun_out/lean_tests.log 2>&1; echo "tests rc=$? $(grep -E 'passed|failed' gpurun_out/lean_tests.log | tail -1)"
bash tools/gpu_sweep2.sh lean "BVG_FUSE_ACT=0" "BVG_FUSE_ACT=1" 2>&1 | grep -E "===|stage [0-9]|pre|steps|FAILED"
bash tools/gpu_trace.sh "BVG_FUSE_ACT=0 BVG_CONV_TRACE=24 BVG_CONV_TRACE_TAPS=11" "BVG_FUSE_ACT=0 BVG_CONV_TRACE=96 BVG_CONV_TRACE_TAPS=3"
