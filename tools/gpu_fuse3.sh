#!/bin/bash
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_ops.py -k "act_conv" > gpurun_out/fuse3_ops.log 2>&1; echo "ops rc=$? $(grep -E 'passed|failed' gpurun_out/fuse3_ops.log | tail -1)"
BVG_FUSE_ACT=1 timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/fuse3_fwd.log 2>&1; echo "fwd(fused) rc=$? $(grep -E 'passed|failed' gpurun_out/fuse3_fwd.log | tail -1)"
grep -E "SNR|FAILED|Error" gpurun_out/fuse3_fwd.log | head -8
bash tools/gpu_sweep2.sh fuse3 "BVG_FUSE_ACT=1" 2>&1 | grep -E "===|stage [0-9]|pre|steps|FAILED"
bash tools/gpu_trace.sh "BVG_FUSE_ACT=1 BVG_CONV_TRACE=24 BVG_CONV_TRACE_TAPS=3" "BVG_FUSE_ACT=1 BVG_CONV_TRACE=96 BVG_CONV_TRACE_TAPS=11"
