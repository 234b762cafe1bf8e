#!/bin/bash
# fused-activation validation: tests with fusion on/off must agree; then bench both.
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/fuse_tests.log 2>&1; echo "fused tests rc=$? $(grep -E 'passed|failed' gpurun_out/fuse_tests.log | tail -1)"
grep -E "SNR|max-abs|FAILED|Error" gpurun_out/fuse_tests.log | head -12
timeout 300 python tools/fuse_compare.py > gpurun_out/fuse_compare.log 2>&1; echo "compare rc=$?"; tail -6 gpurun_out/fuse_compare.log
bash tools/gpu_sweep2.sh fuse "BVG_FUSE_ACT=1" "BVG_FUSE_ACT=0" 2>&1 | grep -E "===|stage [0-9]|pre|steps|FAILED"
