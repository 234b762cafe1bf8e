#!/bin/bash
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
for v in 1 0; do
  echo "== UPLO=$v"; BVG_ACT_MMA_UPLO=$v timeout 600 $PYT -s tests/test_gpu_forward.py -k "cfg1_bf16 or cfg2" 2>&1 | grep -E "SNR|passed|failed"
done
bash tools/gpu_sweep2.sh uplo "BVG_ACT_MMA_UPLO=1" "BVG_ACT_MMA_UPLO=0" 2>&1 | grep -E "===|stage [0-5]|steps" | sed 's/|.*| act/ act/'
bash tools/gpu_prof5.sh mma2 act1d_c8_mma 95 act_mma_s5
