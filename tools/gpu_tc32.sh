#!/bin/bash
# fp32 tensor-core mode: accuracy (forward tests) and speed per activation variant (BVG_TC32_ACT = 1 sinf, 2 polynomial, 0 MUFU)
mkdir -p gpurun_out
for v in 2 1 0; do
  echo "== BVG_TC32_ACT=$v"
  BVG_TC32_ACT=$v timeout 300 python -m pytest tests/test_gpu_forward.py -x -q -k "fp32tc_tiny or fp32tc_cfg2" -s 2>&1 | grep -i "max-abs\|passed\|failed"
  BVG_TC32_ACT=$v timeout 200 python bench.py --precision fp32tc --steps 5 --warmup 3 --no-cpu-baseline --no-srt > gpurun_out/r2g_fp32tc_act$v.json 2> gpurun_out/r2g_fp32tc_act$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_fp32tc_act$v.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "conv ms", d["roofline"]["ms_per_step"], "frac", d["roofline"]["frac"], "act ms", d.get("roofline_activation",{}).get("ms_per_step"))
PY
done
