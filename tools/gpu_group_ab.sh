#!/bin/bash
# A/B of the lockstep AMP blocks / grouped Activation1d launches (BVG_ACT_GROUP) in ONE call: same box, same clocks.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -3
for rep in 1 2; do
for v in 0 1; do
  BVG_ACT_GROUP=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-srt > gpurun_out/r2g_group$v.json 2> gpurun_out/r2g_group$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_group$v.json").read().strip().splitlines()[-1])
print("group=$v value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "conv ms", round(d["roofline"]["ms_per_step"],3), "act ms", round(d.get("roofline_activation",{}).get("ms_per_step"),3), "launches", d.get("gpu_launches"))
PY
done
done
