#!/bin/bash
# A/B of the independent-successor launches (BVG_PDL_INDEP) in ONE call
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -k "lockstep or cfg2 or ragged or graph or handoff or cfg5" 2>&1 | tail -3
for rep in 1 2 3; do
for v in 0 1; do
  BVG_PDL_INDEP=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-srt > gpurun_out/r2g_indep$v.json 2> gpurun_out/r2g_indep$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_indep$v.json").read().strip().splitlines()[-1])
print("indep=$v value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "serial", round(d["e2e"]["serial_one_stream"]["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"])
PY
done
done
for v in 0 1; do for f in 24 118; do BVG_PDL_INDEP=$v python tools/small_step.py --frames $f 2>&1 | tail -1; done; done
