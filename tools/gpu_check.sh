#!/bin/bash
# Regular GPU pass: all -m gpu tests, then bench (bf16) with optional env knobs.  Usage: gpu_check.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -m gpu -p no:cacheprovider -s tests > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/tests_$TAG.log | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
          "conv TF/s", round(d["roofline"]["achieved"]), "act GB/s", round(d["roofline_activation"]["achieved"]),
          d["kernel_classes_ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
