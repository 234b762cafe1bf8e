#!/bin/bash
# pipeline traces of one conv launch.  Usage: gpu_trace.sh "<ENV...>" ...
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-srt > gpurun_out/trace_bench.log 2>&1
  python tools/trace_summary.py
done
