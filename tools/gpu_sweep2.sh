#!/bin/bash
# bench sweep with per-stage breakdown.  Usage: gpu_sweep2.sh <tag> "<ENV=..>" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  out=gpurun_out/bench_${TAG}_$i.log
  env $cfg BVG_PROF_DUMP=gpurun_out/dump_${TAG}_$i.txt timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-srt > $out 2>&1
  echo "=== [$cfg] $(python -c "
import json,sys
try:
    d=json.loads(open('$out').read().strip().splitlines()[-1]); print('value %.0f ms/step %.2f'%(d['value'], d['ms_per_step']))
except Exception as e: print('FAILED', open('$out').read()[-400:])
")"
  python tools/stage_summary.py gpurun_out/dump_${TAG}_$i.txt 2>&1 | tail -8
  i=$((i+1))
done
