#!/bin/bash
# Tests + bench sweep over tuning env knobs.  Usage: gpu_sweep.sh <tag> "<ENV=.. ENV=..>" ...
TAG=$1; shift
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_ops.py > gpurun_out/ops_$TAG.log 2>&1; echo "ops rc=$? $(grep -E 'passed|failed' gpurun_out/ops_$TAG.log | tail -1)"
timeout 900 $PYT -s tests/test_gpu_forward.py > gpurun_out/fwd_$TAG.log 2>&1; echo "fwd rc=$? $(grep -E 'passed|failed' gpurun_out/fwd_$TAG.log | tail -1)"
grep -E "SNR|max-abs" gpurun_out/fwd_$TAG.log | head -8
i=0
for cfg in "" "$@"; do
  out=gpurun_out/bench_${TAG}_$i.log
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out 2>&1
  python - "$out" "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernel_classes_ms_per_step"]
    print(f"[{sys.argv[2]}] value {d['value']:.0f} ms/step {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.0f} conv {k['conv_tcgen05']:.2f} ms ({d['roofline']['achieved']:.0f} TF/s) act {k['activation1d']:.2f} ms ({d['roofline_activation']['achieved']:.0f} GB/s) other {k['other']:.2f}")
except Exception as e:
    print(f"[{sys.argv[2]}] bench failed: {e}"); print(open(sys.argv[1]).read()[-600:])
PY
  i=$((i+1))
done
