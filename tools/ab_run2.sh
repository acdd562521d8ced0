#!/bin/bash
# run on the GPU box: tools/ab_run2.sh <workload> <name>...  alternates the named builds (no default build)
cd "$(dirname "$0")/.."
wl=${1:-cfg2}; shift
for i in 1 2; do
  for v in "$@"; do
    export BTKB200_LIB=$PWD/distantspeechrecognition-mirror_b200/libbtkb200_$v.so
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v', '$wl', round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))
except Exception as e: print('$v failed', e)"
  done
done
