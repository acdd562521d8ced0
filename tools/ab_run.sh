#!/bin/bash
# run on the GPU box: tools/ab_run.sh <workload> <name>...  alternates the default build ("A") with the named builds
cd "$(dirname "$0")/.."
wl=${1:-cfg2}; shift
for i in 1 2; do
  for v in A "$@"; do
    if [ $v = A ]; then unset BTKB200_LIB; else export BTKB200_LIB=$PWD/distantspeechrecognition-mirror_b200/libbtkb200_$v.so; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"
  done
done
