#!/bin/bash
# run on the GPU box: alternate A and B three times each, print kernel ms
cd "$(dirname "$0")/.."
B=$PWD/distantspeechrecognition-mirror_b200/libbtkb200_b.so
for i in 1 2 3; do
  for v in A B; do
    if [ $v = B ]; then export BTKB200_LIB=$B; else unset BTKB200_LIB; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload ${1:-cfg2} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"
  done
done
