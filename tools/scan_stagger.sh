#!/bin/bash
# scan the phase offset between the groups of a CTA in the fused tensor-core kernel
for s in 0 800 1600 2400 3200 4000 5000 6500; do
  v=$(PP_TC_STAGGER=$s python bench.py --steps 10 --warmup 3 --lockstep 4096 --no-cpu-baseline --no-e2e --no-k1 --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])")
  echo "stagger $s: $v"
done
