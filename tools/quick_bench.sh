#!/bin/bash
# quick headline + e2e numbers (no CPU baseline, no secondary configs): tools/quick_bench.sh TAG [extra bench args]
TAG=$1; shift
timeout 300 python bench.py --steps 10 --warmup 3 --lockstep 4096 --no-cpu-baseline --no-k1 --no-secondary --e2e-calls 20 "$@" > gpurun_out/quick_$TAG.json 2> gpurun_out/quick_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/quick_$TAG.json').read().strip().splitlines()[-1])
print("$TAG value %.4e  ms/step %.3f  e2e %.4e  ms/call %.3f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call']))
PY
