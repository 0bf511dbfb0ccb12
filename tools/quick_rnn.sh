#!/bin/bash
# quick recurrent-rollout numbers: tools/quick_rnn.sh TAG
for envs in 32768 262144; do
timeout 200 python bench.py --workload rnn --envs $envs --steps 8 --warmup 3 --lockstep 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 rnn envs $envs: %.4e env-steps/s  %.3f ms/launch' % (d['value'], d['ms_per_step']))"
done
