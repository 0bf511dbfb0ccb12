#!/bin/bash
# N-GPU check (default 2): the full bench line under torchrun, with hard timeouts.  tools/multi_gpu_check.sh TAG [N]
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 --lockstep 4096 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench N=$N rc $?"
tail -3 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}_n$N.json').read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e (ranks %s)" % (d['value'], d['e2e']['value'], d['e2e'].get('ranks')))
    for k in ('config4_rnn','config5_train'):
        c=d.get(k)
        if c: print(k, "%.4e" % c['value'], c.get('updates_per_s'), c.get('grad_allreduce'), c.get('grad_allreduce_us'))
except Exception as e:
    print("no line:", e)
PY
