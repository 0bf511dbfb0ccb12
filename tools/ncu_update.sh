#!/bin/bash
# per-kernel times of one DRQN update and one DQN update (ncu launch list): tools/ncu_update.sh TAG
TAG=$1
cat > /tmp/upd.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer, SequenceSampler
n, T = 4096, 64
ring = pp.ReplayRing(n * T, lockstep_envs=n)
ring.obs.uniform_(-1, 1); ring.next_obs.uniform_(-1, 1)
ring.done.copy_((torch.rand(n * T, device="cuda") < 0.03).to(torch.uint8)); ring.steps_written = T; ring.head.fill_(n * T)
s = SequenceSampler(ring, trace_length=8); s.refresh()
torch.manual_seed(0)
tr = DRQNTrainer(pp.QNetRNN(), batch_size=64, use_graph=False)
for _ in range(3): tr.update(s)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_update_$TAG.csv python /tmp/upd.py > gpurun_out/ncu_update_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_update_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
# last third = third update
names=[(r[4], float(r[-1])) for r in rows]
k=len(names)//3
tot=0
for nme,t in names[-k:]:
    print(f"{t/1000:8.2f} us  {nme[:90]}"); tot+=t
print("sum %.1f us over %d launches" % (tot/1000, k))
PY
