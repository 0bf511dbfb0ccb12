#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU), after the same commands exited 0 without ncu.
#   bash profiles/run_ncu_r02.sh
set -u
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --lockstep 256 --no-cpu-baseline --k1-envs 4194304 --no-secondary --e2e-calls 5"
timeout 300 $CMD > $OUT/ncu_plain_r02.json 2> $OUT/ncu_plain_r02.err || { echo "plain run failed"; tail -5 $OUT/ncu_plain_r02.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r02.csv $CMD > $OUT/ncu_launches_r02.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selfplay_tc -s 3 -c 1 -f -o $OUT/prof_selfplay_tc_r02 \
    python bench.py --steps 2 --warmup 3 --lockstep 256 --no-cpu-baseline --no-e2e --no-k1 --no-secondary > $OUT/ncu_selfplay_r02.log 2>&1
echo "selfplay_tc full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selfplay_rnn_tc -s 3 -c 1 -f -o $OUT/prof_rnn_tc_r02 \
    python bench.py --workload rnn --envs 32768 --steps 2 --warmup 3 --lockstep 8 --no-cpu-baseline > $OUT/ncu_rnn_r02.log 2>&1
echo "selfplay_rnn_tc full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_bwd|lstm_fwd|sgemm" -s 40 -c 8 -f -o $OUT/prof_drqn_r02 \
    python bench.py --workload train_rnn --envs 4096 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_drqn_r02.log 2>&1
echo "drqn kernels full rc=$?"
ls -la $OUT | grep r02
