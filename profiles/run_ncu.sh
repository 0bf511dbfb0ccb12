#!/bin/bash
# ncu evidence for bench.py (run under gpurun, one GPU).  Usage: bash profiles/run_ncu.sh <round-tag>
# 1) plain run must exit 0;  2) launch list with device time per launch;  3) --set full on the two kernels that matter.
set -u
TAG=${1:-r01}
EXTRA=${2:-}            # e.g. "--precision f16"
OUT=gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --k1-envs 4194304 $EXTRA"
$CMD > $OUT/ncu_plain_${TAG}.json 2> $OUT/ncu_plain_${TAG}.err || { echo "plain run failed"; tail -5 $OUT/ncu_plain_${TAG}.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:selfplay -s 3 -c 2 -f -o $OUT/prof_selfplay_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-k1 $EXTRA > $OUT/ncu_selfplay_${TAG}.log 2>&1
echo "selfplay full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -f -o $OUT/prof_k1_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --k1-envs 4194304 $EXTRA > $OUT/ncu_k1_${TAG}.log 2>&1
echo "k1 full rc=$?"
ls -la $OUT
