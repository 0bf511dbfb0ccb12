#!/bin/bash
# One GPU visit: the parity suite, the bench line, and (optionally) sanitizer passes.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh TAG [tests] [bench] [san] [learn]
TAG=$1; shift
mkdir -p gpurun_out
for what in "$@"; do case $what in
  tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log;;
  bench) timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"; tail -c 600 gpurun_out/bench_$TAG.json;;
  learn) timeout 600 python tools/learning_curve.py > gpurun_out/learn_$TAG.log 2>&1; cat gpurun_out/learn_$TAG.log;;
  san) for tool in memcheck racecheck synccheck; do
         timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/san_${tool}_$TAG.log python tools/sanitize_launches.py > gpurun_out/san_${tool}_$TAG.out 2>&1
         echo "$tool rc $?"; tail -3 gpurun_out/san_${tool}_$TAG.log
       done;;
esac; done
