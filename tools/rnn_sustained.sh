for pair in 1 0; do
PP_RNN_PAIR=$pair timeout 200 python bench.py --workload rnn --envs 262144 --steps 40 --warmup 5 --lockstep 32 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair=$pair sustained rnn: %.4e env-steps/s  %.3f ms/launch' % (d['value'], d['ms_per_step']), d.get('clocks'))"
done
