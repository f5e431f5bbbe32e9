#!/bin/bash
# The three configs at their full length of T = 1000 steps (device-timed), one line each
for m in humanoid drone cartpole; do
  lin=--no-linearize; [ $m = cartpole ] && lin=
  python bench.py --model $m $lin --steps 1000 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$m', d['steps'], '%.4g' % d['value'], d['ms_per_step'], d.get('contact_stats'), d.get('bad_env_flags'))"
done
