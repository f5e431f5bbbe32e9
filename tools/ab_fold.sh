#!/bin/bash
# A/B over ab/lib_*.so: headline bench (device-timed step and e2e), two rounds
for round in 1 2; do
  for f in ab/lib_*.so; do
    cp "$f" mujoco-template_b200/libb2mj.so
    echo -n "$(basename $f) r$round: "
    python bench.py --no-cpu-baseline --no-secondary --no-e2e --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cartpole %.4g ms %.4f lin_us %.2f launches %s' % (d['value'], d['ms_per_step'], r['kernel_ms']*1e3, d['gpu_launches']))"
  done
done
