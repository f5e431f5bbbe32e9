#!/bin/bash
# PTX zero/one folding of the specialised kernels: GPU suite on the folded build, then bench A/B against ab/lib_nofold.so
cp mujoco-template_b200/libb2mj.so ab/lib_fold.so
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for round in 1 2; do
  for f in ab/lib_nofold.so ab/lib_fold*.so; do
    cp "$f" mujoco-template_b200/libb2mj.so
    echo -n "$(basename $f) r$round: "
    python bench.py --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cartpole %.4g ms %.4f lin_us %.2f | drone %.4g | tv %.4g | fp64 frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms']*1e3, d['secondary']['drone']['value'], d['secondary']['cartpole_tv_lqr']['value'], d['roofline_fp64']['frac']))"
  done
done
cp ab/lib_fold.so mujoco-template_b200/libb2mj.so
