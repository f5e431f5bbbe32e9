# A/B over ab/lib_*.so x warps per block: humanoid config, 200- and 600-step rollouts
for w in ${WPBS:-4}; do for f in ab/lib_*.so; do cp "$f" mujoco-template_b200/libb2mj.so; for st in 200 600; do echo -n "$(basename $f) wpb=$w steps=$st: "; B2_WARP_LS_WPB=$w python tools/bench_value.py --model humanoid --no-linearize --steps $st 2>&1 | cut -c1-60; done; done; done
