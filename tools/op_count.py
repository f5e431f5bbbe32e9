"""Algorithmic operation counts of one mj_step / one mjd_transitionFD + step per model, from the oracle's op-counting build
(oracle/orc_count.h: the same C source compiled with a counting arithmetic type).
States: bench.py's initial-state distributions (humanoid: after 20 warm-up steps, so that contacts and qacc_warmstart are
realistic).  Writes profiles/op_count_r02.json, which bench.py quotes beside the ncu-counted EXECUTED flops of the kernels.

    python tools/op_count.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import warnings

warnings.simplefilter("ignore")
import numpy as np

import bench
from oracle.oracle import op_count

out = {"weights": "add / sub / mul / div / sqrt = 1 flop, sin / cos / atan2 / ... call = 20 (SURVEY.md section 8d)", "models": {}}
for name in ("pendulum", "cartpole", "drone", "humanoid"):
    m = bench.load_model(name)
    dims = dict(nq=m.nq, nv=m.nv, nu=m.nu, nbody=m.nbody, njnt=m.njnt, ngeom=m.ngeom, nsite=m.nsite, ntendon=m.ntendon)
    n = 16
    qpos, qvel = bench.synth_states(m, name, n, 0)
    rng = np.random.default_rng(1)
    lo, hi = bench.RANDOM_CTRL.get(name, (0.0, 0.0))
    ctrl = rng.uniform(lo, hi, (n, m.nu))
    warm = 20 if name == "humanoid" else 1
    steps = [op_count(m.blob, dims, qpos[i], qvel[i], ctrl[i], warm_steps=warm) for i in range(n)]
    lin = [op_count(m.blob, dims, qpos[i], qvel[i], ctrl[i], linearize=True, warm_steps=warm) for i in range(4 if name == "humanoid" else n)]
    mean = lambda rows, k: float(np.mean([r[k] for r in rows]))
    out["models"][name] = {
        "step": {k: mean(steps, k) for k in ("flops", "add", "mul", "div", "sqrt", "transcendental", "ncon", "nefc", "solver_iter")},
        "linearize_plus_step": {"flops": mean(lin, "flops")},
        "rollouts_per_linearization": 1 + 2 * (2 * m.nv + m.nu),
        "samples": n,
    }
    print(name, json.dumps(out["models"][name]))
json.dump(out, open(os.path.join(ROOT, "profiles", "op_count_r02.json"), "w"), indent=1)
