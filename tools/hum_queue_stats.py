"""Humanoid config #4 rollout: share of envs with branch-coupling (dense) constraint rows and Newton rounds over time
(the warp engine's queue keys, b2_warp_queue_histogram).   python tools/hum_queue_stats.py [nenv]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200"))
import numpy as np
import torch

import bench
from mujoco_template import _mj as mj

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = bench.load_model("humanoid")
qpos, qvel = bench.synth_states(model, "humanoid", n, 0)
d = mj.BatchData(model, n)
d.qpos.copy_(torch.as_tensor(qpos.T.copy(), device="cuda")); d.qvel.copy_(torch.as_tensor(qvel.T.copy(), device="cuda"))
for s in range(1, 401):
    d.ctrl.uniform_(-0.2, 0.2)
    d.backend.step(1, derived=False)
    if s % 25 == 0:
        d.backend.step(1, derived=False)  # the histogram is of the keys written by the step before
        h = np.array(d.backend.batch.warp_queue_histogram())
        dense = h[8:].sum() / h.sum()
        rounds = (np.arange(16) % 8 * h).sum() / h.sum()
        print(f"step {s:4d}: dense-row envs {100 * dense:5.1f} %  mean Newton rounds {rounds:.2f}  hist {h.tolist()}", flush=True)
