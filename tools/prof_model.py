"""Run a few steps (and optionally FD linearisations) of one model so that ncu can capture the kernels.
  python tools/prof_model.py <model> <nenv> [--lin] [--presteps K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from conftest import load_model, random_states
from mujoco_template import _mj as mj

name, n = sys.argv[1], int(sys.argv[2])
model = load_model(name)
d = mj.BatchData(model, n)
qpos, qvel, ctrl = random_states(model, name, min(n, 8192), seed=0)
reps = -(-n // qpos.shape[0])
up = lambda a: torch.as_tensor(a.T.copy(), device="cuda").repeat(1, reps)[:, :n]
d.qpos.copy_(up(qpos)); d.qvel.copy_(up(qvel)); d.ctrl.copy_(up(ctrl))
pre = int(sys.argv[sys.argv.index("--presteps") + 1]) if "--presteps" in sys.argv else 0
lo, hi = {"drone": (0.0, 13.0), "humanoid": (-0.2, 0.2)}.get(name, (-1.0, 1.0))
for _ in range(pre):  # age the batch under random controls (drones reach the floor, the humanoid falls)
    d.ctrl.uniform_(lo, hi)
    d.backend.step(1, derived=False)
for _ in range(3):
    d.backend.step(1, derived=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    d.backend.step(1, derived=False)
e1.record()
torch.cuda.synchronize()
print("step_ms", e0.elapsed_time(e1) / 4, "ncon", float(d.ncon.float().mean()) if hasattr(d, "ncon") else None)
if "--lin" in sys.argv:
    A, B = d.backend.linearize(1e-6, True)
    d.backend.linearize(1e-6, True, out=(A, B))
torch.cuda.synchronize()
print("ok", d.backend.batch.kernel_variant, "bad flags", int((d.flags != 0).sum()))
