"""Humanoid step throughput against resident warps per SM (B2_WARP_EXTRA_SMEM pads the block's shared memory).
  python tools/hum_sweep.py [nenv] [extra_smem ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from conftest import load_model, random_states
from mujoco_template import _mj as mj

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
extras = [int(x) for x in sys.argv[2:]] or [0]
model = load_model("humanoid")
qpos, qvel, ctrl = random_states(model, "humanoid", min(n, 8192), seed=0)
reps = -(-n // qpos.shape[0])
up = lambda a: torch.as_tensor(a.T.copy(), device="cuda").repeat(1, reps)[:, :n]
for extra in extras:
    os.environ["B2_WARP_EXTRA_SMEM"] = str(extra)
    d = mj.BatchData(model, n)
    d.qpos.copy_(up(qpos)); d.qvel.copy_(up(qvel)); d.ctrl.copy_(up(ctrl))
    torch.manual_seed(0)
    for _ in range(30):
        d.ctrl.uniform_(-0.2, 0.2)
        d.backend.step(1, derived=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.backend.step(1, derived=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    d.backend.step(1, derived=True)
    print(f"extra_smem={extra} step_ms={ms:.4f} env_steps_per_s={n / ms * 1e3:.4g} ncon={float(d.ncon.float().mean()):.2f} "
          f"nefc={float(d.nefc.float().mean()):.2f} iters={float(d.solver_iter.float().mean()):.2f} bad={int((d.flags != 0).sum())}", flush=True)
    del d
