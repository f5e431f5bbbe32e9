"""Throughput of b2_dlqr (batched DARE + LQR gain): cartpole (nx 4) and drone (nx 12) batches straight from b2_linearize.
  python tools/dlqr_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200"))
import numpy as np
import torch

import bench
from mujoco_template import _mj as mj
from mujoco_template.batched_controllers import batched_dlqr_gain

for name, n in (("cartpole", 65536), ("drone", 65536), ("humanoid", 256)):
    model = bench.load_model(name)
    qpos, qvel = bench.synth_states(model, name, n, 0)
    d = mj.BatchData(model, n)
    d.qpos.copy_(torch.as_tensor(qpos.T.copy(), device="cuda")); d.qvel.copy_(torch.as_tensor(qvel.T.copy(), device="cuda"))
    if name == "drone":
        d.ctrl.fill_(3.2495625)
    A, B = d.backend.linearize(1e-6, True)        # (nx, nx, N), (nx, nu, N)
    nx, nu = A.shape[0], B.shape[1]
    Q, R = np.eye(nx), np.eye(nu)
    An, Bn = A.permute(2, 0, 1), B.permute(2, 0, 1)
    K, P, st = batched_dlqr_gain(An, Bn, Q, R, return_status=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        batched_dlqr_gain(An, Bn, Q, R)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rho = float(torch.linalg.eigvals(An[:64] - Bn[:64] @ K[:64]).abs().amax())
    print(f"{name}: nx={nx} nu={nu} N={n}: {ms:.3f} ms per batch = {n / ms * 1e3:.3g} gains/s; doublings {int(st.min())}..{int(st.max())}, "
          f"failed {int((st < 0).sum())}, closed-loop spectral radius (first 64 envs) {rho:.6f}", flush=True)
