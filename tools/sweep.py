"""BASELINE config #5: scaling sweep N x {pendulum, cartpole, drone, humanoid} x {FP64, FP32} on one GPU.

Step-only throughput (zero/hover control, one mj_step per launch, CUDA events, L2 flushed between
launches) and FP64 linearizations/s.  Prints one JSON object per line; the committed output lives in
profiles/sweep_r04.jsonl.   python tools/sweep.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from conftest import load_model, random_states
from mujoco_template import _mj as mj

quick = "--quick" in sys.argv
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


sizes = {"pendulum": [2**10, 2**12, 2**14, 2**16, 2**18, 2**20, 2**22], "cartpole": [2**10, 2**12, 2**14, 2**16, 2**18, 2**20, 2**22],
         "drone": [2**10, 2**14, 2**18, 2**20, 2**22], "humanoid": [2**10, 2**12, 2**14, 2**16, 2**18, 2**20, 2**22]}
if quick:
    sizes = {k: v[:2] for k, v in sizes.items()}
for name, ns in sizes.items():
    model = load_model(name)
    for n in ns:
        qpos, qvel, ctrl = random_states(model, name, min(n, 4096), seed=5)
        reps = int(np.ceil(n / qpos.shape[0]))
        qpos, qvel = np.tile(qpos, (reps, 1))[:n], np.tile(qvel, (reps, 1))[:n]
        for prec in (64, 32):
            d = mj.BatchData(model, n, precision=prec)
            dt = d.qpos.dtype
            d.qpos.copy_(torch.as_tensor(qpos.T.copy(), device="cuda").to(dt)); d.qvel.copy_(torch.as_tensor(qvel.T.copy(), device="cuda").to(dt))
            if name == "drone":
                d.ctrl.fill_(3.2495625)
            for _ in range(3):
                mj.mj_step(model, d)
            ms = timed(lambda: d.backend.step(1, derived=False), 7 if name != "humanoid" else (3 if n <= 2**16 else 1))
            out = dict(model=name, nenv=n, precision=prec, kernel=d.backend.batch.kernel_variant, step_ms=ms,
                       env_steps_per_sec=n / ms * 1e3, bytes_per_step=(2 * model.nq + 2 * model.nv + model.nu + 2 * model.nv) * prec // 8)
            out["hbm_gbs"] = out["bytes_per_step"] * n / ms / 1e6
            if prec == 64 and n <= 2**18 and not (name == "humanoid" and n > 2**14):
                A, B = d.backend.linearize(1e-6, True)
                lms = timed(lambda: d.backend.linearize(1e-6, True, out=(A, B)), 3)
                out["linearize_ms"] = lms; out["linearizations_per_sec"] = n / lms * 1e3
                del A, B
            if name == "humanoid":
                mj.mj_step(model, d)
                out["mean_ncon"] = float(d.ncon.float().mean()); out["mean_newton_iter"] = float(d.solver_iter.float().mean())
            out["flags"] = int((d.flags != 0).sum())
            print(json.dumps(out), flush=True)
            del d
            torch.cuda.empty_cache()
