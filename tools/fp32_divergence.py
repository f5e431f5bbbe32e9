"""FP32 mode: divergence of the FP32 trajectory from the FP64 one at 100 / 300 / 1000 steps, per model and regime, with the
exponential growth rate fitted between 100 and 1000 steps (printed as a table and as JSON lines; the numbers are pinned in
tests/test_gpu_parity.py::test_fp32_mode_divergence_bound_1000_steps and quoted in BASELINE.md).

    python tools/fp32_divergence.py [nenv]
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

CHECK = (100, 300, 1000)


def regimes(name, model, n):
    """(label, qpos, qvel, ctrl, presteps): presteps are run in FP64 first and both precisions start from that state."""
    qpos, qvel, ctrl = random_states(model, name, n, seed=21)
    ctrl[:] = 0
    if name == "pendulum":
        q = qpos.copy(); q[:, 0] = np.linspace(-2.0, 2.0, n)
        yield "passive swing, |theta0| <= 2 rad", q, np.zeros_like(qvel), ctrl, 0
        q = qpos.copy(); q[:, 0] = np.pi + np.linspace(-0.05, 0.05, n)
        yield "released within 0.05 rad of upright", q, np.zeros_like(qvel), ctrl, 0
    elif name == "cartpole":
        yield "passive, damped (pole falls and swings)", qpos, qvel, ctrl, 0
    elif name == "drone":
        c = ctrl.copy(); c[:] = 3.2495625
        yield "hover thrust held, small random tilt / rates (open loop)", qpos, qvel, c, 0
        c = np.random.default_rng(3).uniform(0, 13, ctrl.shape)
        yield "random constant thrust per rotor (tumbling)", qpos, qvel, c, 0
    elif name == "humanoid":
        yield "falling from stand_on_left_leg, zero control", qpos, qvel, ctrl, 0
        yield "at rest on the floor (state after 1500 FP64 steps)", qpos, qvel, ctrl, 1500


def run(model, prec, qpos, qvel, ctrl):
    n = qpos.shape[0]
    d = mj.BatchData(model, n, precision=prec)
    dt = d.qpos.dtype
    up = lambda a: torch.as_tensor(np.ascontiguousarray(a.T), device="cuda").to(dt)
    d.qpos.copy_(up(qpos)); d.qvel.copy_(up(qvel))
    if model.nu:
        d.ctrl.copy_(up(ctrl))
    out, done = {}, 0
    for k in CHECK:
        mj.mj_step(model, d, k - done); done = k
        out[k] = (d.qpos.double().cpu().numpy().T.copy(), d.qvel.double().cpu().numpy().T.copy())
    return out, int((d.flags != 0).sum())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for name in ("pendulum", "cartpole", "drone", "humanoid"):
        model = load_model(name)
        h = float(model.opt.timestep)
        for label, qpos, qvel, ctrl, pre in regimes(name, model, n):
            if pre:
                d = mj.BatchData(model, n)
                d.qpos.copy_(torch.as_tensor(qpos.T.copy(), device="cuda")); d.qvel.copy_(torch.as_tensor(qvel.T.copy(), device="cuda"))
                mj.mj_step(model, d, pre)
                qpos, qvel = d.qpos.cpu().numpy().T.copy(), d.qvel.cpu().numpy().T.copy()
            r64, bad64 = run(model, 64, qpos, qvel, ctrl)
            r32, bad32 = run(model, 32, qpos, qvel, ctrl)
            row = {"model": name, "regime": label, "nenv": n, "timestep": h, "flagged_envs": [bad64, bad32]}
            for k in CHECK:
                e = np.abs(r64[k][0] - r32[k][0]).max(axis=1)
                row[f"dq_max_{k}"] = float(e.max()); row[f"dq_median_{k}"] = float(np.median(e))
            lo, hi = max(row["dq_median_100"], 1e-12), max(row["dq_median_1000"], 1e-12)
            row["growth_rate_per_s"] = float(np.log(hi / lo) / (900 * h))
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
