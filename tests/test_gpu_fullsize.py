"""BASELINE.json's full batch sizes, checked through size-independent properties.

The oracle cannot step 65,536 ... 262,144 envs in test time, so each full-size batch is a small set of distinct seeded
states tiled over the batch:
  * every replica of a state must come out bit-identical (an env's result may not depend on its position in the batch,
    on the warp or lock-step partner it shares hardware with, or on the work-queue order -- and every env must have been
    stepped exactly once);
  * the distinct states are compared with the oracle at the usual tolerances;
  * invariants of the domain: unit quaternions, contact counts, (A, B) block structure under semi-implicit Euler.
"""
import numpy as np
import pytest

from conftest import load_model, oracle_for, random_states
from test_gpu_parity import AB_RTOL, STEP_RTOL, _batch, _rel, _upload

pytestmark = pytest.mark.gpu

# (model, full batch size of the BASELINE config, distinct states, steps)
FULL = [("cartpole", 65536, 512, 20), ("drone", 262144, 256, 20), ("humanoid", 16384, 128, 6)]


def _tiled(model, name, n, distinct, seed):
    qpos, qvel, ctrl = random_states(model, name, distinct, seed=seed)
    reps = n // distinct
    assert reps * distinct == n
    return qpos, qvel, ctrl, (np.tile(qpos, (reps, 1)), np.tile(qvel, (reps, 1)), np.tile(ctrl, (reps, 1)))


@pytest.mark.parametrize("name,n,distinct,nsteps", FULL)
def test_full_size_rollout_replicas_identical_and_match_oracle(name, n, distinct, nsteps):
    from mujoco_template import _mj as mj

    model = load_model(name)
    qpos, qvel, ctrl, (Q, V, U) = _tiled(model, name, n, distinct, seed=21)
    data = _batch(model, n)
    _upload(data, Q, V, U)
    for _ in range(nsteps):
        mj.mj_step(model, data)
    gq = data.qpos.cpu().numpy().T.reshape(n // distinct, distinct, model.nq)
    gv = data.qvel.cpu().numpy().T.reshape(n // distinct, distinct, model.nv)
    gw = data.qacc_warmstart.cpu().numpy().T.reshape(n // distinct, distinct, model.nv)
    assert int(data.flags.max().item()) == 0
    # replicas: bit-identical
    assert np.array_equal(gq, np.broadcast_to(gq[0], gq.shape))
    assert np.array_equal(gv, np.broadcast_to(gv[0], gv.shape))
    assert np.array_equal(gw, np.broadcast_to(gw[0], gw.shape))
    # a sample of the distinct states against the oracle (free-running: roundoff accumulates over nsteps)
    om, od = oracle_for(model)
    sample = range(0, distinct, max(1, distinct // 24))
    tol = 1e-7 if name == "humanoid" else 1e-9 * nsteps
    for e in sample:
        od.reset()
        od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
        od.step(nsteps)
        assert _rel(gq[0, e], od.qpos) <= tol, (name, e, _rel(gq[0, e], od.qpos))
        assert _rel(gv[0, e], od.qvel) <= 10 * tol, (name, e, _rel(gv[0, e], od.qvel))
    # invariants
    if name in ("drone", "humanoid"):
        quat = gq[..., 3:7]
        assert np.max(np.abs(np.linalg.norm(quat, axis=-1) - 1.0)) < 1e-12
    if name == "humanoid":
        mj.mj_forward(model, data)
        ncon = data.ncon.cpu().numpy().reshape(n // distinct, distinct)
        assert np.array_equal(ncon, np.broadcast_to(ncon[0], ncon.shape)) and ncon.max() >= 2


@pytest.mark.parametrize("name,n,distinct", [("cartpole", 65536, 256), ("drone", 32768, 64)])
def test_full_size_linearize_replicas_identical_and_match_oracle(name, n, distinct):
    model = load_model(name)
    qpos, qvel, ctrl, (Q, V, U) = _tiled(model, name, n, distinct, seed=22)
    data = _batch(model, n)
    _upload(data, Q, V, U)
    A, B = data.backend.linearize(1e-6, True)
    nx, nu = 2 * model.nv, model.nu
    A = A.cpu().numpy().reshape(nx, nx, n // distinct, distinct)
    B = B.cpu().numpy().reshape(nx, nu, n // distinct, distinct)
    assert np.array_equal(A, np.broadcast_to(A[:, :, :1], A.shape))
    assert np.array_equal(B, np.broadcast_to(B[:, :, :1], B.shape))
    om, od = oracle_for(model)
    for e in range(0, distinct, max(1, distinct // 16)):
        od.reset()
        od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
        Ao, Bo = od.transition_fd(1e-6, True)
        assert _rel(A[:, :, 0, e], Ao) <= AB_RTOL and _rel(B[:, :, 0, e], Bo) <= AB_RTOL
    # semi-implicit Euler: d(qpos')/d(qvel) = h * d(qvel')/d(qvel) for hinge / slide coordinates (cartpole), and the
    # position rows of B are h times its velocity rows
    if name == "cartpole":
        h = float(model.opt.timestep)
        nv = model.nv
        assert np.max(np.abs(A[:nv, nv:] - h * A[nv:, nv:])) < 1e-6
        assert np.max(np.abs(B[:nv] - h * B[nv:])) < 1e-6


def test_full_size_control_tick_equals_separate_launches():
    """cartpole config #2 at full size: the fused control tick (LQR law + FD + advance in one launch) leaves the same
    state and (A, B), to roundoff, as controller launch + b2_linearize + b2_step."""
    import torch
    from mujoco_template import BatchedEnv
    from mujoco_template.batched_controllers import BatchedLQRController

    model = load_model("cartpole")
    n = 65536
    qpos, qvel, _ = random_states(model, "cartpole", n, seed=23)
    out = {}
    for fused in (True, False):
        ctl = BatchedLQRController(Q=np.diag([10.0, 100.0, 1.0, 1.0]), R=np.array([[0.01]]))
        env = BatchedEnv(model, n, controller=ctl, device=0)
        env.reset()
        env.fuse_control_tick = fused
        env.data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device="cuda"))
        env.data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device="cuda"))
        env.forward()
        for _ in range(5):
            res = env.step(return_obs=False)
        out[fused] = (env.data.qpos.clone(), env.data.qvel.clone(), res.info["A"].clone(), res.info["B"].clone())
    # two different kernels advance the env (FMA contraction differs): roundoff-level agreement of the state, and of
    # (A, B) up to that roundoff amplified by 1 / eps
    for k, (a, b) in enumerate(zip(out[True], out[False])):
        bound = (1e-12 if k < 2 else 1e-7) * max(1.0, float(b.abs().max()))
        assert float((a - b).abs().max()) <= bound, (k, float((a - b).abs().max()))
    # the LQR loop keeps every pole up
    assert float(out[True][0][1].abs().max()) < 0.25
