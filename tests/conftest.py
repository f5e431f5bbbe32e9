"""Shared test plumbing.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mujoco-template_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_NAMES = ("pendulum", "cartpole", "drone", "humanoid")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_model(name: str):
    from mujoco_template import _mj as mj

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel.from_compiled(os.path.join(GOLDEN, "models", f"{name}.b2m"))


def oracle_for(model):
    from oracle.oracle import OracleData, OracleModel

    om = OracleModel(model.blob, dict(nq=model.nq, nv=model.nv, nu=model.nu, nbody=model.nbody, njnt=model.njnt,
                                      ngeom=model.ngeom, nsite=model.nsite, ntendon=model.ntendon))
    return om, OracleData(om)


def random_states(model, name: str, n: int, seed: int = 0):
    """Synthetic initial states / controls per SURVEY.md section 8(d).  Returns AoS (n, dim) arrays."""
    rng = np.random.default_rng(seed)
    nq, nv, nu = model.nq, model.nv, model.nu
    qpos = np.tile(model.qpos0, (n, 1))
    qvel = np.zeros((n, nv))
    ctrl = np.zeros((n, nu))
    if name == "pendulum":
        qpos[:, 0] = rng.uniform(-np.pi, np.pi, n)
        qvel[:, 0] = rng.uniform(-2, 2, n)
        ctrl[:, 0] = rng.uniform(-4, 4, n)  # beyond the +-3 force clamp on purpose
    elif name == "cartpole":
        qpos[:, 0] = rng.uniform(-1, 1, n)
        qpos[:, 1] = rng.uniform(-0.2, 0.2, n)
        qvel[:] = rng.uniform(-0.5, 0.5, (n, 2))
        ctrl[:, 0] = rng.uniform(-1, 1, n)
    elif name == "drone":
        key = model.key_qpos[0]
        qpos[:] = key
        qpos[:, 2] = rng.uniform(1, 3, n)
        rv = rng.normal(0, 0.1, (n, 3))
        ang = np.linalg.norm(rv, axis=1, keepdims=True)
        qpos[:, 3] = np.cos(ang[:, 0] / 2)
        qpos[:, 4:7] = rv / np.maximum(ang, 1e-12) * np.sin(ang / 2)
        qvel[:] = rng.normal(0, 0.1, (n, nv))
        ctrl[:] = rng.uniform(0, 13, (n, nu))
    elif name == "humanoid":
        qpos[:] = model.key_qpos[1]  # stand_on_left_leg
        qpos[:, 7:] += rng.normal(0, 0.02, (n, nq - 7))
        qvel[:] = rng.normal(0, 0.01, (n, nv))
        ctrl[:] = rng.uniform(-1, 1, (n, nu)) * 0.2
    else:
        qpos[:, :] += rng.normal(0, 0.1, (n, nq))
        qvel[:] = rng.normal(0, 0.1, (n, nv))
        ctrl[:] = rng.uniform(-1, 1, (n, nu))
    return qpos, qvel, ctrl


@pytest.fixture(params=MODEL_NAMES)
def model_name(request):
    return request.param
