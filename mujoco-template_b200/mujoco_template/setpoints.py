"""Steady-state controls by inverse dynamics (reference ``mujoco_template/setpoints.py:10-58``).

``steady_ctrl0`` asks which controls hold ``(qpos0, qvel0)`` with zero acceleration: one ``mj_inverse``
(a ``b2_inverse`` kernel launch) gives the generalized force, and the least-squares solution through the
pseudo-inverse of the dense actuator moment matrix gives the controls.  The caller's state is preserved.
``batched_steady_ctrl0`` does the same for every env of a ``BatchedEnv`` at its current state.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _mj as mj
from .exceptions import CompatibilityError, ConfigError, TemplateError
from .state_utils import _restore_state, _snapshot_state


def steady_ctrl0(model: Any, data: Any, qpos0: np.ndarray, qvel0: np.ndarray | None = None) -> np.ndarray:
    if qpos0.shape[0] != model.nq:
        raise ConfigError("qpos0 must have length model.nq")
    qvel0 = np.zeros(model.nv) if qvel0 is None else qvel0
    if qvel0.shape[0] != model.nv:
        raise ConfigError("qvel0 must have length model.nv")
    snap = _snapshot_state(data)
    try:
        mj.mj_resetData(model, data)
        data.qpos[:] = qpos0
        data.qvel[:] = qvel0
        mj.mj_forward(model, data)
        data.qacc[:] = 0.0
        mj.mj_inverse(model, data)
        qfrc = np.array(data.qfrc_inverse)
        if model.nu == 0:
            raise CompatibilityError("No actuators to realize inverse dynamics (nu=0).")
        moment = np.array(data.actuator_moment).reshape(model.nu, model.nv)
        u = (np.atleast_2d(qfrc) @ np.linalg.pinv(moment)).ravel()
        sv = np.linalg.svd(moment, compute_uv=False)
        if sv.size == 0 or (sv.min() / sv.max() if sv.max() > 0 else 0.0) < 1e-12:
            raise TemplateError("Actuator moment matrix is singular or ill-conditioned at this state.")
        return u
    finally:
        _restore_state(data, snap)
        mj.mj_forward(model, data)


def batched_steady_ctrl0(env: Any):
    """Controls that hold every env of a ``BatchedEnv`` at its current ``(qpos, qvel)``: ``(nu, nenv)`` tensor."""
    import torch

    model, data = env.model, env.data
    if model.nu == 0:
        raise CompatibilityError("No actuators to realize inverse dynamics (nu=0).")
    data.qacc.zero_()
    mj.mj_inverse(model, data)
    n = data.qpos.shape[1]
    moment = data.actuator_moment.reshape(model.nu, model.nv, n).permute(2, 0, 1)  # (n, nu, nv)
    qfrc = data.qfrc_inverse.t().unsqueeze(1)                                        # (n, 1, nv)
    return (qfrc @ torch.linalg.pinv(moment)).squeeze(1).t().contiguous()


__all__ = ["steady_ctrl0", "batched_steady_ctrl0"]
