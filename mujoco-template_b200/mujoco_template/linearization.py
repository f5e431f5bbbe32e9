"""Discrete-time (A, B) linearisation of one simulation step.

Reference behaviour (``mujoco_template/linearization.py:16-135``):

* ``use_native=True``  -> ``mjd_transitionFD(model, data, eps, centered=True, A, B)``; here a
  single kernel launch evaluates all ``2(2nv+nu)`` perturbed rollouts in parallel.
  ``horizon_steps`` is ignored on this path, as in the reference.
* ``use_native=False`` (or a ``LinearizationError`` from the native path) -> the reference's
  Python centred-difference loop, reproduced with its exact semantics: perturbations are not
  clamped to ``ctrlrange``, ``qacc_warmstart`` is not restored between rollouts
  (``state_utils`` does not snapshot it), ``horizon_steps`` is honoured, and the position rows
  use ``mj_differentiatePos(model, dq, 1, qpos_new, qpos_base)`` -- i.e. base minus new, the
  argument order the reference passes (``linearization.py:10-13``).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _mj as mj
from .exceptions import LinearizationError
from .state_utils import _restore_state, _snapshot_state


def _dqpos(model: Any, qpos2: np.ndarray, qpos1: np.ndarray) -> np.ndarray:
    dq = np.zeros(model.nv)
    mj.mj_differentiatePos(model, dq, 1.0, qpos2, qpos1)
    return dq


def _native_transition_fd(model: Any, data: Any, eps: float = 1e-6, centered: bool = True) -> tuple[np.ndarray, np.ndarray]:
    nx = 2 * model.nv
    A = np.zeros((nx, nx))
    B = np.zeros((nx, model.nu))
    try:
        mj.mjd_transitionFD(model, data, float(eps), bool(centered), A, B, None, None)
    except LinearizationError:
        raise
    except Exception as exc:
        raise LinearizationError(f"mjd_transitionFD failed: {exc}") from exc
    return A, B


def _fd_linearization(model: Any, data: Any, eps: float = 1e-6, horizon_steps: int = 1) -> tuple[np.ndarray, np.ndarray]:
    nv, nu = model.nv, model.nu
    snap = _snapshot_state(data)
    base_q, base_v, base_u = np.array(data.qpos), np.array(data.qvel), np.array(data.ctrl)

    def rollout() -> np.ndarray:
        mj.mj_forward(model, data)
        for _ in range(horizon_steps):
            mj.mj_step(model, data)
        return np.concatenate([_dqpos(model, np.array(data.qpos), base_q), np.array(data.qvel) - base_v])

    def column(apply) -> np.ndarray:
        sides = []
        for sign in (+1.0, -1.0):
            _restore_state(data, snap)
            data.qpos[:] = base_q
            data.qvel[:] = base_v
            apply(sign * eps)
            sides.append(rollout())
        return (sides[0] - sides[1]) / (2.0 * eps)

    def nudge_pos(idx: int):
        def apply(delta: float) -> None:
            q = np.copy(base_q)
            shift = np.zeros(nv)
            shift[idx] = delta
            mj.mj_integratePos(model, q, shift, 1.0)
            data.qpos[:] = q
        return apply

    def nudge_vel(idx: int):
        def apply(delta: float) -> None:
            data.qvel[idx] += delta
        return apply

    def nudge_ctrl(idx: int):
        def apply(delta: float) -> None:
            u = np.copy(base_u)
            u[idx] += delta
            data.ctrl[:] = u
        return apply

    try:
        A = np.zeros((2 * nv, 2 * nv))
        B = np.zeros((2 * nv, nu))
        for i in range(nv):
            A[:, i] = column(nudge_pos(i))
        for i in range(nv):
            A[:, nv + i] = column(nudge_vel(i))
        for i in range(nu):
            B[:, i] = column(nudge_ctrl(i))
        return A, B
    finally:
        _restore_state(data, snap)
        mj.mj_forward(model, data)


def linearize_discrete(model: Any, data: Any, use_native: bool = True, eps: float = 1e-6,
                       horizon_steps: int = 1) -> tuple[np.ndarray, np.ndarray]:
    if use_native:
        try:
            return _native_transition_fd(model, data, eps=eps, centered=True)
        except LinearizationError:
            pass
    return _fd_linearization(model, data, eps=eps, horizon_steps=horizon_steps)


__all__ = ["linearize_discrete"]
