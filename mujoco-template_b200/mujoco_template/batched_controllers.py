"""Vectorised controllers for ``BatchedEnv`` written against the reference ``Controller`` protocol.

``BatchedLQRController`` is the controller BASELINE config #2 needs and the reference does not
ship for the cartpole (its cartpole uses PID, reference ``examples/cartpole/controllers/pid.py``):
a discrete LQR about a fixed setpoint whose gain comes from the FD linearisation at that
setpoint -- the same recipe as the reference's drone / humanoid LQRs
(``examples/drone/controllers/lqr.py:120-133``: ``linearize_discrete`` once in ``prepare`` then a
DARE) -- and which declares ``needs_linearization=True`` so that every control tick also
produces the time-varying ``(A, B)`` of all envs in ``StepResult.info``.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .control import ControlSpace, ControllerCapabilities
from .exceptions import ConfigError


def dlqr_gain(A: np.ndarray, B: np.ndarray, Q: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Infinite-horizon discrete LQR gain K (u = -K x) via SciPy's DARE solver."""
    from scipy.linalg import solve_discrete_are

    P = solve_discrete_are(A, B, Q, R)
    return np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)


class BatchedLQRController:
    def __init__(self, qpos_ref: np.ndarray | None = None, ctrl_ref: np.ndarray | None = None,
                 Q: np.ndarray | None = None, R: np.ndarray | None = None, eps: float = 1e-6,
                 needs_linearization: bool = True, fused: bool = True):
        self.qpos_ref = None if qpos_ref is None else np.asarray(qpos_ref, dtype=float)
        self.ctrl_ref = None if ctrl_ref is None else np.asarray(ctrl_ref, dtype=float)
        self.Q, self.R, self.eps = Q, R, float(eps)
        self.capabilities = ControllerCapabilities(control_space=ControlSpace.TORQUE,
                                                   needs_linearization=bool(needs_linearization))
        self.fused = bool(fused)  # False: the same tick as a handful of torch ops (reference-style arithmetic)
        self._gain_uploaded = False
        self.K: np.ndarray | None = None
        self._dev: dict[str, Any] = {}

    def prepare(self, model: Any, data: Any) -> None:
        from . import _mj as mj

        nv, nu = model.nv, model.nu
        if nu == 0:
            raise ConfigError("BatchedLQRController requires nu>0")
        qref = np.array(model.qpos0) if self.qpos_ref is None else self.qpos_ref
        uref = np.zeros(nu) if self.ctrl_ref is None else self.ctrl_ref
        if self.K is None:
            probe = mj.MjData(model)  # one env in device-mapped host memory
            probe.qpos[:] = qref
            probe.qvel[:] = 0.0
            probe.ctrl[:] = uref
            A = np.zeros((2 * nv, 2 * nv))
            B = np.zeros((2 * nv, nu))
            mj.mjd_transitionFD(model, probe, self.eps, True, A, B, None, None)
            Q = np.eye(2 * nv) if self.Q is None else np.asarray(self.Q, dtype=float)
            R = np.eye(nu) if self.R is None else np.asarray(self.R, dtype=float)
            self.A0, self.B0 = A, B
            self.K = dlqr_gain(A, B, Q, R)
        if hasattr(data.qpos, "device"):
            import torch

            dev, dt = data.qpos.device, data.qpos.dtype
            self._dev = dict(
                K=torch.as_tensor(self.K, device=dev, dtype=dt), qref=torch.as_tensor(qref, device=dev, dtype=dt)[:, None],
                uref=torch.as_tensor(uref, device=dev, dtype=dt)[:, None],
                lo=torch.as_tensor(model.actuator_ctrlrange[:, 0], device=dev, dtype=dt)[:, None],
                hi=torch.as_tensor(model.actuator_ctrlrange[:, 1], device=dev, dtype=dt)[:, None],
                limited=torch.as_tensor(np.asarray(model.actuator_ctrllimited, dtype=bool), device=dev)[:, None],
                qref_full=None, dq=torch.zeros((nv, data.qpos.shape[1]), device=dev, dtype=dt))
        self._gain_uploaded = False
        self._free = bool(np.any(model.jnt_type == 0))
        self._qref_np, self._uref_np = qref, uref

    def device_law_ready(self, data: Any) -> bool:
        """True when the control law lives on the device (b2_lqr_set_gain done): ``BatchedEnv`` may then fold this
        controller's tick, the (A, B) linearisation and the step into one ``b2_control_tick`` launch."""
        b = data.backend
        if not (self.fused and hasattr(b, "batch") and hasattr(data.qpos, "device")):
            return False
        if not self._gain_uploaded:
            b.batch.lqr_set_gain(self.K, self._qref_np, self._uref_np)
            self._gain_uploaded = True
        return True

    def __call__(self, model: Any, data: Any, t: float) -> None:
        b = data.backend
        if self.device_law_ready(data):
            # one kernel: tangent-space state error, gain product, ctrlrange clamp (b2_lqr_control)
            b._launch("lqr_control", b.batch.lqr_control, b.state_struct())
            return
        import torch

        d = self._dev
        if self._free:
            if d["qref_full"] is None:
                d["qref_full"] = d["qref"].expand_as(data.qpos).contiguous()
            b._pre()
            b.batch.differentiate_pos(d["dq"].data_ptr(), 1.0, d["qref_full"].data_ptr(), data.qpos.data_ptr(), b.stream)
            dq = d["dq"]
        else:
            dq = data.qpos - d["qref"]
        x = torch.cat([dq, data.qvel], dim=0)
        u = d["uref"] - d["K"] @ x
        u = torch.where(d["limited"], torch.minimum(torch.maximum(u, d["lo"]), d["hi"]), u)
        data.ctrl.copy_(u)


__all__ = ["BatchedLQRController", "dlqr_gain"]
