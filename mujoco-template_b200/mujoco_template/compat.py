"""Controller / model compatibility report (host-side validation only, no arithmetic).

Same outcomes as reference ``mujoco_template/compat.py:31-127``: ``nu == 0`` and "all
actuators disabled" are hard failures; everything else is a warning string.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .control import ControlSpace, ControllerCapabilities
from .exceptions import CompatibilityError, ConfigError

_SERVO_SPACES = (ControlSpace.POSITION, ControlSpace.VELOCITY, ControlSpace.INTVELOCITY)


@dataclass
class CompatibilityReport:
    ok: bool
    reasons: list[str] = field(default_factory=list)
    warnings: list[str] = field(default_factory=list)

    def assert_ok(self) -> None:
        if self.ok:
            return
        lines = ["Incompatible controller/model:"] + [f"- {r}" for r in self.reasons]
        raise CompatibilityError("\n".join(lines))


def _valid_range(pair) -> bool:
    lo, hi = pair
    return bool(np.isfinite(lo) and np.isfinite(hi) and hi > lo)


def check_controller_compat(model: Any, ctrl_cap: ControllerCapabilities, enabled_mask: np.ndarray | None) -> CompatibilityReport:
    reasons: list[str] = []
    notes: list[str] = []
    nu = model.nu
    if nu == 0:
        reasons.append("Model has no actuators (nu=0).")
    mask = np.ones(nu, dtype=bool) if enabled_mask is None else enabled_mask
    if mask.shape[0] != nu:
        raise ConfigError("enabled_mask must have length model.nu")
    if mask.sum() == 0:
        reasons.append("All actuators are disabled by group selection.")
    active = np.flatnonzero(mask)

    if ctrl_cap.actuator_groups is not None:
        wanted = {int(g) for g in ctrl_cap.actuator_groups}
        live = {int(g) for g in np.asarray(model.actuator_group)[active]}
        if not live:
            notes.append("Controller declared actuator groups but none are currently enabled; continuing without additional group gating.")
        missing, extra = sorted(wanted - live), sorted(live - wanted)
        if missing:
            notes.append(f"Controller requested actuator groups {missing} but they are not enabled; controller will still run with the available groups.")
        if extra:
            notes.append(f"Enabled actuators include groups {extra} beyond the controller request; behaviour matches MuJoCo but may require controller-side masking.")

    space = ctrl_cap.control_space
    if space in _SERVO_SPACES:
        limited = np.asarray(model.actuator_ctrllimited, dtype=bool)
        for idx in active:
            if not limited[idx]:
                notes.append(f"Enabled actuator {idx} lacks ctrlrange limits required for servo control.")
            elif not _valid_range(model.actuator_ctrlrange[idx]):
                lo, hi = model.actuator_ctrlrange[idx]
                notes.append(f"Invalid ctrlrange for enabled actuator {idx}: [{lo}, {hi}]")
    if space == ControlSpace.INTVELOCITY:
        actlim = np.asarray(model.actuator_actlimited, dtype=bool)
        for idx in active:
            if not actlim[idx]:
                notes.append(f"Enabled actuator {idx} has no activation limits (actlimited=0) under intvelocity control.")
            elif not _valid_range(model.actuator_actrange[idx]):
                lo, hi = model.actuator_actrange[idx]
                notes.append(f"Invalid actrange for enabled actuator {idx}: [{lo}, {hi}]")
    if space == ControlSpace.TORQUE:
        flim = np.asarray(model.actuator_forcelimited, dtype=bool)
        for idx in np.flatnonzero(flim & mask):
            if not _valid_range(model.actuator_forcerange[idx]):
                lo, hi = model.actuator_forcerange[idx]
                notes.append(f"Invalid forcerange for enabled actuator {idx}: [{lo}, {hi}]")

    notes.append("Note: joint/tendon constraints or other clamps may still limit motion/force beyond actuator-level checks.")
    return CompatibilityReport(ok=not reasons, reasons=reasons, warnings=notes)


__all__ = ["CompatibilityReport", "check_controller_compat"]
