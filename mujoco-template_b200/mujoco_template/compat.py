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


def _range_notes(model: Any, indices: np.ndarray, flag_field: str, range_field: str, *, unflagged: str | None, label: str) -> list[str]:
    """One pass over the enabled actuators for a (limited-flag, range) field pair of the model.

    ``unflagged``: message template for an actuator whose flag is off (``None``: such actuators are fine and skipped);
    a flagged actuator must have a finite, non-empty range, else it is reported under ``label``."""
    flags = np.asarray(getattr(model, flag_field), dtype=bool)
    ranges = np.asarray(getattr(model, range_field), dtype=float)
    out: list[str] = []
    for idx in indices:
        if not flags[idx]:
            if unflagged is not None:
                out.append(unflagged.format(idx=idx))
            continue
        lo, hi = ranges[idx]
        if not (np.isfinite(lo) and np.isfinite(hi) and hi > lo):
            out.append(f"Invalid {label} for enabled actuator {idx}: [{lo}, {hi}]")
    return out


def _group_notes(model: Any, declared, active: np.ndarray) -> list[str]:
    wanted = {int(g) for g in declared}
    live = {int(g) for g in np.asarray(model.actuator_group)[active]}
    out: list[str] = []
    if not live:
        out.append("Controller declared actuator groups but none are currently enabled; continuing without additional group gating.")
    if wanted - live:
        out.append(f"Controller requested actuator groups {sorted(wanted - live)} but they are not enabled; controller will still run with the available groups.")
    if live - wanted:
        out.append(f"Enabled actuators include groups {sorted(live - wanted)} beyond the controller request; behaviour matches MuJoCo but may require controller-side masking.")
    return out


def check_controller_compat(model: Any, ctrl_cap: ControllerCapabilities, enabled_mask: np.ndarray | None) -> CompatibilityReport:
    """Hard failures: no actuators, or none enabled.  Everything else becomes a warning string (reference wording)."""
    nu = int(model.nu)
    mask = np.ones(nu, dtype=bool) if enabled_mask is None else np.asarray(enabled_mask)
    if mask.shape[0] != nu:
        raise ConfigError("enabled_mask must have length model.nu")
    reasons = [msg for bad, msg in ((nu == 0, "Model has no actuators (nu=0)."),
                                    (not mask.any(), "All actuators are disabled by group selection.")) if bad]
    active = np.flatnonzero(mask)
    notes: list[str] = []
    if ctrl_cap.actuator_groups is not None:
        notes += _group_notes(model, ctrl_cap.actuator_groups, active)
    space = ctrl_cap.control_space
    if space in _SERVO_SPACES:
        notes += _range_notes(model, active, "actuator_ctrllimited", "actuator_ctrlrange", label="ctrlrange",
                              unflagged="Enabled actuator {idx} lacks ctrlrange limits required for servo control.")
    if space == ControlSpace.INTVELOCITY:
        notes += _range_notes(model, active, "actuator_actlimited", "actuator_actrange", label="actrange",
                              unflagged="Enabled actuator {idx} has no activation limits (actlimited=0) under intvelocity control.")
    if space == ControlSpace.TORQUE:
        notes += _range_notes(model, active, "actuator_forcelimited", "actuator_forcerange", label="forcerange", unflagged=None)
    notes.append("Note: joint/tendon constraints or other clamps may still limit motion/force beyond actuator-level checks.")
    return CompatibilityReport(ok=not reasons, reasons=reasons, warnings=notes)


__all__ = ["CompatibilityReport", "check_controller_compat"]
