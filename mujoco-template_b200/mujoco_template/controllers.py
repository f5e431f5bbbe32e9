"""Built-in controllers (reference ``mujoco_template/controllers.py:12-46``).

Both work unchanged on a single ``MjData`` and on a ``BatchData``: ``data.ctrl`` has the
actuator index as its leading dimension in either case.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .control import ControlSpace, ControllerCapabilities
from .exceptions import CompatibilityError, ConfigError, TemplateError


def _require_actuators(model: Any, who: str) -> None:
    if model.nu == 0:
        raise CompatibilityError(f"{who} requires nu>0 to write controls.")


def _check_ctrl_shape(model: Any, data: Any) -> None:
    if data.ctrl.shape[0] != model.nu:
        raise TemplateError("data.ctrl size does not match model.nu")


@dataclass
class ZeroController:
    """Writes zeros to every actuator (the ``--zero`` CLI controller)."""

    capabilities: ControllerCapabilities = field(
        default_factory=lambda: ControllerCapabilities(control_space=ControlSpace.TORQUE))

    def prepare(self, model: Any, data: Any) -> None:
        _require_actuators(model, "ZeroController")

    def __call__(self, model: Any, data: Any, t: float) -> None:
        _check_ctrl_shape(model, data)
        data.ctrl[:] = 0.0


@dataclass
class PositionTargetDemo:
    """Holds fixed servo targets; defaults to the controls present at ``prepare`` time."""

    targets: np.ndarray | None = None
    capabilities: ControllerCapabilities = field(
        default_factory=lambda: ControllerCapabilities(control_space=ControlSpace.POSITION))

    def prepare(self, model: Any, data: Any) -> None:
        _require_actuators(model, "PositionTargetDemo")
        if self.targets is None:
            ctrl = np.asarray(data.ctrl.cpu() if hasattr(data.ctrl, "cpu") else data.ctrl)
            self.targets = np.array(ctrl[:, 0] if ctrl.ndim == 2 else ctrl) if ctrl.shape[0] == model.nu else np.zeros(model.nu)
        if self.targets.shape[0] != model.nu:
            raise ConfigError("targets must have length model.nu")

    def __call__(self, model: Any, data: Any, t: float) -> None:
        _check_ctrl_shape(model, data)
        if getattr(data.ctrl, "ndim", 1) == 2:
            if hasattr(data.ctrl, "device"):
                import torch

                data.ctrl[:] = torch.as_tensor(self.targets, dtype=data.ctrl.dtype, device=data.ctrl.device)[:, None]
            else:
                data.ctrl[:] = self.targets[:, None]
        else:
            data.ctrl[:] = self.targets


__all__ = ["ZeroController", "PositionTargetDemo"]
