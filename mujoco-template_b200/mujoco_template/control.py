"""Controller boundary of the hot path.

Keeps the reference's contract (reference ``mujoco_template/control.py:9-32``): a controller
exposes ``capabilities`` and two calls, ``prepare(model, data)`` once per reset and
``__call__(model, data, t)`` per control tick, and communicates only by writing
``data.ctrl``.  ``needs_linearization`` / ``needs_jacobians`` are the switches that make
``Env.step`` / ``BatchedEnv.step`` launch the FD-linearization and Jacobian kernels.

For ``BatchedEnv`` the same protocol applies with ``data`` being a ``BatchData`` whose
arrays are ``(dim, nenv)`` CUDA tensors; a vectorised controller writes ``data.ctrl[:]``.
"""

from __future__ import annotations

from collections.abc import Iterable
from dataclasses import dataclass, field
from typing import Any, Protocol, runtime_checkable


class ControlSpace:
    """String tokens naming what ``data.ctrl`` means to the actuators."""

    TORQUE = "torque"
    POSITION = "position"
    VELOCITY = "velocity"
    INTVELOCITY = "intvelocity"

    ALL = (TORQUE, POSITION, VELOCITY, INTVELOCITY)


@dataclass(frozen=True)
class ControllerCapabilities:
    """What a controller needs from the environment before each control tick."""

    control_space: str = ControlSpace.TORQUE
    needs_linearization: bool = False
    needs_jacobians: Iterable[str] = field(default_factory=tuple)
    actuator_groups: Iterable[int] | None = None


@runtime_checkable
class Controller(Protocol):
    capabilities: ControllerCapabilities

    def prepare(self, model: Any, data: Any) -> None:
        """Called at construction and after every reset."""

    def __call__(self, model: Any, data: Any, t: float) -> None:
        """Write ``data.ctrl`` for the state found in ``data`` at time ``t``."""


__all__ = ["ControlSpace", "ControllerCapabilities", "Controller"]
