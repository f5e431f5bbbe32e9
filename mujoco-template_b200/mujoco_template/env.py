"""Single-environment step loop (reference ``mujoco_template/env.py:20-260``).

Behaviour kept from the reference:

* per substep, and only on control ticks (``_substep % control_decimation == 0``):
  controller -> optional (A, B) -> optional Jacobians; then one physics step;
* ``info['A']``/``info['B']``/``info['jacobians']`` hold a single object when exactly one was
  produced in this call, otherwise a list;
* ``compat_warnings`` appear once, in the first ``StepResult`` after construction/reset;
* ``info_fn`` keys may not collide with built-in ones.

The physics, the FD linearisation and the Jacobians run as CUDA kernels on the env's
device-mapped state (``_mj.MjData``).
"""

from __future__ import annotations

import warnings
from collections.abc import Callable, Iterable, Iterator
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _mj as mj
from ._typing import InfoDict, JacobiansDict, Observation
from .compat import check_controller_compat
from .control import Controller
from .exceptions import ConfigError, TemplateError
from .jacobians import compute_requested_jacobians
from .linearization import linearize_discrete
from .model import ModelHandle
from .observations import ObservationExtractor, ObservationSpec


@dataclass
class StepResult:
    """What one ``Env.step`` call hands back (reference ``env.py:20-25``)."""

    obs: Observation | None
    reward: float | None
    done: bool
    info: InfoDict


def _one_or_many(items: list) -> Any:
    """The reference's shape rule for per-tick products: the object itself if there is exactly one, else the list."""
    return items[0] if len(items) == 1 else items


class _TickProducts:
    """Linearisations and Jacobians gathered over the control ticks of one ``step(n)`` call."""

    __slots__ = ("A", "B", "jac")

    def __init__(self) -> None:
        self.A: list[np.ndarray] = []
        self.B: list[np.ndarray] = []
        self.jac: list[JacobiansDict] = []

    def into(self, info: InfoDict) -> None:
        if self.A:
            info["A"] = _one_or_many(self.A)
            info["B"] = _one_or_many(self.B)
        if self.jac:
            info["jacobians"] = _one_or_many(self.jac)


def _declared_groups(controller: Controller | None) -> tuple[int, ...] | None:
    groups = None if controller is None else controller.capabilities.actuator_groups
    return None if groups is None else tuple(int(g) for g in groups)


class Env:
    """One environment: model handle + controller + observation extractor + optional reward / done / info hooks."""

    def __init__(
        self,
        handle: ModelHandle,
        obs_spec: ObservationSpec | None = None,
        controller: Controller | None = None,
        reward_fn: Callable[[Any, Any, Observation | None], float] | None = None,
        done_fn: Callable[[Any, Any, Observation | None], bool] | None = None,
        info_fn: Callable[[Any, Any, Observation | None], dict] | None = None,
        enabled_groups: Iterable[int] | None = None,
        control_decimation: int = 1,
    ):
        if control_decimation < 1:
            raise ConfigError("control_decimation must be >= 1")
        self.handle, self.model, self.data = handle, handle.model, handle.data
        self.controller = controller
        self.control_decimation = int(control_decimation)
        self.reward_fn, self.done_fn, self.info_fn = reward_fn, done_fn, info_fn
        self._obs_spec = obs_spec
        self.extractor: ObservationExtractor | None = None if obs_spec is None else ObservationExtractor(handle.model, obs_spec)
        self._substep = 0
        self._compat_warnings: list[str] = []
        self._added_warnings = False  # compat_warnings go into the first StepResult after construction / reset only
        self._apply_group_selection(enabled_groups)
        if controller is not None:
            self._prepare_and_check(controller)

    # ------------------------------------------------------------------ construction helpers
    def _warn(self, msg: str) -> None:
        warnings.warn(msg, RuntimeWarning)
        self._compat_warnings.append(msg)

    def _apply_group_selection(self, enabled_groups: Iterable[int] | None) -> None:
        declared = _declared_groups(self.controller)
        if enabled_groups is None:
            if declared is not None:
                self._warn(f"Controller declares actuator groups {sorted(set(declared))} but Env leaves actuator "
                           "availability unchanged by default.")
            return
        chosen = tuple(int(g) for g in enabled_groups)
        if declared is not None and set(declared) != set(chosen):
            self._warn(f"Controller declares actuator groups {sorted(set(declared))} but user requested "
                       f"{sorted(set(chosen))}; proceeding with the user selection.")
        self.handle.set_enabled_actuator_groups(chosen)

    def _prepare_and_check(self, controller: Controller) -> None:
        controller.prepare(self.model, self.data)
        report = check_controller_compat(self.model, controller.capabilities, self.handle.enabled_actuator_mask())
        self._compat_warnings = list(report.warnings)  # replaces the group warnings, as reference env.py:92 does
        report.assert_ok()

    @property
    def compat_warnings(self) -> list[str]:
        return list(self._compat_warnings)

    @classmethod
    def from_xml_path(
        cls,
        xml_path: str,
        *,
        obs_spec: ObservationSpec | None = None,
        controller: Controller | None = None,
        reward_fn=None,
        done_fn=None,
        info_fn=None,
        enabled_groups: Iterable[int] | None = None,
        control_decimation: int = 1,
        auto_reset: bool = True,
        keyframe: int | str | None = None,
    ) -> "Env":
        """Compile the MJCF file, build the env around it and reset it (unless ``auto_reset=False``).

        Without an ``obs_spec`` the observation leaves out ``sensordata`` (reference ``env.py:122-123``).
        """
        if keyframe is not None and not auto_reset:
            raise ConfigError("auto_reset=False is incompatible with specifying a keyframe")
        env = cls(ModelHandle.from_xml_path(xml_path),
                  obs_spec=ObservationSpec(include_sensordata=False) if obs_spec is None else obs_spec,
                  controller=controller, reward_fn=reward_fn, done_fn=done_fn, info_fn=info_fn,
                  enabled_groups=enabled_groups, control_decimation=control_decimation)
        if auto_reset:
            env.reset(keyframe)
        return env

    # ------------------------------------------------------------------ observations
    def _observe(self) -> Observation:
        if self.extractor is None:
            self._obs_spec = self._obs_spec or ObservationSpec()
            self.extractor = ObservationExtractor(self.model, self._obs_spec)
        return self.extractor(self.data)

    # ------------------------------------------------------------------ reset / step
    def reset(self, keyframe: int | str | None = None) -> Observation:
        """``mj_resetData[Keyframe]`` + ``mj_forward``; the controller is prepared again (reference ``env.py:152-162``)."""
        if keyframe is None:
            self.handle.reset()
        else:
            self.handle.reset_keyframe(keyframe)
        self.handle.forward()
        self._substep, self._added_warnings = 0, False
        if self.controller is not None:
            self.controller.prepare(self.model, self.data)
        return self._observe()

    def _control_tick(self, out: _TickProducts) -> None:
        """Controller, then what its capabilities ask for -- (A, B) at the new controls, then Jacobians."""
        controller = self.controller
        controller(self.model, self.data, float(self.data.time))
        caps = controller.capabilities
        if caps.needs_linearization:
            A, B = linearize_discrete(self.model, self.data, use_native=True)
            out.A.append(A)
            out.B.append(B)
        if caps.needs_jacobians:
            out.jac.append(compute_requested_jacobians(self.model, self.data, caps.needs_jacobians))

    def step(self, n: int = 1, *, return_obs: bool = True) -> StepResult:
        if n < 1:
            raise ConfigError("Env.step(n): n must be >= 1")
        info: InfoDict = {}
        if self._compat_warnings and not self._added_warnings:
            info["compat_warnings"] = list(self._compat_warnings)
            self._added_warnings = True
        products = _TickProducts()
        for _ in range(n):
            if self.controller is not None and self._substep % self.control_decimation == 0:
                self._control_tick(products)
            self.handle.step()
            self._substep += 1
        products.into(info)
        obs = self._observe() if return_obs else None
        reward = self.reward_fn(self.model, self.data, obs) if self.reward_fn else None
        done = bool(self.done_fn(self.model, self.data, obs)) if self.done_fn else False
        if self.info_fn:
            extra = self.info_fn(self.model, self.data, obs)
            clash = next((k for k in extra if k in info), None)
            if clash is not None:
                raise TemplateError(f"info key collision: {clash}")
            info.update(extra)
        return StepResult(obs=obs, reward=reward, done=done, info=info)

    def linearize(self, eps: float = 1e-6, horizon_steps: int = 1) -> tuple[np.ndarray, np.ndarray]:
        """(A, B) of one step at the current state and controls: ``b2_linearize`` behind ``linearize_discrete``."""
        return linearize_discrete(self.model, self.data, use_native=True, eps=eps, horizon_steps=horizon_steps)

    def passive(self, *, duration: float | None = None, max_steps: int | None = None, hooks=None,
                return_obs: bool = True) -> Iterator[StepResult]:
        """Generator over steps until ``duration`` / ``max_steps`` (``runtime.iterate_passive``)."""
        from .runtime import iterate_passive

        yield from iterate_passive(self, duration=duration, max_steps=max_steps, hooks=hooks, return_obs=return_obs)


__all__ = ["Env", "StepResult"]
