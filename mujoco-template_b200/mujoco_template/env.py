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
    obs: Observation | None
    reward: float | None
    done: bool
    info: InfoDict


def _one_or_many(items: list) -> Any:
    return items[0] if len(items) == 1 else items


class Env:
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
        self.handle = handle
        self.model = handle.model
        self.data = handle.data
        self._obs_spec = obs_spec
        self.extractor: ObservationExtractor | None = ObservationExtractor(handle.model, obs_spec) if obs_spec is not None else None
        self.controller = controller
        self.reward_fn, self.done_fn, self.info_fn = reward_fn, done_fn, info_fn
        self.control_decimation = int(control_decimation)
        self._substep = 0
        self._added_warnings = False
        self._compat_warnings: list[str] = []
        self._configure_groups(enabled_groups)
        if controller is not None:
            controller.prepare(self.model, self.data)
            report = check_controller_compat(self.model, controller.capabilities, self.handle.enabled_actuator_mask())
            self._compat_warnings = list(report.warnings)
            report.assert_ok()

    def _configure_groups(self, enabled_groups: Iterable[int] | None) -> None:
        caps_groups = None
        if self.controller is not None and self.controller.capabilities.actuator_groups is not None:
            caps_groups = tuple(int(g) for g in self.controller.capabilities.actuator_groups)
        if enabled_groups is not None:
            chosen = tuple(int(g) for g in enabled_groups)
            if caps_groups is not None and set(caps_groups) != set(chosen):
                self._note(
                    f"Controller declares actuator groups {sorted(set(caps_groups))} but user requested "
                    f"{sorted(set(chosen))}; proceeding with the user selection.")
            self.handle.set_enabled_actuator_groups(chosen)
        elif caps_groups is not None:
            self._note(
                f"Controller declares actuator groups {sorted(set(caps_groups))} but Env leaves actuator "
                "availability unchanged by default.")

    def _note(self, msg: str) -> None:
        warnings.warn(msg, RuntimeWarning)
        self._compat_warnings.append(msg)

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
        """Load an MJCF file, build the env and (by default) reset it.

        ``obs_spec`` defaults to ``ObservationSpec(include_sensordata=False)``.
        """
        if obs_spec is None:
            obs_spec = ObservationSpec(include_sensordata=False)
        env = cls(
            ModelHandle.from_xml_path(xml_path), obs_spec=obs_spec, controller=controller, reward_fn=reward_fn,
            done_fn=done_fn, info_fn=info_fn, enabled_groups=enabled_groups, control_decimation=control_decimation)
        if keyframe is not None and not auto_reset:
            raise ConfigError("auto_reset=False is incompatible with specifying a keyframe")
        if auto_reset:
            env.reset(keyframe)
        return env

    def _ensure_extractor(self) -> ObservationExtractor:
        if self.extractor is None:
            if self._obs_spec is None:
                self._obs_spec = ObservationSpec()
            self.extractor = ObservationExtractor(self.model, self._obs_spec)
        return self.extractor

    def reset(self, keyframe: int | str | None = None) -> Observation:
        if keyframe is None:
            self.handle.reset()
        else:
            self.handle.reset_keyframe(keyframe)
        self.handle.forward()
        self._substep = 0
        self._added_warnings = False
        if self.controller is not None:
            self.controller.prepare(self.model, self.data)
        return self._ensure_extractor()(self.data)

    def step(self, n: int = 1, *, return_obs: bool = True) -> StepResult:
        if n < 1:
            raise ConfigError("Env.step(n): n must be >= 1")
        info: InfoDict = {}
        if self._compat_warnings and not self._added_warnings:
            info["compat_warnings"] = list(self._compat_warnings)
            self._added_warnings = True

        lin_A: list[np.ndarray] = []
        lin_B: list[np.ndarray] = []
        jacs: list[JacobiansDict] = []
        for _ in range(n):
            if self.controller is not None and self._substep % self.control_decimation == 0:
                self.controller(self.model, self.data, float(self.data.time))
                caps = self.controller.capabilities
                if caps.needs_linearization:
                    A, B = linearize_discrete(self.model, self.data, use_native=True)
                    lin_A.append(A)
                    lin_B.append(B)
                if caps.needs_jacobians:
                    jacs.append(compute_requested_jacobians(self.model, self.data, caps.needs_jacobians))
            self.handle.step()
            self._substep += 1
        if lin_A:
            info["A"], info["B"] = _one_or_many(lin_A), _one_or_many(lin_B)
        if jacs:
            info["jacobians"] = _one_or_many(jacs)

        obs = self._ensure_extractor()(self.data) if return_obs else None
        reward: float | None = None
        done = False
        if self.reward_fn:
            reward = self.reward_fn(self.model, self.data, obs)
        if self.done_fn:
            done = bool(self.done_fn(self.model, self.data, obs))
        if self.info_fn:
            for key, value in self.info_fn(self.model, self.data, obs).items():
                if key in info:
                    raise TemplateError(f"info key collision: {key}")
                info[key] = value
        return StepResult(obs=obs, reward=reward, done=done, info=info)

    def linearize(self, eps: float = 1e-6, horizon_steps: int = 1) -> tuple[np.ndarray, np.ndarray]:
        return linearize_discrete(self.model, self.data, use_native=True, eps=eps, horizon_steps=horizon_steps)

    def passive(self, *, duration: float | None = None, max_steps: int | None = None, hooks=None,
                return_obs: bool = True) -> Iterator[StepResult]:
        """Yield steps via :func:`runtime.iterate_passive` (``return_obs=False`` skips extraction)."""
        from .runtime import iterate_passive

        yield from iterate_passive(self, duration=duration, max_steps=max_steps, hooks=hooks, return_obs=return_obs)


__all__ = ["Env", "StepResult"]
