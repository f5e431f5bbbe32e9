"""ModelHandle: owns one ``(model, data)`` pair and forwards the hot-path calls to the GPU.

Same surface as reference ``mujoco_template/model.py:11-105``; ``step``/``forward``/``reset``
land in ``libb2mj.so`` kernels instead of ``mj_step``/``mj_forward``/``mj_resetData``.
``from_binary_path``/``save_binary`` use this package's own compiled-model file, not MJB.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from . import _mj as mj
from .exceptions import CompatibilityError, ConfigError, NameLookupError, TemplateError


class ModelHandle:
    def __init__(self, model: mj.MjModel, data: mj.MjData | None = None):
        self.model = model
        if data is None:
            data = mj.MjData(model)
        elif data.model is not model:
            raise ConfigError("Provided mj.MjData must reference the supplied model.")
        self.data = data

    # ---- loaders
    @classmethod
    def from_xml_path(cls, xml_path: str) -> "ModelHandle":
        return cls(mj.MjModel.from_xml_path(xml_path))

    @classmethod
    def from_xml_string(cls, xml_text: str) -> "ModelHandle":
        return cls(mj.MjModel.from_xml_string(xml_text))

    @classmethod
    def from_binary_path(cls, mjb_path: str) -> "ModelHandle":
        """Load a compiled-model file written by :meth:`save_binary` (not MuJoCo's MJB format)."""
        return cls(mj.MjModel.from_compiled(mjb_path))

    @classmethod
    def from_model_and_data(cls, model: mj.MjModel, data: mj.MjData) -> "ModelHandle":
        """Adopt an existing pair; no buffers are allocated."""
        return cls(model, data=data)

    def save_binary(self, mjb_path: str) -> None:
        try:
            self.model.save_compiled(mjb_path)
        except OSError as exc:
            raise TemplateError(f"mj_saveModel failed for {mjb_path}: {exc}") from exc

    # ---- hot path
    def forward(self) -> None:
        mj.mj_forward(self.model, self.data)

    def step(self) -> None:
        mj.mj_step(self.model, self.data)

    def reset(self) -> None:
        mj.mj_resetData(self.model, self.data)

    def reset_keyframe(self, key: int | str) -> None:
        if isinstance(key, str):
            idx = mj.mj_name2id(self.model, mj.mjtObj.mjOBJ_KEY, key)
            if idx < 0:
                raise NameLookupError(f"Keyframe name not found: {key}")
        else:
            idx = int(key)
            if idx < 0 or idx >= self.model.nkey:
                raise ConfigError(f"Keyframe index out of range: {idx}")
        mj.mj_resetDataKeyframe(self.model, self.data, idx)

    # ---- actuator groups (API surface; the mask is applied inside the actuation stage)
    @property
    def actuator_groups(self) -> np.ndarray:
        return np.array(self.model.actuator_group, dtype=int)

    def set_enabled_actuator_groups(self, enabled_groups: Iterable[int]) -> None:
        wanted = {int(g) for g in enabled_groups}
        if not wanted:
            raise CompatibilityError("At least one actuator group must be enabled.")
        if min(wanted) < 0 or max(wanted) > 31:
            raise ConfigError("Actuator groups must be in [0, 31].")
        if self.model.nu == 0:
            raise CompatibilityError("Model has no actuators (nu=0).")
        present = {int(g) for g in self.model.actuator_group[: self.model.nu]}
        if not (wanted & present):
            raise CompatibilityError("None of the requested groups exist in this model.")
        mask = 0
        for grp in present - wanted:
            mask |= 1 << grp
        self.model.opt.disableactuator = mask
        mj.mj_forward(self.model, self.data)
        if not self.enabled_actuator_mask().any():
            raise CompatibilityError("All actuators disabled by group selection.")

    def enabled_actuator_mask(self) -> np.ndarray:
        disabled = int(self.model.opt.disableactuator)
        return np.array([not ((disabled >> int(g)) & 1) for g in self.actuator_groups], dtype=bool).reshape(self.model.nu)


__all__ = ["ModelHandle"]
