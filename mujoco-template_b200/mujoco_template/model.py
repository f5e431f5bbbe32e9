"""``ModelHandle``: one compiled model plus one data block, with the hot-path calls routed to the GPU library.

Public surface of reference ``mujoco_template/model.py:11-105`` (loaders, ``forward`` / ``step`` / ``reset`` /
``reset_keyframe``, actuator-group masks).  ``step`` / ``forward`` / ``reset`` end in ``libb2mj.so`` kernels
(``b2_step`` / ``b2_forward``) rather than in ``mj_step`` / ``mj_forward`` / ``mj_resetData``; the "binary" model file is
this package's own compiled-model blob (``MjModel.save_compiled``), not MuJoCo's MJB.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from . import _mj as mj
from .exceptions import CompatibilityError, ConfigError, NameLookupError, TemplateError

_GROUP_BITS = 32  # width of opt.disableactuator


def _keyframe_index(model: "mj.MjModel", key: int | str) -> int:
    """Resolve a keyframe given by name or by index; errors follow the reference (name -> NameLookupError)."""
    if isinstance(key, str):
        found = mj.mj_name2id(model, mj.mjtObj.mjOBJ_KEY, key)
        if found < 0:
            raise NameLookupError(f"Keyframe name not found: {key}")
        return found
    index = int(key)
    if not 0 <= index < model.nkey:
        raise ConfigError(f"Keyframe index out of range: {index}")
    return index


class ModelHandle:
    """Owns ``(model, data)``; every method below is a thin, fail-fast forwarder."""

    def __init__(self, model: "mj.MjModel", data: "mj.MjData | None" = None):
        if data is not None and data.model is not model:
            raise ConfigError("Provided mj.MjData must reference the supplied model.")
        self.model = model
        self.data = mj.MjData(model) if data is None else data

    # ------------------------------------------------------------------ construction
    @classmethod
    def _from(cls, loader, source: str) -> "ModelHandle":
        return cls(loader(source))

    @classmethod
    def from_xml_path(cls, xml_path: str) -> "ModelHandle":
        return cls._from(mj.MjModel.from_xml_path, xml_path)

    @classmethod
    def from_xml_string(cls, xml_text: str) -> "ModelHandle":
        return cls._from(mj.MjModel.from_xml_string, xml_text)

    @classmethod
    def from_binary_path(cls, mjb_path: str) -> "ModelHandle":
        """Load what :meth:`save_binary` wrote (a compiled-model blob of this package)."""
        return cls._from(mj.MjModel.from_compiled, mjb_path)

    @classmethod
    def from_model_and_data(cls, model: "mj.MjModel", data: "mj.MjData") -> "ModelHandle":
        """Adopt a pair that already exists: nothing is allocated."""
        return cls(model, data)

    def save_binary(self, mjb_path: str) -> None:
        try:
            self.model.save_compiled(mjb_path)
        except OSError as exc:
            raise TemplateError(f"mj_saveModel failed for {mjb_path}: {exc}") from exc

    # ------------------------------------------------------------------ hot path (GPU)
    def step(self) -> None:
        mj.mj_step(self.model, self.data)

    def forward(self) -> None:
        mj.mj_forward(self.model, self.data)

    def reset(self) -> None:
        mj.mj_resetData(self.model, self.data)

    def reset_keyframe(self, key: int | str) -> None:
        mj.mj_resetDataKeyframe(self.model, self.data, _keyframe_index(self.model, key))

    # ------------------------------------------------------------------ actuator groups
    # The mask lives in model.opt.disableactuator and is honoured by the actuation stage of the kernels.
    @property
    def actuator_groups(self) -> np.ndarray:
        return np.asarray(self.model.actuator_group, dtype=int).copy()

    def enabled_actuator_mask(self) -> np.ndarray:
        """Boolean (nu,) array: True where the actuator's group is not disabled."""
        off_bits = int(self.model.opt.disableactuator)
        groups = self.actuator_groups[: self.model.nu]
        return ((off_bits >> groups) & 1) == 0 if groups.size else np.zeros(0, dtype=bool)

    def set_enabled_actuator_groups(self, enabled_groups: Iterable[int]) -> None:
        keep = sorted({int(g) for g in enabled_groups})
        if not keep:
            raise CompatibilityError("At least one actuator group must be enabled.")
        if keep[0] < 0 or keep[-1] >= _GROUP_BITS:
            raise ConfigError("Actuator groups must be in [0, 31].")
        if self.model.nu == 0:
            raise CompatibilityError("Model has no actuators (nu=0).")
        in_model = set(self.actuator_groups[: self.model.nu].tolist())
        if in_model.isdisjoint(keep):
            raise CompatibilityError("None of the requested groups exist in this model.")
        self.model.opt.disableactuator = sum(1 << g for g in in_model.difference(keep))
        self.forward()
        if not self.enabled_actuator_mask().any():
            raise CompatibilityError("All actuators disabled by group selection.")


__all__ = ["ModelHandle"]
