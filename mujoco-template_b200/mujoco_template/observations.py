"""Declarative observations (reference ``mujoco_template/observations.py:34-174``).

Contract kept: dict keys ``qpos qvel act ctrl sensordata time sites_pos bodies_pos geoms_pos
subtree_com`` + user extras; the flattened form concatenates ``sorted(keys)``; state slices
are zero-copy views unless ``copy=True``; position groups are ``(k, 3)`` copies.
Derived positions are those of the forward pass preceding the last integration (they lag
``qpos`` by one step, as in MuJoCo).
"""

from __future__ import annotations

import warnings
from collections.abc import Callable, Mapping, Sequence
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _mj as mj
from ._typing import Observation, ObservationDict
from .exceptions import NameLookupError

_Fn = Callable[[Any, Any], "np.ndarray | Sequence[float]"]


@dataclass(frozen=True)
class ObservationProducer:
    """User callback producing one observation entry."""

    fn: _Fn
    copy: bool | None = None

    def produce(self, model: Any, data: Any, default_copy: bool) -> np.ndarray:
        value = self.fn(model, data)
        want_copy = default_copy if self.copy is None else bool(self.copy)
        arr = value if isinstance(value, np.ndarray) else np.asarray(value)
        return np.array(arr, copy=True) if want_copy else arr


@dataclass
class ObservationSpec:
    include_qpos: bool = True
    include_qvel: bool = True
    include_act: bool = False
    include_ctrl: bool = False
    include_sensordata: bool = False
    include_time: bool = False
    sites_pos: Sequence[str] = field(default_factory=tuple)
    bodies_pos: Sequence[str] = field(default_factory=tuple)
    geoms_pos: Sequence[str] = field(default_factory=tuple)
    subtree_com: Sequence[str] = field(default_factory=tuple)
    as_dict: bool = True
    bodies_inertial: bool = False
    extras: Mapping[str, "ObservationProducer | _Fn"] = field(default_factory=dict)
    copy: bool = False


class ObservationExtractor:
    def __init__(self, model: Any, spec: ObservationSpec):
        self.model = model
        self.spec = spec
        O = mj.mjtObj
        self.site_ids = self._ids(O.mjOBJ_SITE, spec.sites_pos)
        self.body_ids = self._ids(O.mjOBJ_BODY, spec.bodies_pos)
        self.geom_ids = self._ids(O.mjOBJ_GEOM, spec.geoms_pos)
        self.subtree_ids = self._ids(O.mjOBJ_BODY, spec.subtree_com)
        self.extra_items = tuple((name, self._as_producer(name, p)) for name, p in spec.extras.items())
        self._warned: set[str] = set()

    def _ids(self, objtype: int, names: Sequence[str]) -> tuple[int, ...]:
        out = []
        for name in names:
            idx = int(mj.mj_name2id(self.model, objtype, name))
            if idx < 0:
                raise NameLookupError(f"Name not found in model: {name}")
            out.append(idx)
        return tuple(out)

    @staticmethod
    def _as_producer(name: str, p: Any) -> ObservationProducer:
        if isinstance(p, ObservationProducer):
            return p
        if callable(p):
            return ObservationProducer(p)
        raise TypeError(f"extras[{name!r}] must be callable or ObservationProducer")

    def _warn_once(self, key: str, msg: str) -> None:
        if key not in self._warned:
            self._warned.add(key)
            warnings.warn(msg, RuntimeWarning)

    def _slice(self, arr: Any) -> np.ndarray:
        return np.array(arr, copy=True) if self.spec.copy else np.asarray(arr)

    @staticmethod
    def _gather(src: Any, ids: tuple[int, ...]) -> np.ndarray:
        pos = np.zeros((len(ids), 3))
        for row, idx in enumerate(ids):
            pos[row] = src[idx]
        return pos

    def __call__(self, data: Any) -> Observation:
        s = self.spec
        out: ObservationDict = {}
        if s.include_qpos:
            out["qpos"] = self._slice(data.qpos)
        if s.include_qvel:
            out["qvel"] = self._slice(data.qvel)
        if s.include_act:
            if hasattr(data, "act"):
                out["act"] = self._slice(data.act)
            else:
                self._warn_once("act", "ObservationSpec requested activations but data.act is missing; returning an empty array instead.")
                out["act"] = np.zeros(0, dtype=float)
        if s.include_ctrl:
            out["ctrl"] = self._slice(data.ctrl)
        if s.include_sensordata:
            if self.model.nsensordata == 0:
                self._warn_once("sensordata", "ObservationSpec requested sensordata but model has none; returning an empty array instead.")
                out["sensordata"] = np.zeros(0, dtype=float)
            else:
                out["sensordata"] = self._slice(data.sensordata)
        if s.include_time:
            out["time"] = np.array([data.time], dtype=float)
        if self.site_ids:
            out["sites_pos"] = self._gather(data.site_xpos, self.site_ids)
        if self.body_ids:
            src = data.xipos if (s.bodies_inertial and hasattr(data, "xipos")) else data.xpos
            out["bodies_pos"] = self._gather(src, self.body_ids)
        if self.geom_ids:
            out["geoms_pos"] = self._gather(data.geom_xpos, self.geom_ids)
        if self.subtree_ids:
            mj.mj_subtreeCoM(self.model, data)
            out["subtree_com"] = self._gather(data.subtree_com, self.subtree_ids)
        for name, producer in self.extra_items:
            if name in out:
                raise ValueError(f"extras[{name!r}] duplicates an existing observation key")
            out[name] = producer.produce(self.model, data, s.copy)
        if s.as_dict:
            return out
        parts = [out[k].ravel() for k in sorted(out)]
        return np.concatenate(parts) if parts else np.zeros(0)


__all__ = ["ObservationSpec", "ObservationExtractor", "ObservationProducer"]
