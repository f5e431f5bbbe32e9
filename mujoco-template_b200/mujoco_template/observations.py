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
    """Compiles an ``ObservationSpec`` into a fixed plan -- a list of ``(key, getter)`` pairs resolved once against the
    model (names -> ids, which state slices, which position tables) -- and replays that plan on every call."""

    _STATE_FIELDS = (("include_qpos", "qpos"), ("include_qvel", "qvel"))

    def __init__(self, model: Any, spec: ObservationSpec):
        self.model, self.spec = model, spec
        self._warned: set[str] = set()
        kinds = mj.mjtObj
        self.site_ids = self._resolve(kinds.mjOBJ_SITE, spec.sites_pos)
        self.body_ids = self._resolve(kinds.mjOBJ_BODY, spec.bodies_pos)
        self.geom_ids = self._resolve(kinds.mjOBJ_GEOM, spec.geoms_pos)
        self.subtree_ids = self._resolve(kinds.mjOBJ_BODY, spec.subtree_com)
        self.extra_items = tuple((name, self._producer(name, p)) for name, p in spec.extras.items())
        self._plan: list[tuple[str, Callable[[Any], np.ndarray]]] = self._compile()

    # ------------------------------------------------------------------ plan construction
    def _resolve(self, objtype: int, names: Sequence[str]) -> tuple[int, ...]:
        ids = tuple(int(mj.mj_name2id(self.model, objtype, name)) for name in names)
        for name, idx in zip(names, ids):
            if idx < 0:
                raise NameLookupError(f"Name not found in model: {name}")
        return ids

    @staticmethod
    def _producer(name: str, p: Any) -> ObservationProducer:
        if isinstance(p, ObservationProducer):
            return p
        if not callable(p):
            raise TypeError(f"extras[{name!r}] must be callable or ObservationProducer")
        return ObservationProducer(p)

    def _view(self, arr: Any) -> np.ndarray:
        """State slice: a zero-copy view of the (device-mapped) buffer unless the spec asks for copies."""
        return np.array(arr, copy=True) if self.spec.copy else np.asarray(arr)

    def _warn_once(self, key: str, msg: str) -> None:
        if key not in self._warned:
            self._warned.add(key)
            warnings.warn(msg, RuntimeWarning)

    @staticmethod
    def _rows(table: Any, ids: tuple[int, ...]) -> np.ndarray:
        out = np.empty((len(ids), 3))
        for row, idx in enumerate(ids):
            out[row] = table[idx]
        return out

    def _compile(self) -> list[tuple[str, Callable[[Any], np.ndarray]]]:
        s, plan = self.spec, []
        for flag, name in self._STATE_FIELDS:
            if getattr(s, flag):
                plan.append((name, lambda d, name=name: self._view(getattr(d, name))))
        if s.include_act:
            plan.append(("act", self._act))
        if s.include_ctrl:
            plan.append(("ctrl", lambda d: self._view(d.ctrl)))
        if s.include_sensordata:
            plan.append(("sensordata", self._sensordata))
        if s.include_time:
            plan.append(("time", lambda d: np.array([d.time], dtype=float)))
        if self.site_ids:
            plan.append(("sites_pos", lambda d: self._rows(d.site_xpos, self.site_ids)))
        if self.body_ids:
            plan.append(("bodies_pos", self._bodies))
        if self.geom_ids:
            plan.append(("geoms_pos", lambda d: self._rows(d.geom_xpos, self.geom_ids)))
        if self.subtree_ids:
            plan.append(("subtree_com", self._subtree))
        return plan

    # ------------------------------------------------------------------ getters with a story
    def _act(self, data: Any) -> np.ndarray:
        if hasattr(data, "act"):
            return self._view(data.act)
        self._warn_once("act", "ObservationSpec requested activations but data.act is missing; returning an empty array instead.")
        return np.zeros(0, dtype=float)

    def _sensordata(self, data: Any) -> np.ndarray:
        if self.model.nsensordata:
            return self._view(data.sensordata)
        self._warn_once("sensordata", "ObservationSpec requested sensordata but model has none; returning an empty array instead.")
        return np.zeros(0, dtype=float)

    def _bodies(self, data: Any) -> np.ndarray:
        inertial = self.spec.bodies_inertial and hasattr(data, "xipos")
        return self._rows(data.xipos if inertial else data.xpos, self.body_ids)

    def _subtree(self, data: Any) -> np.ndarray:
        mj.mj_subtreeCoM(self.model, data)
        return self._rows(data.subtree_com, self.subtree_ids)

    # ------------------------------------------------------------------ per-step call
    def __call__(self, data: Any) -> Observation:
        out: ObservationDict = {key: get(data) for key, get in self._plan}
        for name, producer in self.extra_items:
            if name in out:
                raise ValueError(f"extras[{name!r}] duplicates an existing observation key")
            out[name] = producer.produce(self.model, data, self.spec.copy)
        if self.spec.as_dict:
            return out
        if not out:
            return np.zeros(0)
        return np.concatenate([out[key].ravel() for key in sorted(out)])  # flattened form: keys in sorted order


__all__ = ["ObservationSpec", "ObservationExtractor", "ObservationProducer"]
