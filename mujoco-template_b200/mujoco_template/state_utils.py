"""Snapshot / restore of the integration state (reference ``mujoco_template/state_utils.py:9-31``).

Deliberately mirrors the reference's choice of fields: qpos, qvel, act, ctrl, time -- and
NOT ``qacc_warmstart`` (the Python FD fallback therefore warm-starts from whatever the
previous rollout left behind, exactly like the reference).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from ._typing import StateSnapshot


def _snapshot_state(data: Any) -> StateSnapshot:
    snap: StateSnapshot = {name: np.array(getattr(data, name)) for name in ("qpos", "qvel", "ctrl")}
    snap["act"] = np.array(data.act) if hasattr(data, "act") else None
    snap["time"] = float(data.time)
    return snap


def _restore_state(data: Any, snap: StateSnapshot) -> None:
    for name in ("qpos", "qvel", "ctrl"):
        getattr(data, name)[:] = snap[name]
    if snap.get("act") is not None and hasattr(data, "act"):
        data.act[:] = snap["act"]
    t = snap.get("time")
    data.time = 0.0 if t is None else float(np.asarray(t).reshape(-1)[0])


__all__ = ["_snapshot_state", "_restore_state"]
