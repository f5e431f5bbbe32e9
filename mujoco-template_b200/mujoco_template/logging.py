"""State/control recorder producing the reference's CSV schema.

Column order (reference ``mujoco_template/logging.py:81-178``): ``time_s``; then for every joint
in id order its qpos columns followed by its qvel columns (free joints are labelled
``pos_x..quat_z`` / ``lin_x..ang_z``, hinge/slide carry no component suffix); then
``ctrl[<actuator>]`` per actuator (``ctrl[none]`` when ``nu == 0``); then probe columns.
"""

from __future__ import annotations

from collections.abc import Callable, Iterator, Sequence
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from . import _mj as mj
from .exceptions import ConfigError
from .runtime import TrajectoryLogger

_QPOS_LABELS = {
    mj.mjtJoint.mjJNT_FREE: ("pos_x", "pos_y", "pos_z", "quat_w", "quat_x", "quat_y", "quat_z"),
    mj.mjtJoint.mjJNT_BALL: ("quat_w", "quat_x", "quat_y", "quat_z"),
    mj.mjtJoint.mjJNT_SLIDE: ("",),
    mj.mjtJoint.mjJNT_HINGE: ("",),
}
_QVEL_LABELS = {
    mj.mjtJoint.mjJNT_FREE: ("lin_x", "lin_y", "lin_z", "ang_x", "ang_y", "ang_z"),
    mj.mjtJoint.mjJNT_BALL: ("ang_x", "ang_y", "ang_z"),
    mj.mjtJoint.mjJNT_SLIDE: ("",),
    mj.mjtJoint.mjJNT_HINGE: ("",),
}


@dataclass(frozen=True)
class DataProbe:
    """Extra CSV column: ``extractor(env, result)`` must return a scalar (or ``None`` for an empty cell)."""

    name: str
    extractor: Callable[[Any, Any], Any]


@dataclass(frozen=True)
class ArrayProbe:
    """Extra CSV column of the batched recorder backed by one row of a ``(dim, nenv)`` array of ``env.data``
    (``site_xpos``, ``subtree_com``, ``xpos``, ``sensordata``, ``qacc``, ...): gathered on the device with the state
    columns, no Python per step.  ``row`` indexes the flattened leading dimension (site 2, z: ``3 * 2 + 2``)."""

    name: str
    array: str
    row: int


def _labels(table: dict, joint_type: int, count: int) -> Sequence[str]:
    known = table.get(joint_type)
    if known is None or len(known) != count:
        return [f"c{i}" if count > 1 else "" for i in range(count)]
    return known


def build_schema(model: Any, probe_names: Sequence[str] = ()):
    """Reference CSV schema: (columns, qpos indices, qvel indices, ctrl column names)."""
    m = model
    columns = ["time_s"]
    qpos_idx: list[int] = []
    qvel_idx: list[int] = []
    for j in range(m.njnt):
        name = mj.mj_id2name(m, mj.mjtObj.mjOBJ_JOINT, j) or f"joint_{j}"
        jtype = int(m.jnt_type[j])
        q0 = int(m.jnt_qposadr[j])
        q1 = int(m.jnt_qposadr[j + 1]) if j + 1 < m.njnt else int(m.nq)
        v0 = int(m.jnt_dofadr[j])
        v1 = int(m.jnt_dofadr[j + 1]) if j + 1 < m.njnt else int(m.nv)
        for idx, comp in zip(range(q0, q1), _labels(_QPOS_LABELS, jtype, q1 - q0)):
            columns.append(f"qpos[{name}].{comp}" if comp else f"qpos[{name}]")
            qpos_idx.append(idx)
        for idx, comp in zip(range(v0, v1), _labels(_QVEL_LABELS, jtype, v1 - v0)):
            columns.append(f"qvel[{name}].{comp}" if comp else f"qvel[{name}]")
            qvel_idx.append(idx)
    if m.nu > 0:
        ctrl_cols = tuple(f"ctrl[{mj.mj_id2name(m, mj.mjtObj.mjOBJ_ACTUATOR, a) or f'actuator_{a}'}]" for a in range(m.nu))
    else:
        ctrl_cols = ("ctrl[none]",)
    columns.extend(ctrl_cols)
    columns.extend(probe_names)
    return tuple(columns), qpos_idx, qvel_idx, ctrl_cols


class StateControlRecorder:
    """StepHook that logs time, generalized coordinates, velocities, controls and probes."""

    def __init__(self, env: Any, *, log_path: str | Path | None = None, store_rows: bool = True,
                 probes: Sequence[DataProbe] = ()) -> None:
        self._env = env
        self._model = env.model
        self._store_rows = bool(store_rows)
        self._rows: list[tuple[object, ...]] = []
        self._probes = self._validate_probes(probes)
        self._columns, self._qpos_indices, self._qvel_indices, self._ctrl_columns = self._schema()
        self._has_actuators = self._model.nu > 0
        self._column_index_map = {name: i for i, name in enumerate(self._columns)}
        self._logger = TrajectoryLogger(log_path, self._columns, self._format_row)

    @staticmethod
    def _validate_probes(probes: Sequence[DataProbe]) -> tuple[DataProbe, ...]:
        seen: set[str] = set()
        for probe in probes:
            if not isinstance(probe, DataProbe):
                raise ConfigError("All probes must be instances of DataProbe.")
            if not probe.name:
                raise ConfigError("Probe names must be non-empty strings.")
            if probe.name in seen:
                raise ConfigError(f"Duplicate probe name detected: {probe.name}")
            if not callable(probe.extractor):
                raise ConfigError(f"Probe '{probe.name}' extractor must be callable.")
            seen.add(probe.name)
        return tuple(probes)

    def _schema(self):
        return build_schema(self._model, [p.name for p in self._probes])

    @property
    def columns(self) -> tuple[str, ...]:
        return self._columns

    @property
    def column_index(self) -> dict[str, int]:
        return dict(self._column_index_map)

    @property
    def rows(self) -> list[tuple[object, ...]]:
        return self._rows

    def __enter__(self) -> "StateControlRecorder":
        self._logger.__enter__()
        return self

    def __exit__(self, exc_type, exc, exc_tb) -> None:
        self._logger.__exit__(exc_type, exc, exc_tb)

    def close(self) -> None:
        self._logger.close()

    def _format_row(self, result: Any) -> tuple[object, ...]:
        data, m = self._env.data, self._model
        if len(self._qpos_indices) != m.nq:
            raise ConfigError("Internal recorder error: qpos index coverage mismatch.")
        if len(self._qvel_indices) != m.nv:
            raise ConfigError("Internal recorder error: qvel index coverage mismatch.")
        row: list[object] = [float(data.time)]
        row += [float(data.qpos[i]) for i in self._qpos_indices]
        row += [float(data.qvel[i]) for i in self._qvel_indices]
        if self._has_actuators:
            row += [float(data.ctrl[a]) for a in range(m.nu)]
        else:
            row.append("")
        for probe in self._probes:
            value = probe.extractor(self._env, result)
            if isinstance(value, np.ndarray):
                if value.size != 1:
                    raise ConfigError(f"Probe '{probe.name}' returned array with {value.size} elements; expected scalar.")
                value = float(value.item())
            elif isinstance(value, (list, tuple)):
                raise ConfigError(f"Probe '{probe.name}' returned a non-scalar sequence; expected scalar-compatible value.")
            elif value is None:
                value = ""
            row.append(value)
        return tuple(row)

    def __call__(self, result: Any) -> None:
        row = self._logger.log(result)
        if self._store_rows:
            self._rows.append(row)

    def as_dicts(self) -> Iterator[dict[str, object]]:
        for row in self._rows:
            yield {name: row[i] for name, i in self._column_index_map.items()}


class BatchedStateControlRecorder:
    """StepHook for ``BatchedEnv``: the reference's CSV rows for a selection of envs.

    Every step appends ``[time, qpos.., qvel.., ctrl.., probes..]`` of the selected envs to a device-side ring buffer
    with ONE launch of the library's gather kernel (``b2_recorder_record``: a column table of (array, row) pairs built
    once, no host synchronisation in the step loop); full chunks are copied to pinned host memory asynchronously and
    written out as CSV with the reference schema plus a leading ``env`` column.

    Probes come in two kinds.  ``ArrayProbe(name, array, row)`` names one row of a ``(dim, nenv)`` array of
    ``env.data`` -- e.g. ``ArrayProbe("imu_z_m", "site_xpos", 3 * site_id + 2)`` for what the reference's drone example
    logs through ``e.data.site_xpos[imu, 2]`` (``examples/drone/drone_common.py:52-57``) -- and is just another column
    of the gather.  A ``DataProbe`` with a callable is vectorised: ``extractor(env, result)`` must return one value per
    selected env (tensor or array); it is evaluated after the gather and written into its column."""

    def __init__(self, env: Any, *, log_path: str | Path | None = None, env_indices: Sequence[int] | None = None,
                 chunk_steps: int = 256, store_rows: bool = False, probes: Sequence[DataProbe] = ()) -> None:
        import torch

        self._env, self._model = env, env.model
        m = env.model
        self._array_probes = {k: p for k, p in enumerate(probes) if isinstance(p, ArrayProbe)}
        self._probes = StateControlRecorder._validate_probes([DataProbe(p.name, lambda e, r: None) if isinstance(p, ArrayProbe) else p
                                                              for p in probes])
        cols, self._qi, self._vi, _ = build_schema(m, [p.name for p in self._probes])
        self.columns = ("env",) + cols
        n = env.data.qpos.shape[1]
        sel = list(range(n)) if env_indices is None else [int(i) for i in env_indices]
        if not sel or min(sel) < 0 or max(sel) >= n:
            raise ConfigError("env_indices must be a non-empty list of valid env indices")
        dev, dt = env.data.qpos.device, env.data.qpos.dtype
        self._sel = torch.as_tensor(sel, device=dev, dtype=torch.long)
        self._sel_host = sel
        self._qi_t = torch.as_tensor(self._qi, device=dev, dtype=torch.long)
        self._vi_t = torch.as_tensor(self._vi, device=dev, dtype=torch.long)
        self._nrow = 1 + m.nq + m.nv + max(m.nu, 1) + len(self._probes)
        self._chunk = int(chunk_steps)
        if self._chunk < 1:
            raise ConfigError("chunk_steps must be >= 1")
        self._ring = torch.zeros((self._chunk, self._nrow, len(sel)), device=dev, dtype=dt)
        self._host = torch.zeros((self._chunk, self._nrow, len(sel)), dtype=dt).pin_memory() if dev.type == "cuda" else None
        self._fill = 0
        self._store_rows = bool(store_rows)
        self.rows: list[tuple[object, ...]] = []
        self._path = None if log_path is None else Path(log_path)
        self._file = None
        self._writer = None
        self.steps = 0
        for p in self._array_probes.values():
            arr = getattr(env.data, p.array, None)
            if arr is None or not 0 <= int(p.row) < arr.shape[0]:
                raise ConfigError(f"ArrayProbe {p.name!r}: env.data has no row {p.row} of array {p.array!r}")
        # device path: the column table of the library's gather kernel (one launch per recorded step)
        self._native = None
        self._needs_derived = any(p.array not in ("qpos", "qvel", "ctrl", "qacc_warmstart") for p in self._array_probes.values())
        if dev.type == "cuda":
            from . import _capi

            d = env.data
            table: list[tuple[int | None, int, int]] = [(None, 0, 1)]  # time_s
            table += [(d.qpos.data_ptr(), int(r), 0) for r in self._qi]
            table += [(d.qvel.data_ptr(), int(r), 0) for r in self._vi]
            table += [(d.ctrl.data_ptr(), a, 0) for a in range(m.nu)] if m.nu else [(None, 0, 2)]
            for k in range(len(self._probes)):
                p = self._array_probes.get(k)
                table.append((getattr(d, p.array).data_ptr(), int(p.row), 0) if p is not None else (None, 0, 2))
            assert len(table) == self._nrow
            self._native = _capi.NativeRecorder(d.backend.batch, table, sel)

    def __enter__(self) -> "BatchedStateControlRecorder":
        if self._path is not None:
            import csv

            self._path.parent.mkdir(parents=True, exist_ok=True)
            self._file = self._path.open("w", newline="", encoding="utf-8")
            self._writer = csv.writer(self._file)
            self._writer.writerow(self.columns)
        return self

    def __exit__(self, exc_type, exc, exc_tb) -> None:
        self.close()

    def __call__(self, result: Any) -> None:
        import torch

        data, m = self._env.data, self._model
        slot = self._ring[self._fill]
        if self._native is not None:
            if self._needs_derived:
                data.backend.ensure_derived()  # lazily stepped envs: produce xpos / site_xpos / ... of this step first
            stream = torch.cuda.current_stream(slot.device).cuda_stream
            self._native.record(float(data.time), slot.data_ptr(), stream)
            base = 1 + m.nq + m.nv + max(m.nu, 1)
            for k, probe in enumerate(self._probes):
                if k not in self._array_probes:
                    slot[base + k] = torch.as_tensor(probe.extractor(self._env, result), device=slot.device, dtype=slot.dtype)
            self._fill += 1
            self.steps += 1
            if self._fill == self._chunk:
                self.flush()
            return
        slot[0] = float(data.time)
        slot[1: 1 + m.nq] = data.qpos[self._qi_t][:, self._sel]
        slot[1 + m.nq: 1 + m.nq + m.nv] = data.qvel[self._vi_t][:, self._sel]
        base = 1 + m.nq + m.nv
        if m.nu:
            slot[base: base + m.nu] = data.ctrl[:, self._sel]
        else:
            slot[base] = float("nan")
        base += max(m.nu, 1)
        for k, probe in enumerate(self._probes):
            if k in self._array_probes:
                p = self._array_probes[k]
                slot[base + k] = getattr(data, p.array)[int(p.row)][self._sel]
            else:
                slot[base + k] = torch.as_tensor(probe.extractor(self._env, result), device=slot.device, dtype=slot.dtype)
        self._fill += 1
        self.steps += 1
        if self._fill == self._chunk:
            self.flush()

    def flush(self) -> None:
        """Copy the filled part of the ring to the host and emit its rows (env-major within each step)."""
        if self._fill == 0:
            return
        k = self._fill
        self._fill = 0
        if self._host is not None:
            self._host[:k].copy_(self._ring[:k], non_blocking=True)
            import torch

            torch.cuda.current_stream(self._ring.device).synchronize()
            block = self._host[:k].numpy()
        else:
            block = self._ring[:k].numpy()
        m = self._model
        nu1 = max(m.nu, 1)
        for t in range(k):
            for c, e in enumerate(self._sel_host):
                vals = block[t, :, c]
                row: list[object] = [e] + [float(x) for x in vals[: 1 + m.nq + m.nv]]
                row += [float(x) for x in vals[1 + m.nq + m.nv: 1 + m.nq + m.nv + nu1]] if m.nu else [""]
                row += [float(x) for x in vals[1 + m.nq + m.nv + nu1:]]
                tup = tuple(row)
                if self._writer is not None:
                    self._writer.writerow(tup)
                if self._store_rows:
                    self.rows.append(tup)

    def close(self) -> None:
        self.flush()
        if self._file is not None:
            self._file.close()
        self._file = None
        self._writer = None


__all__ = ["DataProbe", "ArrayProbe", "StateControlRecorder", "BatchedStateControlRecorder", "build_schema"]
