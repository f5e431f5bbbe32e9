"""Step drivers and the CSV sink of the hot path.

In scope (component 10): ``TrajectoryLogger``, ``iterate_passive``,
``run_passive_headless`` and the ``StepHook`` alias (reference
``mujoco_template/runtime.py:21,559-685``).  The reference's run harness, viewer, video and
CLI-flag plumbing are orchestration glue outside the hot path and are not rebuilt here.
"""

from __future__ import annotations

import csv
from collections.abc import Callable, Iterable, Iterator, Sequence
from pathlib import Path
from typing import IO, Any

from .exceptions import ConfigError

StepHook = Callable[[Any], None]


class TrajectoryLogger:
    """Writes one CSV row per step; usable as a context manager and reusable across runs."""

    def __init__(self, path: str | Path | None, columns: Sequence[str], formatter: Callable[[Any], Sequence[Any]]) -> None:
        if not columns:
            raise ConfigError("TrajectoryLogger requires at least one column name.")
        if formatter is None:
            raise ConfigError("TrajectoryLogger requires a formatter callable.")
        self._path = None if path is None else Path(path)
        self._columns = tuple(columns)
        self._formatter = formatter
        self._file: IO[str] | None = None
        self._writer = None

    @property
    def columns(self) -> tuple[str, ...]:
        return self._columns

    @property
    def enabled(self) -> bool:
        return self._path is not None

    def __enter__(self) -> "TrajectoryLogger":
        if self._path is not None:
            self._path.parent.mkdir(parents=True, exist_ok=True)
            self._file = self._path.open("w", newline="", encoding="utf-8")
            self._writer = csv.writer(self._file)
            self._writer.writerow(self._columns)
        return self

    def __exit__(self, exc_type, exc, exc_tb) -> None:
        self.close()

    def close(self) -> None:
        if self._file is not None:
            self._file.close()
        self._file = None
        self._writer = None

    def log(self, result: Any) -> tuple[Any, ...]:
        row = tuple(self._formatter(result))
        if len(row) != len(self._columns):
            raise ConfigError(
                f"Formatter returned a row of unexpected length ({len(row)} received, expected {len(self._columns)}).")
        if self._writer is not None:
            self._writer.writerow(row)
        return row


def _normalize_hooks(hooks: StepHook | Iterable[StepHook] | None) -> tuple[StepHook, ...]:
    if hooks is None:
        return ()
    if callable(hooks):
        return (hooks,)
    out = tuple(hooks)
    if any(not callable(h) for h in out):
        raise ConfigError("All hooks must be callables accepting StepResult.")
    return out


def iterate_passive(env: Any, *, duration: float | None = None, max_steps: int | None = None,
                    hooks: StepHook | Iterable[StepHook] | None = None, return_obs: bool = True) -> Iterator[Any]:
    """Step ``env`` until ``max_steps`` steps ran or ``env.data.time >= duration``.

    At least one step is always taken; hooks run after every step, before the result is
    yielded.  Works for ``Env`` and ``BatchedEnv`` alike (both expose ``step`` and ``data.time``).
    """
    if duration is not None and duration < 0:
        raise ConfigError("duration must be >= 0 when provided.")
    if max_steps is not None and max_steps < 1:
        raise ConfigError("max_steps must be >= 1 when provided.")
    callbacks = _normalize_hooks(hooks)
    taken = 0
    while True:
        result = env.step(return_obs=return_obs)
        taken += 1
        for cb in callbacks:
            cb(result)
        yield result
        if max_steps is not None and taken >= max_steps:
            return
        if duration is not None and env.data.time >= duration:
            return


def run_passive_headless(env: Any, *, duration: float | None = None, max_steps: int | None = None,
                         hooks: StepHook | Iterable[StepHook] | None = None, return_obs: bool = True) -> int:
    """Run :func:`iterate_passive` to exhaustion; returns the number of steps executed."""
    return sum(1 for _ in iterate_passive(env, duration=duration, max_steps=max_steps, hooks=hooks, return_obs=return_obs))


__all__ = ["StepHook", "TrajectoryLogger", "iterate_passive", "run_passive_headless"]
