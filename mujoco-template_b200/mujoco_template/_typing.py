"""Type aliases shared across the package (names follow reference ``mujoco_template/_typing.py``)."""

from __future__ import annotations

from typing import Any, Union

import numpy as np

ObservationDict = dict[str, np.ndarray]
ObservationArray = np.ndarray
Observation = Union[ObservationDict, ObservationArray]
JacobianDict = dict[str, np.ndarray]
JacobiansDict = dict[str, JacobianDict]
InfoDict = dict[str, Any]
StateSnapshot = dict[str, Any]

__all__ = ["ObservationDict", "ObservationArray", "Observation", "JacobianDict", "JacobiansDict", "InfoDict", "StateSnapshot"]
