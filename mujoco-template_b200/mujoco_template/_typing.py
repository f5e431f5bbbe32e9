"""Type vocabulary of the package.

The first block carries the reference's alias names (``mujoco_template/_typing.py``) because user annotations import
them; the second block names what the batched path adds: SoA device arrays and their views.
"""

from __future__ import annotations

from typing import Any, Protocol, Union, runtime_checkable

import numpy as np

# ---- single-environment side (NumPy views over device-mapped host memory)
ObservationArray = np.ndarray
ObservationDict = dict[str, np.ndarray]
Observation = Union[ObservationDict, ObservationArray]
JacobianDict = dict[str, np.ndarray]          # {"jacp": (3, nv)[, "jacr": (3, nv)]}
JacobiansDict = dict[str, JacobianDict]       # keyed by the request token ("site:tip", "bodycom:torso", ...)
InfoDict = dict[str, Any]
StateSnapshot = dict[str, Any]                # qpos / qvel / act / ctrl / time copies (state_utils)


# ---- batched side (CUDA tensors, env index fastest in memory)
@runtime_checkable
class DeviceArray(Protocol):
    """What the batched path needs from a state buffer: a shape and a raw device pointer (``torch.Tensor`` fits)."""

    shape: Any

    def data_ptr(self) -> int: ...


BatchedObservation = Union[dict[str, Any], Any]  # dict of (N, ...) tensors, or one flattened (N, dim) tensor
LinearizationPair = tuple[Any, Any]              # (A, B): (N, 2nv, 2nv), (N, 2nv, nu) views of env-fastest buffers

__all__ = ["ObservationDict", "ObservationArray", "Observation", "JacobianDict", "JacobiansDict", "InfoDict", "StateSnapshot",
           "DeviceArray", "BatchedObservation", "LinearizationPair"]
