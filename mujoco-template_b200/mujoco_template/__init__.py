"""B200-native drop-in for the hot path of ``mujoco_template``.

Same public names as the reference package for everything on the path
(reference ``mujoco_template/__init__.py:81-135``), plus ``BatchedEnv``.  ``mj`` is the
MuJoCo-shaped facade over ``libb2mj.so`` (``_mj.py``), since the real ``mujoco`` module is
neither required nor used.  Out-of-scope reference components (viewer, video, adaptive
camera, run harness) are intentionally absent; see DESIGN.md.
"""

from __future__ import annotations

from . import _mj as mj
from ._typing import InfoDict, JacobianDict, JacobiansDict, Observation, ObservationArray, ObservationDict, StateSnapshot
from .batched import BatchedEnv, BatchedObservationExtractor, shard_range
from .compat import CompatibilityReport, check_controller_compat
from .control import ControlSpace, Controller, ControllerCapabilities
from .controllers import PositionTargetDemo, ZeroController
from .env import Env, StepResult
from .exceptions import CompatibilityError, ConfigError, LinearizationError, NameLookupError, TemplateError
from .jacobians import compute_requested_jacobians
from .linearization import linearize_discrete
from .logging import ArrayProbe, BatchedStateControlRecorder, DataProbe, StateControlRecorder
from .model import ModelHandle
from .observations import ObservationExtractor, ObservationProducer, ObservationSpec
from .runtime import StepHook, TrajectoryLogger, iterate_passive, run_passive_headless
from .setpoints import batched_steady_ctrl0, steady_ctrl0

__version__ = "0.1.0"

__all__ = [
    "TemplateError", "NameLookupError", "CompatibilityError", "LinearizationError", "ConfigError",
    "ControlSpace", "Controller", "ControllerCapabilities", "ObservationSpec", "ObservationExtractor",
    "ObservationProducer", "ModelHandle", "CompatibilityReport", "StepResult", "Env", "BatchedEnv",
    "BatchedObservationExtractor", "shard_range", "ZeroController", "PositionTargetDemo",
    "check_controller_compat", "linearize_discrete", "compute_requested_jacobians", "steady_ctrl0",
    "batched_steady_ctrl0", "TrajectoryLogger",
    "DataProbe", "ArrayProbe", "StateControlRecorder", "BatchedStateControlRecorder", "StepHook", "iterate_passive", "run_passive_headless",
    "ObservationDict", "ObservationArray", "Observation", "JacobianDict", "JacobiansDict", "InfoDict",
    "StateSnapshot", "mj", "__version__",
]
