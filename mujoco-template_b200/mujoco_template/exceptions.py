"""Error types of the boundary.

The class names and their hierarchy are the reference's (``mujoco_template/exceptions.py:4-21``) so that user code
catching ``TemplateError`` / ``ConfigError`` / ... keeps working; what this module adds is the other half of the
boundary: the mapping from the C-ABI's status codes (``include/b2mj.h``: ``B2_ERR_*``) onto those types, used by
``_capi.check``.
"""

from __future__ import annotations


class TemplateError(RuntimeError):
    """Root of everything this package raises; also what a CUDA failure (``B2_ERR_CUDA``) surfaces as."""

    status: int | None = -4


class ConfigError(TemplateError):
    """Bad arguments, malformed or oversized model blobs, MJCF features outside the compiled subset."""

    status = -1  # also -2 (blob) and -3 (capacity): see STATUS_TO_ERROR


class NameLookupError(TemplateError):
    """A body / joint / site / geom / actuator / keyframe name that the model does not contain."""

    status = None  # raised on the Python side only


class CompatibilityError(TemplateError):
    """Controller and model do not fit together (no actuators, every group disabled, ...)."""

    status = None


class LinearizationError(TemplateError):
    """``(A, B)`` could not be produced (``B2_ERR_LINEARIZE``)."""

    status = -5


# C status code -> exception type (0 is success and never reaches this table)
STATUS_TO_ERROR: dict[int, type[TemplateError]] = {
    -1: ConfigError,         # B2_ERR_ARG
    -2: ConfigError,         # B2_ERR_BLOB
    -3: ConfigError,         # B2_ERR_CAPACITY
    -4: TemplateError,       # B2_ERR_CUDA
    -5: LinearizationError,  # B2_ERR_LINEARIZE
}


def error_for_status(status: int, message: str) -> TemplateError:
    """The exception instance a non-zero C status code stands for."""
    return STATUS_TO_ERROR.get(int(status), TemplateError)(message)


__all__ = ["TemplateError", "NameLookupError", "CompatibilityError", "LinearizationError", "ConfigError",
           "STATUS_TO_ERROR", "error_for_status"]
