"""Exception hierarchy of the boundary (mirrors reference mujoco_template/exceptions.py:4-21)."""

from __future__ import annotations


class TemplateError(RuntimeError):
    """Base class for every error raised by this package."""


class NameLookupError(TemplateError):
    """A named body/joint/site/geom/keyframe does not exist in the model."""


class CompatibilityError(TemplateError):
    """Controller and model cannot work together."""


class LinearizationError(TemplateError):
    """The (A, B) linearization could not be produced."""


class ConfigError(TemplateError):
    """Invalid configuration, arguments, or unsupported MJCF feature."""


__all__ = ["TemplateError", "NameLookupError", "CompatibilityError", "LinearizationError", "ConfigError"]
