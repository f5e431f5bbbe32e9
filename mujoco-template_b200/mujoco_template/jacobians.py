"""Requested point Jacobians (reference ``mujoco_template/jacobians.py:12-83``).

Tokens: ``site:<name>``, ``body:<name>`` -> ``jacp`` and ``jacr``; ``bodycom:<name>``,
``subtreecom:<name>`` -> ``jacp`` only.  A bare ``com`` is rejected as ambiguous.
"""

from __future__ import annotations

from collections.abc import Iterable
from typing import Any

import numpy as np

from . import _mj as mj
from ._typing import JacobiansDict
from .exceptions import ConfigError, NameLookupError

_PREFIXES = ("site", "body", "bodycom", "subtreecom")


def _parse_jacobian_token(token: str) -> tuple[str, str | None]:
    if token == "com":
        return ("com", None)
    head, sep, tail = token.partition(":")
    if sep and head in _PREFIXES:
        return (head, tail)
    raise ConfigError(f"Unknown jacobian token: {token}")


def resolve_jacobian_token(model: Any, token: str) -> tuple[str, int]:
    """Token -> (kind, object id); shared by the single-env and batched paths."""
    kind, name = _parse_jacobian_token(token)
    if kind == "com":
        raise ConfigError("'com' jacobian is ambiguous; request 'bodycom:<name>' or 'subtreecom:<name>'.")
    if kind == "site":
        idx = mj.mj_name2id(model, mj.mjtObj.mjOBJ_SITE, name)
        if idx < 0:
            raise NameLookupError(f"Site not found: {name}")
    else:
        idx = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, name)
        if idx < 0:
            raise NameLookupError(f"Body not found: {name}")
    return kind, idx


def compute_requested_jacobians(model: Any, data: Any, tokens: Iterable[str]) -> JacobiansDict:
    out: JacobiansDict = {}
    for token in tokens:
        kind, idx = resolve_jacobian_token(model, token)
        jacp = np.zeros((3, model.nv))
        if kind in ("site", "body"):
            jacr = np.zeros((3, model.nv))
            (mj.mj_jacSite if kind == "site" else mj.mj_jacBody)(model, data, jacp, jacr, idx)
            out[token] = {"jacp": jacp, "jacr": jacr}
        elif kind == "bodycom":
            mj.mj_jacBodyCom(model, data, jacp, None, idx)
            out[token] = {"jacp": jacp}
        else:
            mj.mj_jacSubtreeCom(model, data, jacp, idx)
            out[token] = {"jacp": jacp}
    return out


__all__ = ["compute_requested_jacobians", "resolve_jacobian_token"]
