"""Command-line smoke run on the B200 path.

Same flags and output as the reference's ``python -m mujoco_template`` (``mujoco_template/__main__.py:10-33``):

    python -m mujoco_template examples/pendulum/pendulum.xml --steps 300 --zero

BASELINE config #1 is exactly this command on the pendulum model.
"""

from __future__ import annotations

import argparse
import sys

from .controllers import ZeroController
from .env import Env

_FLAGS = (
    # name, kwargs
    ("--steps", dict(type=int, default=300)),
    ("--zero", dict(action="store_true", help="Use ZeroController")),
    ("--groups", dict(type=int, nargs="*", default=None, help="Enable only these actuator groups")),
    ("--decim", dict(type=int, default=1, help="Control decimation (>=1)")),
)


def _parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="python -m mujoco_template",
                                     description="MuJoCo template smoke test on the B200 path (fail-fast)")
    parser.add_argument("xml", help="Path to MJCF XML")
    for flag, kwargs in _FLAGS:
        parser.add_argument(flag, **kwargs)
    return parser


def main(argv: list[str] | None = None) -> int:
    """Build the env the flags describe, run ``--steps`` passive steps, report how many were taken."""
    opts = _parser().parse_args(argv)
    controller = ZeroController() if opts.zero else None
    env = Env.from_xml_path(opts.xml, controller=controller, enabled_groups=opts.groups, control_decimation=opts.decim)
    taken = 0
    for _ in env.passive(max_steps=opts.steps):
        taken += 1
    print(f"Completed {taken} steps.")
    return taken


if __name__ == "__main__":
    main(sys.argv[1:])
