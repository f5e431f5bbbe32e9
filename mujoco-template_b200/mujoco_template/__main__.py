"""CLI smoke run (reference ``mujoco_template/__main__.py:10-33``):

    python -m mujoco_template examples/pendulum/pendulum.xml --steps 300 --zero
"""

from __future__ import annotations

import argparse

from .controllers import ZeroController
from .env import Env


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description="MuJoCo template smoke test on the B200 path (fail-fast)")
    ap.add_argument("xml", help="Path to MJCF XML")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--zero", action="store_true", help="Use ZeroController")
    ap.add_argument("--groups", type=int, nargs="*", default=None, help="Enable only these actuator groups")
    ap.add_argument("--decim", type=int, default=1, help="Control decimation (>=1)")
    args = ap.parse_args(argv)
    env = Env.from_xml_path(args.xml, controller=ZeroController() if args.zero else None,
                            enabled_groups=args.groups, control_decimation=args.decim)
    steps = sum(1 for _ in env.passive(max_steps=args.steps))
    print(f"Completed {steps} steps.")
    return steps


if __name__ == "__main__":
    main()
