"""Generate tests/golden/traj_<model>.npz from the CPU oracle (regression pins, not MuJoCo goldens:
mujoco is not importable in the build container -- see DESIGN.md "Oracle")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import MODEL_NAMES, load_model, oracle_for, random_states  # noqa: E402

if __name__ == "__main__":
    for name in MODEL_NAMES:
        model = load_model(name)
        qpos, qvel, ctrl = random_states(model, name, 1, seed=2024)
        if name == "drone":
            ctrl[:] = 3.3
        om, d = oracle_for(model)
        d.qpos[:] = qpos[0]; d.qvel[:] = qvel[0]; d.ctrl[:] = ctrl[0]
        A, B = d.transition_fd(1e-6, True)
        T = 60
        qt, vt = [], []
        for _ in range(T):
            d.step()
            qt.append(np.array(d.qpos)); vt.append(np.array(d.qvel))
        np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), qpos0=qpos[0], qvel0=qvel[0], ctrl=ctrl[0], A=A, B=B,
                            qpos_traj=np.array(qt), qvel_traj=np.array(vt))
        print(name, "ok", "ncon", d.ncon)
