"""BASELINE config #1 (pendulum, ZeroController, 300 steps from reset, N = 1 Env in device-mapped host memory): wall time of
the reference-style Python loop on the GPU path and on the CPU oracle backend."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mujoco-template_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mujoco_template as mt
from conftest import load_model

model = load_model("pendulum")
env = mt.Env(mt.ModelHandle(model), obs_spec=mt.ObservationSpec(include_time=True), controller=mt.ZeroController())
for rep in range(3):
    env.reset()
    t0 = time.perf_counter()
    steps = sum(1 for _ in env.passive(max_steps=300))
    dt = time.perf_counter() - t0
    print(f"gpu path: {steps} steps in {dt * 1e3:.2f} ms = {dt / steps * 1e6:.1f} us/step")
from oracle.oracle import OracleData, OracleModel
om = OracleModel(model.blob, dict(nq=1, nv=1, nu=1, nbody=model.nbody, njnt=1, ngeom=model.ngeom, nsite=model.nsite, ntendon=0))
od = OracleData(om)
t0 = time.perf_counter()
for _ in range(300):
    od.ctrl[:] = 0.0
    od.step()
dt = time.perf_counter() - t0
print(f"oracle (C restatement, bare ctypes loop): 300 steps in {dt * 1e3:.2f} ms = {dt / 300 * 1e6:.1f} us/step")
