"""Pins for the CPU oracle (parity against a live MuJoCo is unpinned: see DESIGN.md).

Known answers and invariants the oracle must satisfy independently of MuJoCo:
closed-form pendulum under classical RK4, the 2-DoF cart-pole equations under semi-implicit
Euler with implicit joint damping, the drone hover thrust, conservation laws, (A, B) structure,
and committed regression trajectories (tests/golden/traj_*.npz, made by make_golden.py)."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, MODEL_NAMES, load_model, oracle_for, random_states


def test_pendulum_closed_form_rk4():
    model = load_model("pendulum")
    om, d = oracle_for(model)
    mass, lc, g, h = float(model.body_mass[1]), 0.25, 9.81, 0.005
    I = float(model.body_inertia[1][0]) + mass * lc * lc  # about the hinge
    tau = 0.7

    def f(th, tq):
        return (-mass * g * lc * math.sin(th) + max(-3.0, min(3.0, tq))) / I

    th, w = 1.2, 0.3
    d.qpos[0], d.qvel[0], d.ctrl[0] = th, w, tau
    for _ in range(400):
        d.step()
        k1 = (w, f(th, tau)); k2 = (w + h / 2 * k1[1], f(th + h / 2 * k1[0], tau))
        k3 = (w + h / 2 * k2[1], f(th + h / 2 * k2[0], tau)); k4 = (w + h * k3[1], f(th + h * k3[0], tau))
        th += h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0]); w += h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
    assert abs(d.qpos[0] - th) < 1e-12 and abs(d.qvel[0] - w) < 1e-12
    # force clamp: |tau| > 3 saturates (pendulum.xml forcerange)
    d.reset(); d.ctrl[0] = 10.0; d.forward()
    assert abs(d.qacc[0] - 3.0 / I) < 1e-12


def test_cartpole_manipulator_equations_implicit_damping():
    model = load_model("cartpole")
    om, d = oracle_for(model)
    mc, mp = float(model.body_mass[1]), float(model.body_mass[2])
    lc = 0.3                                   # pole CoM above the hinge
    Ip = float(model.body_inertia[2][0])       # about the pole CoM, transverse axis
    g, h, gear = 9.81, 0.01, 50.0
    Dmp = np.diag([1.0, 0.1])
    rng = np.random.default_rng(0)
    for _ in range(20):
        x, th, xd, thd, u = rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-2, 2), rng.uniform(-3, 3)
        d.reset(); d.qpos[:] = [x, th]; d.qvel[:] = [xd, thd]; d.ctrl[0] = u
        # hinge about +y, pole along +z at th=0: CoM at (x + lc sin th, lc cos th)
        M = np.array([[mc + mp, mp * lc * math.cos(th)], [mp * lc * math.cos(th), Ip + mp * lc * lc]])
        bias = np.array([-mp * lc * math.sin(th) * thd * thd, -mp * g * lc * math.sin(th)])
        frc = np.array([gear * u, 0.0]) - Dmp @ np.array([xd, thd]) - bias
        qacc = np.linalg.solve(M, frc)
        d.forward()
        assert np.allclose(d.qacc, qacc, rtol=1e-11, atol=1e-11)
        v_new = np.array([xd, thd]) + h * np.linalg.solve(M + h * Dmp, frc)
        d.step()
        assert np.allclose(d.qvel, v_new, rtol=1e-11, atol=1e-12)
        assert np.allclose(d.qpos, np.array([x, th]) + h * v_new, rtol=1e-12, atol=1e-13)


def test_drone_hover_and_free_fall():
    model = load_model("drone")
    om, d = oracle_for(model)
    total_mass = 4 * 0.25 + 0.325
    assert abs(model.body_mass.sum() - total_mass) < 1e-12
    d.reset(0)
    assert np.allclose(d.ctrl, total_mass * 9.81 / 4)
    d.forward()
    assert np.max(np.abs(d.qacc)) < 1e-10 and d.ncon == 0
    d.ctrl[:] = 0; d.qpos[2] = 5.0; d.forward()
    assert abs(d.qacc[2] + 9.81) < 1e-9 and np.max(np.abs(d.qacc[[0, 1, 3, 4, 5]])) < 1e-9
    for _ in range(100):
        d.step()
        assert abs(np.linalg.norm(d.qpos[3:7]) - 1.0) < 1e-12


def test_pendulum_energy_drift_is_rk4_small():
    model = load_model("pendulum")
    om, d = oracle_for(model)
    mass, lc, g = float(model.body_mass[1]), 0.25, 9.81
    I = float(model.body_inertia[1][0]) + mass * lc * lc
    d.qpos[0] = math.pi / 2
    E0 = -mass * g * lc * math.cos(d.qpos[0])
    for _ in range(600):
        d.step()
    E = 0.5 * I * d.qvel[0] ** 2 - mass * g * lc * math.cos(d.qpos[0])
    assert abs(E - E0) < 1e-8


def test_linearization_structure_and_analytic_pendulum():
    model = load_model("pendulum")
    om, d = oracle_for(model)
    mass, lc, g, h = float(model.body_mass[1]), 0.25, 9.81, 0.005
    I = float(model.body_inertia[1][0]) + mass * lc * lc

    def rk4_map(x, u):
        def f(s):
            return np.array([s[1], (-mass * g * lc * math.sin(s[0]) + u) / I])
        k1 = f(x); k2 = f(x + h / 2 * k1); k3 = f(x + h / 2 * k2); k4 = f(x + h * k3)
        return x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)

    x0, u0 = np.array([0.8, -0.4]), 0.5
    d.qpos[0], d.qvel[0], d.ctrl[0] = x0[0], x0[1], u0
    A, B = d.transition_fd(1e-6, True)
    e = 1e-6
    Aref = np.stack([(rk4_map(x0 + e * np.eye(2)[i], u0) - rk4_map(x0 - e * np.eye(2)[i], u0)) / (2 * e) for i in range(2)], axis=1)
    Bref = ((rk4_map(x0, u0 + e) - rk4_map(x0, u0 - e)) / (2 * e))[:, None]
    assert np.allclose(A, Aref, rtol=1e-6, atol=1e-9) and np.allclose(B, Bref, rtol=1e-6, atol=1e-9)
    assert d.qpos[0] == x0[0] and d.qvel[0] == x0[1] and d.time == 0.0  # state restored


@pytest.mark.parametrize("name", ["cartpole", "drone"])
def test_euler_linearization_blocks(name):
    """Semi-implicit Euler: q' = q (+) h v'  =>  A[:nv] = [I 0] + h A[nv:] (tangent space, small h)."""
    model = load_model(name)
    om, d = oracle_for(model)
    qpos, qvel, ctrl = random_states(model, name, 1, seed=11)
    d.qpos[:] = qpos[0]; d.qvel[:] = qvel[0] * (0 if name == "drone" else 1); d.ctrl[:] = ctrl[0]
    A, B = d.transition_fd(1e-6, True)
    nv, h = model.nv, float(model.opt.timestep)
    assert np.all(np.isfinite(A)) and np.all(np.isfinite(B)) and A.shape == (2 * nv, 2 * nv) and B.shape == (2 * nv, model.nu)
    if name == "cartpole":
        assert np.allclose(A[:nv], np.hstack([np.eye(nv), np.zeros((nv, nv))]) + h * A[nv:], atol=1e-7)
        assert np.allclose(B[:nv], h * B[nv:], atol=1e-9)


def test_ctrlrange_one_sided_differences():
    """At a ctrlrange bound mjd_transitionFD falls back to a one-sided difference (drone ctrl in [0, 13])."""
    model = load_model("drone")
    om, d = oracle_for(model)
    d.reset(0)
    d.ctrl[0] = 0.0
    A0, B0 = d.transition_fd(1e-6, True)
    d.ctrl[0] = 0.5
    A1, B1 = d.transition_fd(1e-6, True)
    assert np.allclose(B0[:, 0], B1[:, 0], rtol=1e-4, atol=1e-8)  # thrust map is linear in ctrl
    d.ctrl[0] = 14.0  # outside the range: clamped, derivative is zero
    _, B2 = d.transition_fd(1e-6, True)
    assert np.allclose(B2[:, 0], 0.0)


def test_humanoid_standing_contacts_and_warmstart():
    model = load_model("humanoid")
    om, d = oracle_for(model)
    d.reset(1)
    d.forward()
    assert d.ncon == 4 and d.nefc == 16
    geoms = sorted({(c["geom1"], c["geom2"]) for c in d.contacts()})
    assert geoms == [(0, 12), (0, 13)]  # floor vs the two left-foot capsules
    f = d.efc("efc_force")
    assert np.all(f >= 0) and f.sum() > 0
    # vertical force balance at the first instant is within the soft-contact regime
    total_mass = float(model.body_mass.sum())
    assert 0.2 * total_mass * 9.81 < d.qfrc_constraint[2] < 3 * total_mass * 9.81  # net upward force on the root
    w = np.array(d.qacc_warmstart)
    assert np.allclose(w, d.qacc)
    for _ in range(50):
        d.step()
    assert d.warnings == 0 and d.solver_iter <= 10


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_golden_trajectories(name):
    """Regression pin: committed oracle trajectories (bit-for-bit on the same compiler flags, 1e-12 otherwise)."""
    path = os.path.join(GOLDEN, f"traj_{name}.npz")
    z = np.load(path)
    model = load_model(name)
    om, d = oracle_for(model)
    d.qpos[:] = z["qpos0"]; d.qvel[:] = z["qvel0"]; d.ctrl[:] = z["ctrl"]
    A, B = d.transition_fd(1e-6, True)
    assert np.allclose(A, z["A"], rtol=1e-9, atol=1e-11) and np.allclose(B, z["B"], rtol=1e-9, atol=1e-11)
    for k in range(z["qpos_traj"].shape[0]):
        d.step()
        assert np.allclose(d.qpos, z["qpos_traj"][k], rtol=1e-10, atol=1e-12), (name, k)
        assert np.allclose(d.qvel, z["qvel_traj"][k], rtol=1e-9, atol=1e-11), (name, k)


def test_setconst_cross_check_oracle_vs_compiler():
    """invweight0 / meaninertia: NumPy Jacobian-sum M (compiler) vs the oracle's CRB + L'DL solve."""
    for name in MODEL_NAMES:
        model = load_model(name)
        om, _ = oracle_for(model)
        mean, dof, body, ten = om.setconst()
        assert abs(mean - model.stat.meaninertia) < 1e-12 * max(1, abs(mean))
        assert np.allclose(dof, model.dof_invweight0, rtol=1e-10)
        assert np.allclose(body, model.body_invweight0, rtol=1e-10, atol=1e-14)
        if model.ntendon:
            assert np.allclose(ten, model.tendon_invweight0, rtol=1e-10)


SENSOR_XML = """
<mujoco model="sensor-rig">
  <option timestep="{dt}" gravity="0.3 -0.2 -9.81"/>
  <worldbody>
    <body name="base" pos="0 0 1">
      <joint name="swing" type="hinge" axis="0 1 0" damping="0.1"/>
      <geom name="arm" type="capsule" fromto="0 0 0 0.4 0 0" size="0.03" mass="0.7"/>
      <body name="probe" pos="0.4 0 0" quat="0.9 0.1 -0.3 0.2">
        <joint name="slide" type="slide" axis="0 0 1"/>
        <joint name="twist" type="hinge" axis="1 0 0" pos="0 0.05 0"/>
        <geom name="tip" type="sphere" size="0.05" pos="0.02 0.1 -0.03" mass="0.4"/>
        <site name="imu" pos="0.03 -0.07 0.11" quat="0.7 0.2 0.5 -0.4"/>
      </body>
    </body>
    <body name="floater" pos="1 1 2">
      <freejoint name="root"/>
      <geom name="hull" type="box" size="0.1 0.2 0.05" mass="1.3" contype="0" conaffinity="0"/>
      <site name="imu2" pos="0.05 0.02 -0.01" quat="0.5 0.5 -0.5 0.5"/>
    </body>
  </worldbody>
  <actuator>
    <motor name="m0" joint="swing"/>
    <motor name="m1" joint="twist"/>
  </actuator>
  <sensor>
    <jointpos name="s_pos" joint="slide"/>
    <jointvel name="s_vel" joint="twist"/>
    <framepos name="p_site" objtype="site" objname="imu"/>
    <framepos name="p_body" objtype="body" objname="probe"/>
    <framepos name="p_xbody" objtype="xbody" objname="probe"/>
    <framepos name="p_geom" objtype="geom" objname="tip"/>
    <framequat name="q_site" objtype="site" objname="imu"/>
    <framequat name="q_xbody" objtype="xbody" objname="probe"/>
    <framequat name="q_geom" objtype="geom" objname="tip"/>
    <gyro name="gyro" site="imu"/>
    <velocimeter name="velo" site="imu"/>
    <accelerometer name="acc" site="imu"/>
    <gyro name="gyro2" site="imu2"/>
    <velocimeter name="velo2" site="imu2"/>
    <accelerometer name="acc2" site="imu2" cutoff="5"/>
  </sensor>
</mujoco>
"""


def _sensor_rig(dt):
    import warnings
    from mujoco_template import _mj as mj

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel.from_xml_string(SENSOR_XML.format(dt=dt))


def _sens(model, d, name):
    i = model.names["sensor"].index(name)
    return np.array(d.sensordata[model.sensor_adr[i]: model.sensor_adr[i] + model.sensor_dim[i]])


def _quat_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_sensors_against_kinematic_identities():
    """jointpos/jointvel read the state; frame sensors equal the pose arrays; gyro / velocimeter / accelerometer equal the
    site-frame angular velocity, d/dt(site position) and d/dt(site velocity) - g, checked by differencing a tiny step."""
    dt = 1e-6
    model = _sensor_rig(dt)
    assert model.nsensor == 15 and model.nsensordata == 2 + 4 * 3 + 3 * 4 + 6 * 3
    om, d = oracle_for(model)
    rng = np.random.default_rng(5)
    d.qpos[:3] = rng.uniform(-0.5, 0.5, 3)
    q = rng.normal(size=4); d.qpos[6:10] = q / np.linalg.norm(q)
    d.qvel[:] = rng.normal(size=model.nv)
    d.ctrl[:] = [0.3, -0.2]
    d.forward()
    s0 = {n: _sens(model, d, n) for n in model.names["sensor"]}
    site = d.site_xpos.copy()
    assert s0["s_pos"][0] == d.qpos[1] and s0["s_vel"][0] == d.qvel[2]
    assert np.array_equal(s0["p_site"], d.site_xpos[0]) and np.array_equal(s0["p_xbody"], d.xpos[2])
    assert np.array_equal(s0["p_body"], d.xipos[2]) and np.array_equal(s0["p_geom"], d.geom_xpos[1])
    assert np.allclose(_quat_mat(s0["q_site"]), d.site_xmat[0].reshape(3, 3), atol=1e-14)
    assert np.allclose(_quat_mat(s0["q_geom"]), d.geom_xmat[1].reshape(3, 3), atol=1e-14)
    assert np.array_equal(s0["q_xbody"], d.xquat[2])
    # free body: qvel[3:6] is the body-frame angular velocity; the site frame is a fixed rotation of it
    Rs2 = _quat_mat([0.5, 0.5, -0.5, 0.5])
    assert np.allclose(s0["gyro2"], Rs2.T @ d.qvel[6:9], atol=1e-13)
    R1, R2 = d.site_xmat[0].reshape(3, 3).copy(), d.site_xmat[1].reshape(3, 3).copy()
    v1_world, v2_world = R1 @ s0["velo"], R2 @ s0["velo2"]
    d.step()
    d.forward()
    # velocimeter: site position differenced over the step (semi-implicit Euler moves positions with the NEW velocity)
    s1 = {n: _sens(model, d, n) for n in model.names["sensor"]}
    R1n, R2n = d.site_xmat[0].reshape(3, 3), d.site_xmat[1].reshape(3, 3)
    assert np.allclose((d.site_xpos[0] - site[0]) / dt, R1n @ s1["velo"], atol=2e-5)
    assert np.allclose((d.site_xpos[1] - site[1]) / dt, R2n @ s1["velo2"], atol=2e-5)
    # accelerometer: world-frame site acceleration minus gravity, rotated into the site frame
    g = np.array([0.3, -0.2, -9.81])
    a1 = (R1n @ s1["velo"] - v1_world) / dt
    a2 = (R2n @ s1["velo2"] - v2_world) / dt
    assert np.allclose(R1.T @ (a1 - g), s0["acc"], atol=5e-4)
    assert np.allclose(np.clip(R2.T @ (a2 - g), -5, 5), s0["acc2"], atol=5e-4)


def test_sensors_rest_reading_and_rk4_stage_zero():
    model = load_model("drone")
    om, d = oracle_for(model)
    d.reset(0); d.forward()
    assert np.allclose(d.sensordata, [0, 0, 0, 0, 0, 9.81, 1, 0, 0, 0], atol=1e-10)  # hover: gyro 0, accel +g, level
    # RK4: sensordata after a step is that of the step's initial state (sub-stages skip sensors)
    pend = load_model("pendulum")
    assert pend.nsensordata == 0


def test_free_body_momentum_is_conserved_without_gravity_and_fluid():
    """SURVEY.md 8c pin 3: a torque-free rigid body keeps its linear momentum and world-frame angular momentum."""
    import warnings
    from mujoco_template import _mj as mj

    xml = """<mujoco><option timestep="0.0005" gravity="0 0 0"/><worldbody><body pos="0 0 1"><freejoint/>
      <geom type="box" size="0.1 0.25 0.05" mass="1.7" contype="0" conaffinity="0"/></body></worldbody></mujoco>"""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = mj.MjModel.from_xml_string(xml)
    om, d = oracle_for(model)
    d.qvel[:] = [0.3, -0.2, 0.1, 2.0, -1.0, 3.0]
    inertia = np.asarray(model.body_inertia[1])

    def momenta():
        R = d.xmat[1].reshape(3, 3)
        return 1.7 * d.qvel[:3].copy(), R @ (inertia * d.qvel[3:6])

    d.forward()
    p0, L0 = momenta()
    for _ in range(2000):
        d.step()
    d.forward()
    p1, L1 = momenta()
    assert np.allclose(p1, p0, atol=1e-13)
    assert np.linalg.norm(L1 - L0) <= 2e-3 * np.linalg.norm(L0)        # first-order integrator: small drift only
    assert abs(np.linalg.norm(d.qpos[3:7]) - 1.0) < 1e-12
