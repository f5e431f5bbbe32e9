"""Independent pins of the oracle's tree and constraint paths (VERDICT r01 "next" item 1).

Everything expected here comes from tests/indep_dynamics.py: NumPy written from the documented equations along
routes that share no code and no algorithm with oracle/mjstep_oracle.c (Jacobian-sum mass matrix instead of
composite bodies, virtual power with numerically differentiated kinematics instead of recursive Newton-Euler,
closed-form impedance / aref / R, an active-set solve with a KKT acceptance test instead of the Newton solver).
The same cases run against the CUDA path in tests/test_gpu_independent.py.

What this cannot replace is a run of the real engine (not installable offline: DESIGN.md section 5)."""
import warnings

import numpy as np
import pytest

import indep_dynamics as ind
from conftest import load_model, oracle_for, random_states

SPHERE_XML = """<mujoco><option timestep="0.002" tolerance="1e-14"/>
<worldbody>
 <geom name="floor" type="plane" size="0 0 1" condim="{cd}" friction="0.7 0.005 0.0001" solref="0.015 0.8" solimp="0.85 0.97 0.004 0.4 3"/>
 <body name="ball" pos="0 0 0.099"><freejoint/>
  <geom name="g" type="sphere" size="0.1" mass="2.5" condim="{cd}" friction="0.7 0.005 0.0001" solref="0.015 0.8" solimp="0.85 0.97 0.004 0.4 3"/>
 </body>
</worldbody></mujoco>"""

HINGE_XML = """<mujoco><option timestep="0.004" gravity="0 0 -9.81" tolerance="1e-14"/>
<worldbody>
 <body name="arm" pos="0 0 1"><joint name="h" type="hinge" axis="0 1 0" limited="true" range="-30 40" damping="0.05"
   solreflimit="0.03 0.9" solimplimit="0.8 0.95 0.02 0.3 2" margin="0.01"/>
  <geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.04" mass="1.2" contype="0" conaffinity="0"/>
 </body>
</worldbody>
<actuator><motor joint="h" gear="2"/></actuator></mujoco>"""


def compile_xml(text):
    from mujoco_template import _mj as mj

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel.from_xml_string(text)


def with_tolerance(model, tol):
    """the same compiled model with another solver tolerance (so that the Newton iterate is the fixed point)"""
    from mujoco_template import _mj as mj

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel({**model._c, "tolerance": float(tol)})


# ------------------------------------------------------------------ (ii) tree path: M and bias forces
def scrambled_state(model, name, seed):
    """a generic configuration (random joint angles, random base orientation) at large generalised velocities"""
    rng = np.random.default_rng(seed)
    qpos, _, _ = random_states(model, name, 1, seed=seed)
    q = qpos[0].copy()
    for j in range(model.njnt):
        qa = int(model.jnt_qposadr[j])
        if int(model.jnt_type[j]) == ind.FREE:
            q[qa + 2] += 2.0  # clear of the floor: no contact forces in qacc
            quat = rng.normal(size=4)
            q[qa + 3:qa + 7] = quat / np.linalg.norm(quat)
        else:
            q[qa] += rng.normal(0, 0.3)
    return q, rng.normal(0, 1.5, model.nv)


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "drone", "humanoid"])
def test_mass_matrix_and_bias_match_jacobian_sum_and_virtual_power(name):
    model = load_model(name)
    om, d = oracle_for(model)
    for seed in range(3):
        q, v = scrambled_state(model, name, seed)
        d.reset(); d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = 0
        d.forward()
        M = ind.mass_matrix(model, q)
        assert np.max(np.abs(M - d.qM)) <= 1e-13 * np.max(np.abs(M))
        bias = ind.bias_forces(model, q, v)
        assert np.max(np.abs(bias - d.qfrc_bias)) <= 1e-9 * max(1.0, np.max(np.abs(bias))), (name, seed)


def test_humanoid_kinematics_match_frame_composition():
    model = load_model("humanoid")
    om, d = oracle_for(model)
    q, v = scrambled_state(model, "humanoid", 7)
    d.qpos[:] = q; d.qvel[:] = v; d.forward()
    f = ind.Frames(model, q)
    assert np.allclose(np.array(f.p), d.xpos, atol=1e-13)
    assert np.allclose(np.array(f.com), d.xipos, atol=1e-13)
    assert np.allclose(np.array([R.ravel() for R in f.R]), d.xmat, atol=1e-13)
    for g in range(model.ngeom):
        p, R = f.geom_pose(g)
        assert np.allclose(p, d.geom_xpos[g], atol=1e-13) and np.allclose(R.ravel(), d.geom_xmat[g], atol=1e-13)
    # subtree centres of mass by direct summation
    for b in range(model.nbody):
        members = [k for k in range(model.nbody) if b in _ancestors(model, k)]
        mass = sum(model.body_mass[k] for k in members)
        if mass > 0:
            com = sum(model.body_mass[k] * f.com[k] for k in members) / mass
            assert np.allclose(com, d.subtree_com[b], atol=1e-13)


def _ancestors(model, b):
    out = {b}
    while b > 0:
        b = int(model.body_parentid[b])
        out.add(b)
    return out


# ------------------------------------------------------------------ (i) closed-form constraint rows
def expected_sphere_case(model, q, v):
    """independent (J, D, aref, qacc) of the ball-on-floor model"""
    f = ind.Frames(model, q)
    pp, pR = f.geom_pose(0)
    centre, _ = f.geom_pose(1)
    dist, pos, n = ind.plane_sphere(pp, pR, centre, float(model.geom_size[1][0]))
    dim, mu = int(model.pair_dim[0]), float(model.pair_friction[0][0])
    frame = np.array([n, [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])  # right-handed, tangents as the engine orders them for n = z
    J = ind.contact_rows(model, f, 0, 1, pos, frame, dim, mu)
    tran = float(model.body_invweight0[0][0] + model.body_invweight0[1][0])
    rows = []
    for r in range(J.shape[0]):
        approx = tran if dim == 1 else tran + mu * mu * tran
        rows.append(ind.row_parameters(model.pair_solref[0], model.pair_solimp[0], dist - float(model.pair_margin[0]), float(J[r] @ v),
                                       approx, float(model.opt.timestep), pyramid_mu=None if dim == 1 else mu))
    D = np.array([r[0] for r in rows]); aref = np.array([r[1] for r in rows])
    M = ind.mass_matrix(model, q)
    smooth = -ind.bias_forces(model, q, v)
    qacc, force = ind.solve_rows(M, smooth, J, D, aref)
    return dict(dist=dist, pos=pos, J=J, D=D, aref=aref, qacc=qacc, force=force, M=M, smooth=smooth)


SPHERE_Q = np.array([0.1, -0.2, 0.0985, 1, 0, 0, 0], float)
SPHERE_V = np.array([0.4, -0.3, -0.2, 1.0, 2.0, -0.5])


@pytest.mark.parametrize("condim", [1, 3])
def test_sphere_on_plane_rows_and_acceleration(condim):
    model = compile_xml(SPHERE_XML.format(cd=condim))
    # free sphere: the invweight0 family has closed forms
    mass, inertia = 2.5, 0.4 * 2.5 * 0.1 ** 2
    assert np.allclose(model.body_invweight0[1], [1 / mass, 1 / inertia], rtol=1e-12)
    assert np.allclose(model.dof_invweight0, [1 / mass] * 3 + [1 / inertia] * 3, rtol=1e-12)
    exp = expected_sphere_case(model, SPHERE_Q, SPHERE_V)
    om, d = oracle_for(model)
    d.qpos[:] = SPHERE_Q; d.qvel[:] = SPHERE_V
    d.forward()
    assert d.ncon == 1 and d.nefc == (1 if condim == 1 else 4)
    c = d.contacts()[0]
    assert abs(c["dist"] - exp["dist"]) < 1e-15 and np.allclose(c["pos"], exp["pos"], atol=1e-15)
    assert np.allclose(c["frame"], [[0, 0, 1], [0, 1, 0], [-1, 0, 0]], atol=1e-15)
    assert np.allclose(d.efc("efc_J"), exp["J"], atol=1e-14)
    assert np.allclose(d.efc("efc_D"), exp["D"], rtol=1e-13)
    assert np.allclose(d.efc("efc_aref"), exp["aref"], rtol=1e-12, atol=1e-12)
    assert np.allclose(d.efc("efc_force"), exp["force"], rtol=1e-9, atol=1e-9)
    assert np.allclose(d.qacc, exp["qacc"], rtol=1e-10, atol=1e-9)
    if condim == 1:
        # one active row: the closed form of the normal acceleration, written out
        D, aref = exp["D"][0], exp["aref"][0]
        az = (mass * (-9.81) + D * aref) / (mass + D)
        assert abs(d.qacc[2] - az) < 1e-10
        # and the numbers themselves, from the formulas in the documentation
        dimp = ind.impedance(model.pair_solimp[0], exp["dist"])
        x = 0.0015 / 0.004
        assert abs(dimp - (0.85 + (x ** 3 / 0.4 ** 2) * (0.97 - 0.85))) < 1e-15
        assert abs(1 / D - (1 - dimp) / dimp * (1 / mass)) < 1e-15
        K, B = 1 / (0.97 ** 2 * 0.015 ** 2 * 0.8 ** 2), 2 / (0.97 * 0.015)
        assert abs(aref - (-B * SPHERE_V[2] - K * dimp * exp["dist"])) < 1e-10


def expected_hinge_case(model, q, v, u):
    f = ind.Frames(model, q)
    lo, hi = model.jnt_range[0]
    margin = float(model.jnt_margin[0])
    rows = []
    for side, dist in ((1.0, q[0] - lo), (-1.0, hi - q[0])):
        if dist < margin:
            J = np.array([[side]])
            D, aref, _ = ind.row_parameters(model.jnt_solref[0], model.jnt_solimp[0], dist - margin, float(J[0] @ v),
                                            float(model.dof_invweight0[0]), float(model.opt.timestep))
            rows.append((J, D, aref, dist))
    M = ind.mass_matrix(model, q)
    smooth = -ind.bias_forces(model, q, v) - model.dof_damping * v + np.array([2.0 * u])
    J = np.concatenate([r[0] for r in rows]) if rows else np.zeros((0, 1))
    D = np.array([r[1] for r in rows]); aref = np.array([r[2] for r in rows])
    qacc, force = ind.solve_rows(M, smooth, J, D, aref)
    return dict(J=J, D=D, aref=aref, qacc=qacc, force=force, pos=np.array([r[3] for r in rows]), M=M, smooth=smooth)


HINGE_CASES = [(np.deg2rad(41.5), 0.8, 0.3), (np.deg2rad(-31.0), -0.5, -1.0), (np.deg2rad(39.7), 0.2, 0.0), (0.1, 1.0, 0.5)]


def test_hinge_limit_rows_and_acceleration():
    model = compile_xml(HINGE_XML)
    # rod about its end: I = m (l^2 / 3 + ...) -- the capsule's own inertia comes from the compiler; dof weight = 1 / M
    M0 = ind.mass_matrix(model, model.qpos0)
    assert abs(model.dof_invweight0[0] - 1 / M0[0, 0]) < 1e-12
    om, d = oracle_for(model)
    for q0, v0, u in HINGE_CASES:
        q, v = np.array([q0]), np.array([v0])
        exp = expected_hinge_case(model, q, v, u)
        d.reset(); d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u
        d.forward()
        assert d.nefc == len(exp["D"])
        if d.nefc:
            assert np.allclose(d.efc("efc_J"), exp["J"]) and np.allclose(d.efc("efc_pos"), exp["pos"], atol=1e-15)
            assert np.allclose(d.efc("efc_D"), exp["D"], rtol=1e-13) and np.allclose(d.efc("efc_aref"), exp["aref"], rtol=1e-12)
            # one row: closed form
            D, aref, s = exp["D"][0], exp["aref"][0], exp["J"][0, 0]
            a0 = exp["smooth"][0] / exp["M"][0, 0]
            a = (exp["smooth"][0] + s * D * aref) / (exp["M"][0, 0] + D) if s * a0 - aref < 0 else a0
            assert abs(d.qacc[0] - a) < 1e-10 * max(1, abs(a))
        assert np.allclose(d.qacc, exp["qacc"], rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------ (iii) humanoid: rows from geometry, KKT fixed point
def humanoid_contact_state(seed):
    model = load_model("humanoid")
    qpos, qvel, ctrl = random_states(model, "humanoid", 1, seed=seed)
    return model, qpos[0], qvel[0] * 20, ctrl[0]


def expected_humanoid_rows(model, q, v):
    """limit rows (joints, tendons) then contact rows of plane-capsule / plane-sphere pairs, from geometry alone"""
    f = ind.Frames(model, q)
    h = float(model.opt.timestep)
    J, D, aref, pos = [], [], [], []
    for j in range(model.njnt):
        if not model.jnt_limited[j] or int(model.jnt_type[j]) < ind.SLIDE:
            continue
        val, margin, dof = q[int(model.jnt_qposadr[j])], float(model.jnt_margin[j]), int(model.jnt_dofadr[j])
        for side, dist in ((1.0, val - model.jnt_range[j][0]), (-1.0, model.jnt_range[j][1] - val)):
            if dist < margin:
                row = np.zeros(model.nv); row[dof] = side
                dd, ar, _ = ind.row_parameters(model.jnt_solref[j], model.jnt_solimp[j], dist - margin, float(row @ v),
                                               float(model.dof_invweight0[dof]), h)
                J.append(row); D.append(dd); aref.append(ar); pos.append(dist)
    for t in range(model.ntendon):
        if not model.tendon_limited[t]:
            continue
        row = np.zeros(model.nv); length = 0.0
        for w in range(int(model.tendon_adr[t]), int(model.tendon_adr[t]) + int(model.tendon_num[t])):
            jid = int(model.wrap_jntid[w])
            row[int(model.jnt_dofadr[jid])] = model.wrap_coef[w]
            length += model.wrap_coef[w] * q[int(model.jnt_qposadr[jid])]
        margin = float(model.tendon_margin[t])
        for side, dist in ((1.0, length - model.tendon_range[t][0]), (-1.0, model.tendon_range[t][1] - length)):
            if dist < margin:
                dd, ar, _ = ind.row_parameters(model.tendon_solref[t], model.tendon_solimp[t], dist - margin, float(side * row @ v),
                                               float(model.tendon_invweight0[t]), h)
                J.append(side * row); D.append(dd); aref.append(ar); pos.append(dist)
    contacts = []
    for p in range(len(model.pair_geom1)):
        g1, g2 = int(model.pair_geom1[p]), int(model.pair_geom2[p])
        if int(model.geom_type[g1]) != 0:
            continue  # body-body pairs: asserted inactive by the caller through the contact count
        pp, pR = f.geom_pose(g1)
        c2, R2 = f.geom_pose(g2)
        size = model.geom_size[g2]
        hits = [ind.plane_sphere(pp, pR, c2, float(size[0]))] if int(model.geom_type[g2]) == 2 else \
            ind.plane_capsule(pp, pR, c2, R2, float(size[0]), float(size[1]))
        for dist, cpos, n in hits:
            incl = float(model.pair_margin[p] - model.pair_gap[p])
            if dist < incl:
                contacts.append((p, dist, cpos, n))
    return f, J, D, aref, pos, contacts


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_humanoid_rows_from_geometry_and_kkt_fixed_point(seed):
    model, q, v, u = humanoid_contact_state(seed)
    tight = with_tolerance(model, 1e-15)
    om, d = oracle_for(tight)
    d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u
    d.forward()
    f, J, D, aref, pos, contacts = expected_humanoid_rows(model, q, v)
    assert d.ncon == len(contacts) and d.ncon >= 1
    oc = d.contacts()
    h = float(model.opt.timestep)
    for k, (p, dist, cpos, n) in enumerate(contacts):
        assert (oc[k]["geom1"], oc[k]["geom2"]) == (int(model.pair_geom1[p]), int(model.pair_geom2[p]))
        assert abs(oc[k]["dist"] - dist) < 1e-14 and np.allclose(oc[k]["pos"], cpos, atol=1e-14)
        frame = oc[k]["frame"]  # tangent choice is the engine's; it must complete n to a right-handed orthonormal frame
        assert np.allclose(frame[0], n, atol=1e-14) and np.allclose(frame @ frame.T, np.eye(3), atol=1e-14)
        assert np.allclose(np.cross(frame[0], frame[1]), frame[2], atol=1e-14)
        mu, dim = float(model.pair_friction[p][0]), int(model.pair_dim[p])
        b1, b2 = int(model.geom_bodyid[int(model.pair_geom1[p])]), int(model.geom_bodyid[int(model.pair_geom2[p])])
        rows = ind.contact_rows(model, f, b1, b2, cpos, frame, dim, mu)
        tran = float(model.body_invweight0[b1][0] + model.body_invweight0[b2][0])
        for r in rows:
            dd, ar, _ = ind.row_parameters(model.pair_solref[p], model.pair_solimp[p], dist - float(model.pair_margin[p]), float(r @ v),
                                           tran if dim == 1 else tran * (1 + mu * mu), h, pyramid_mu=None if dim == 1 else mu)
            J.append(r); D.append(dd); aref.append(ar); pos.append(dist)
    J, D, aref = np.array(J), np.array(D), np.array(aref)
    assert d.nefc == len(D)
    assert np.allclose(d.efc("efc_J"), J, atol=1e-12)
    assert np.allclose(d.efc("efc_pos"), pos, atol=1e-14)
    assert np.allclose(d.efc("efc_D"), D, rtol=1e-12)
    assert np.allclose(d.efc("efc_aref"), aref, rtol=1e-10, atol=1e-9)
    # smooth force from independent pieces: bias by virtual power, actuation and passive terms by their definitions
    M = ind.mass_matrix(model, q)
    assert np.allclose(d.qfrc_bias, ind.bias_forces(model, q, v), rtol=1e-9, atol=1e-9)
    smooth = np.array(d.qfrc_smooth)
    # (iii) the Newton result is the minimiser: stationarity of the convex cost to 1e-10 of the force scale
    res, force = ind.kkt_residual(M, smooth, J, D, aref, np.array(d.qacc))
    scale = max(1.0, np.max(np.abs(smooth)))
    assert res <= 1e-10 * scale, (res, scale)
    assert np.allclose(d.efc("efc_force"), force, rtol=1e-9, atol=1e-9 * scale)
    # and it is THE minimiser the independent active-set solve finds
    qacc, _ = ind.solve_rows(M, smooth, J, D, aref)
    assert np.allclose(d.qacc, qacc, rtol=1e-9, atol=1e-9 * np.max(np.abs(qacc)))
    # default tolerance (1e-8): same answer within what the stopping rule allows
    om2, d2 = oracle_for(model)
    d2.qpos[:] = q; d2.qvel[:] = v; d2.ctrl[:] = u
    d2.forward()
    assert np.allclose(d2.qacc, qacc, rtol=1e-6, atol=1e-6 * np.max(np.abs(qacc)))


def test_humanoid_smooth_force_terms():
    """qfrc_smooth = passive + actuator - bias with each term from its definition (joint springs / dampers, fixed-tendon
    springs, affine position / motor actuators with ctrl and force clamps)."""
    model, q, v, u = humanoid_contact_state(5)
    om, d = oracle_for(model)
    d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u * 8  # beyond ctrlrange on purpose
    d.forward()
    passive = -np.asarray(model.dof_damping) * v
    for j in range(model.njnt):
        if int(model.jnt_type[j]) >= ind.SLIDE and model.jnt_stiffness[j] != 0:
            qa = int(model.jnt_qposadr[j])
            passive[int(model.jnt_dofadr[j])] -= model.jnt_stiffness[j] * (q[qa] - model.qpos_spring[qa])
    for t in range(model.ntendon):
        row = np.zeros(model.nv); length = 0.0
        for w in range(int(model.tendon_adr[t]), int(model.tendon_adr[t]) + int(model.tendon_num[t])):
            jid = int(model.wrap_jntid[w])
            row[int(model.jnt_dofadr[jid])] = model.wrap_coef[w]; length += model.wrap_coef[w] * q[int(model.jnt_qposadr[jid])]
        lo, hi = model.tendon_lengthspring[t]
        frc = model.tendon_stiffness[t] * ((hi - length) if length > hi else (lo - length) if length < lo else 0.0)
        passive += row * (frc - model.tendon_damping[t] * float(row @ v))
    act = np.zeros(model.nv)
    for a in range(model.nu):
        jid = int(model.actuator_trnid[a][0]); dof = int(model.jnt_dofadr[jid]); gear = float(model.actuator_gear[a][0])
        c = float(np.clip(u[a] * 8, *model.actuator_ctrlrange[a])) if model.actuator_ctrllimited[a] else float(u[a] * 8)
        frc = model.actuator_gainprm[a] * c + model.actuator_biasprm[a][0] + model.actuator_biasprm[a][1] * gear * q[int(model.jnt_qposadr[jid])] \
            + model.actuator_biasprm[a][2] * gear * v[dof]
        if model.actuator_forcelimited[a]:
            frc = float(np.clip(frc, *model.actuator_forcerange[a]))
        act[dof] += gear * frc
    assert np.allclose(d.qfrc_passive, passive, atol=1e-12)
    assert np.allclose(d.qfrc_actuator, act, atol=1e-12)
    assert np.allclose(d.qfrc_smooth, passive + act - ind.bias_forces(model, q, v), rtol=1e-9, atol=1e-9)


def test_cartpole_slider_limit_and_floor_contact_kkt():
    """cartpole pushed past its slider range with the pole on the floor: limit + box/capsule-plane rows, KKT residual"""
    model = with_tolerance(load_model("cartpole"), 1e-15)
    om, d = oracle_for(model)
    seen = 0
    for q, v, u in (([2.05, 0.1], [1.0, 0.0], 100.0), ([-2.1, -0.3], [-0.5, 0.2], -250.0), ([0.0, 1.75], [0.0, 1.0], 0.0),
                    ([1.99, -1.8], [0.3, -1.0], 10.0), ([2.02, 1.78], [0.5, 0.5], 3.0)):
        d.reset(); d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u
        d.forward()
        seen += d.nefc
        assert d.nefc > 0
        M = ind.mass_matrix(model, np.array(q))
        assert np.allclose(M, d.qM, atol=1e-14)
        res, _ = ind.kkt_residual(M, np.array(d.qfrc_smooth), d.efc("efc_J"), d.efc("efc_D"), d.efc("efc_aref"), np.array(d.qacc))
        assert res <= 1e-10 * max(1.0, np.max(np.abs(d.qfrc_smooth)))
    assert seen >= 8


# ------------------------------------------------------------------ whole steps from independent pieces
def independent_sphere_step(model, q, v):
    exp = expected_sphere_case(model, q, v) if q[2] - 0.1 < 0 else None
    if exp is None:
        smooth = -ind.bias_forces(model, q, v)
        return ind.euler_step(model, q, v, smooth, np.zeros((0, model.nv)), np.zeros(0))
    return ind.euler_step(model, q, v, exp["smooth"], exp["J"], exp["force"])


@pytest.mark.parametrize("condim", [1, 3])
def test_sphere_bounce_trajectory_from_independent_stepper(condim):
    """120 steps of a spinning, sliding ball settling on the floor: oracle mj_step vs the NumPy stepper, per step from a
    shared state (<= 1e-9 relative, the north-star tolerance) and as a free-running trajectory."""
    model = compile_xml(SPHERE_XML.format(cd=condim))
    om, d = oracle_for(model)
    q, v = np.array([0.0, 0.0, 0.13, 1, 0, 0, 0], float), np.array([0.6, -0.4, -0.8, 3.0, -2.0, 1.0])
    d.qpos[:] = q; d.qvel[:] = v
    touched = 0
    for s in range(120):
        qn, vn = independent_sphere_step(model, np.array(d.qpos), np.array(d.qvel))
        q, v = independent_sphere_step(model, q, v)
        d.step()
        touched += d.ncon
        assert np.max(np.abs(d.qpos - qn)) <= 1e-9 and np.max(np.abs(d.qvel - vn)) <= 1e-9 * max(1, np.max(np.abs(vn))), s
    assert touched > 30
    assert np.max(np.abs(d.qpos - q)) <= 1e-7 and np.max(np.abs(d.qvel - v)) <= 1e-6  # free-running drift stays small


@pytest.mark.parametrize("seed", [0, 3])
def test_humanoid_step_from_independent_pieces(seed):
    model, q, v, u = humanoid_contact_state(seed)
    om, d = oracle_for(with_tolerance(model, 1e-15))
    d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u
    d.forward()
    J, D, aref = d.efc("efc_J"), d.efc("efc_D"), d.efc("efc_aref")  # pinned against geometry in the test above
    M = ind.mass_matrix(model, q)
    smooth = np.array(d.qfrc_smooth)
    qacc, force = ind.solve_rows(M, smooth, J, D, aref)
    qn, vn = ind.euler_step(model, q, v, smooth, J, force)
    d.step()
    assert np.max(np.abs(d.qvel - vn)) <= 1e-9 * max(1.0, np.max(np.abs(vn)))
    assert np.max(np.abs(d.qpos - qn)) <= 1e-9
