"""The independent pins of tests/test_oracle_independent.py, run against the CUDA path (through the C-ABI).

Expected values come from tests/indep_dynamics.py (NumPy, written from the documented equations along routes that share
nothing with the oracle or the kernels); the oracle only supplies the contact-frame tangents for the humanoid rows, which
the CPU test checks to be a right-handed completion of the geometric normal."""
import numpy as np
import pytest

import indep_dynamics as ind
from conftest import load_model, oracle_for
from test_gpu_parity import _batch, _upload
from test_oracle_independent import (HINGE_CASES, HINGE_XML, SPHERE_Q, SPHERE_V, SPHERE_XML, compile_xml, expected_hinge_case,
                                     expected_sphere_case, humanoid_contact_state, independent_sphere_step, scrambled_state,
                                     with_tolerance)

pytestmark = pytest.mark.gpu


def _forward(model, qpos, qvel, ctrl):
    from mujoco_template import _mj as mj

    n = qpos.shape[0]
    data = _batch(model, n)
    _upload(data, qpos, qvel, ctrl)
    mj.mj_forward(model, data)
    return data


@pytest.mark.parametrize("condim", [1, 3])
def test_sphere_on_plane_acceleration_and_trajectory(condim):
    from mujoco_template import _mj as mj

    model = compile_xml(SPHERE_XML.format(cd=condim))
    exp = expected_sphere_case(model, SPHERE_Q, SPHERE_V)
    data = _forward(model, SPHERE_Q[None], SPHERE_V[None], np.zeros((1, 0)))
    assert int(data.ncon[0, 0]) == 1 and int(data.nefc[0, 0]) == (1 if condim == 1 else 4)
    assert np.allclose(data.qacc.cpu().numpy()[:, 0], exp["qacc"], rtol=1e-10, atol=1e-9)
    # settling trajectory: every CUDA step from a shared state against the NumPy stepper
    q, v = np.array([0.0, 0.0, 0.13, 1, 0, 0, 0], float), np.array([0.6, -0.4, -0.8, 3.0, -2.0, 1.0])
    data = _batch(model, 1)
    _upload(data, q[None], v[None], np.zeros((1, 0)))
    touched = 0
    for s in range(120):
        gq, gv = data.qpos.cpu().numpy()[:, 0], data.qvel.cpu().numpy()[:, 0]
        qn, vn = independent_sphere_step(model, gq, gv)
        mj.mj_step(model, data)
        touched += int(data.ncon[0, 0])
        assert np.max(np.abs(data.qpos.cpu().numpy()[:, 0] - qn)) <= 1e-9, s
        assert np.max(np.abs(data.qvel.cpu().numpy()[:, 0] - vn)) <= 1e-9 * max(1.0, np.max(np.abs(vn))), s
    assert touched > 30


def test_hinge_limit_acceleration():
    model = compile_xml(HINGE_XML)
    n = len(HINGE_CASES)
    qpos = np.array([[c[0]] for c in HINGE_CASES]); qvel = np.array([[c[1]] for c in HINGE_CASES]); ctrl = np.array([[c[2]] for c in HINGE_CASES])
    data = _forward(model, qpos, qvel, ctrl)
    qacc = data.qacc.cpu().numpy()
    for e, (q0, v0, u) in enumerate(HINGE_CASES):
        exp = expected_hinge_case(model, np.array([q0]), np.array([v0]), u)
        assert int(data.nefc[0, e]) == len(exp["D"])
        assert abs(qacc[0, e] - exp["qacc"][0]) <= 1e-10 * max(1.0, abs(exp["qacc"][0]))


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "drone", "humanoid"])
def test_kinematics_and_bias_forces_match_independent_routes(name):
    model = load_model(name)
    states = [scrambled_state(model, name, s) for s in range(4)]
    qpos = np.array([s[0] for s in states]); qvel = np.array([s[1] for s in states])
    data = _forward(model, qpos, qvel, np.zeros((len(states), model.nu)))
    bias = data.qfrc_bias.cpu().numpy()
    xpos = data.xpos.cpu().numpy(); xipos = data.xipos.cpu().numpy(); gx = data.geom_xpos.cpu().numpy()
    for e, (q, v) in enumerate(states):
        f = ind.Frames(model, q)
        assert np.allclose(xpos[:, e].reshape(-1, 3), np.array(f.p), atol=1e-12)
        assert np.allclose(xipos[:, e].reshape(-1, 3), np.array(f.com), atol=1e-12)
        assert np.allclose(gx[:, e].reshape(-1, 3), np.array([f.geom_pose(g)[0] for g in range(model.ngeom)]), atol=1e-12)
        b = ind.bias_forces(model, q, v)
        assert np.max(np.abs(bias[:, e] - b)) <= 1e-9 * max(1.0, np.max(np.abs(b))), (name, e)
        if int(data.ncon[0, e]) == 0 and int(data.nefc[0, e]) == 0:
            # unconstrained: qacc = M^-1 (passive - bias) with M from the Jacobian sum
            passive = -np.asarray(model.dof_damping) * v
            for j in range(model.njnt):
                if int(model.jnt_type[j]) >= ind.SLIDE and model.jnt_stiffness[j] != 0:
                    qa = int(model.jnt_qposadr[j])
                    passive[int(model.jnt_dofadr[j])] -= model.jnt_stiffness[j] * (q[qa] - model.qpos_spring[qa])
            if name in ("pendulum", "cartpole"):  # no tendon springs, no fluid forces, zero ctrl -> no actuator force
                a = np.linalg.solve(ind.mass_matrix(model, q), passive - b)
                assert np.allclose(data.qacc.cpu().numpy()[:, e], a, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_humanoid_contact_solve_is_the_independent_minimiser(seed):
    """warp engine: qacc with active foot contacts = the active-set minimiser of the convex problem built from rows the CPU
    test pins against geometry; the step = the NumPy semi-implicit Euler update."""
    from mujoco_template import _mj as mj

    model, q, v, u = humanoid_contact_state(seed)
    tight = with_tolerance(model, 1e-15)
    om, d = oracle_for(tight)
    d.qpos[:] = q; d.qvel[:] = v; d.ctrl[:] = u
    d.forward()
    J, D, aref = d.efc("efc_J"), d.efc("efc_D"), d.efc("efc_aref")
    M = ind.mass_matrix(model, q)
    smooth = np.array(d.qfrc_smooth)
    qacc, force = ind.solve_rows(M, smooth, J, D, aref)
    data = _forward(tight, q[None], v[None], u[None])
    assert int(data.ncon[0, 0]) == d.ncon and int(data.nefc[0, 0]) == d.nefc
    g = data.qacc.cpu().numpy()[:, 0]
    res, _ = ind.kkt_residual(M, smooth, J, D, aref, g)
    assert res <= 1e-9 * max(1.0, np.max(np.abs(smooth)))
    assert np.allclose(g, qacc, rtol=1e-8, atol=1e-8 * np.max(np.abs(qacc)))
    qn, vn = ind.euler_step(model, q, v, smooth, J, force)
    mj.mj_step(tight, data)
    assert np.max(np.abs(data.qvel.cpu().numpy()[:, 0] - vn)) <= 1e-9 * max(1.0, np.max(np.abs(vn)))
    assert np.max(np.abs(data.qpos.cpu().numpy()[:, 0] - qn)) <= 1e-9
