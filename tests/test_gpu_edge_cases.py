"""GPU edge cases: ragged / tiny batches, models outside the compiled-in specialisations (generic kernels),
no-actuator models, joint limits, RK4 with springs and servo actuators, capacity errors."""
import os
import warnings

import numpy as np
import pytest

from conftest import load_model, oracle_for
from test_mjcf_compiler import ARM_XML
from test_host_logic import BASE_XML

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def _compile(xml):
    from mujoco_template import _mj as mj

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel.from_xml_string(xml)


def _parity_rollout(model, qpos, qvel, ctrl, nsteps, tol=1e-9, check_lin=True):
    import torch
    from mujoco_template import _mj as mj

    n = qpos.shape[0]
    data = mj.BatchData(model, n)
    dev = data.qpos.device
    om, od = oracle_for(model)
    warm = np.zeros((n, model.nv))
    for s in range(nsteps):
        data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=dev)); data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=dev))
        if model.nu:
            data.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=dev))
        data.qacc_warmstart.copy_(torch.as_tensor(warm.T.copy(), device=dev))
        if check_lin and s % 10 == 0:
            A, B = data.backend.linearize(1e-6, True)
            od.reset(); od.qpos[:] = qpos[0]; od.qvel[:] = qvel[0]; od.qacc_warmstart[:] = warm[0]
            if model.nu:
                od.ctrl[:] = ctrl[0]
            Ao, Bo = od.transition_fd(1e-6, True)
            assert _rel(A[:, :, 0].cpu().numpy(), Ao) <= 1e-6
            if model.nu:
                assert _rel(B[:, :, 0].cpu().numpy(), Bo) <= 1e-6
            assert tuple(B.shape) == (2 * model.nv, model.nu, n)
        mj.mj_step(model, data)
        for e in range(n):
            od.reset(); od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.qacc_warmstart[:] = warm[e]
            if model.nu:
                od.ctrl[:] = ctrl[e]
            od.step()
            qpos[e] = od.qpos; qvel[e] = od.qvel; warm[e] = od.qacc_warmstart
        assert _rel(data.qpos.cpu().numpy().T, qpos) <= tol, s
        assert _rel(data.qvel.cpu().numpy().T, qvel) <= tol, s
    return data


@pytest.mark.parametrize("n", [1, 2, 31, 129, 1000])
def test_ragged_batch_sizes(n):
    model = load_model("cartpole")
    rng = np.random.default_rng(n)
    qpos = rng.uniform(-0.3, 0.3, (n, 2)); qvel = rng.uniform(-1, 1, (n, 2)); ctrl = rng.uniform(-2, 2, (n, 1))
    _parity_rollout(model, qpos, qvel, ctrl, 5)


def test_reference_fixture_model_limits_and_servo():
    """The reference's own test fixture (hinge limited to +-1 degree, motor + position servo): generic kernels."""
    model = _compile(BASE_XML)
    n = 8
    rng = np.random.default_rng(0)
    qpos = rng.uniform(-0.03, 0.03, (n, 1))          # beyond the +-0.01745 rad limit for several envs
    qvel = rng.uniform(-0.5, 0.5, (n, 1)); ctrl = rng.uniform(-1, 1, (n, 2))
    data = _parity_rollout(model, qpos, qvel, ctrl, 30)
    assert data.backend.batch.kernel_variant == "generic"
    assert int(data.nefc.max()) >= 1                   # the joint limit was active somewhere


def test_two_link_arm_rk4_springs_generic():
    model = _compile(ARM_XML)
    n = 16
    rng = np.random.default_rng(1)
    qpos = rng.uniform(-1.2, 1.2, (n, 2)); qvel = rng.uniform(-2, 2, (n, 2)); ctrl = rng.uniform(-1.5, 1.5, (n, 2))
    qpos[0, 0] = 1.7                                   # past the +-90 degree shoulder limit
    _parity_rollout(model, qpos, qvel, ctrl, 25)


def test_jit_specialised_kernels_for_a_user_model(tmp_path, monkeypatch):
    """A model that is not one of the pre-compiled examples gets register-resident kernels at load time
    (MjModel.specialize -> nvcc -> registered with libb2mj.so), with the same parity as the generic kernels."""
    import shutil

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    monkeypatch.setenv("B2_SPEC_CACHE", str(tmp_path))
    xml = ARM_XML.replace('timestep="0.004" integrator="RK4"', 'timestep="0.003"')  # its own blob (and Euler): other tests keep the generic kernels
    assert xml != ARM_XML
    model = _compile(xml)
    so = model.specialize("user_arm")
    assert so is not None and os.path.exists(so)
    n = 40
    rng = np.random.default_rng(2)
    qpos = rng.uniform(-1.2, 1.2, (n, 2)); qvel = rng.uniform(-2, 2, (n, 2)); ctrl = rng.uniform(-1.5, 1.5, (n, 2))
    qpos[0, 0] = 1.7
    data = _parity_rollout(model, qpos, qvel, ctrl, 25)
    assert data.backend.batch.kernel_variant == "user_arm"
    with pytest.raises(Exception):
        model.specialize()  # too late: the model is already on the device


def test_model_without_actuators():
    xml = BASE_XML[: BASE_XML.index("<actuator>")] + "</mujoco>"
    model = _compile(xml)
    assert model.nu == 0
    n = 4
    qpos = np.array([[0.0], [0.005], [-0.01], [0.012]]); qvel = np.array([[0.1], [0.0], [-0.2], [0.3]])
    _parity_rollout(model, qpos, qvel, np.zeros((n, 0)), 10)


def test_capacity_and_argument_errors():
    import mujoco_template as mt
    from mujoco_template import _capi, _mj as mj

    bodies = "".join(f'<body pos="{i} 0 1"><joint type="slide"/><geom size=".1"/></body>' for i in range(40))
    big = _compile(f"<mujoco><worldbody>{bodies}</worldbody></mujoco>")
    with pytest.raises(mt.ConfigError, match="size class"):
        mj.BatchData(big, 4)
    model = load_model("cartpole")
    with pytest.raises(mt.ConfigError):
        mj.BatchData(model, 0)
    data = mj.BatchData(model, 4)
    with pytest.raises(mt.ConfigError):
        data.backend.batch.jacobian(data.backend.state_struct(), _capi.JAC_SITE, 7, data.qpos.data_ptr(), None)
    with pytest.raises(mt.LinearizationError):
        data.backend.linearize(0.0, True)
    with pytest.raises(mt.ConfigError):
        data.backend.step(0)


def test_actuator_group_mask_reaches_the_kernels():
    import mujoco_template as mt

    model = _compile(BASE_XML)
    env = mt.Env(mt.ModelHandle(model))
    env.reset()
    env.data.ctrl[:] = [5.0, 0.0]
    env.step()
    moved = float(env.data.qvel[0])
    assert moved > 0
    env.reset()
    env.handle.set_enabled_actuator_groups([1])      # disables the torque motor (group 0)
    env.data.ctrl[:] = [5.0, 0.0]
    env.step()
    assert abs(float(env.data.qvel[0])) < 1e-12
    assert env.data.backend.batch.kernel_variant == "generic"


def test_humanoid_lying_down_many_contacts_warp_vs_lane(monkeypatch):
    """Prone keyframe: ~10+ contacts, 40+ rows.  Warp engine and lane engine agree with the oracle step by step."""
    import torch
    from mujoco_template import _mj as mj

    model = load_model("humanoid")
    n = 6
    qpos = np.tile(model.key_qpos[2], (n, 1)); qpos[:, 2] += np.linspace(0.0, 0.05, n)
    qvel = np.zeros((n, model.nv)); ctrl = np.zeros((n, model.nu))
    data = _parity_rollout(model, qpos.copy(), qvel.copy(), ctrl, 40, tol=1e-8, check_lin=False)
    assert data.backend.batch.kernel_variant == "generic-warp" and int(data.ncon.max()) >= 8 and int(data.flags.max()) == 0
    monkeypatch.setenv("B2_DISABLE_WARP", "1")
    data2 = _parity_rollout(model, qpos.copy(), qvel.copy(), ctrl, 40, tol=1e-8, check_lin=False)
    assert data2.backend.batch.kernel_variant == "generic"


@pytest.mark.parametrize("name", ["pendulum", "drone", "humanoid"])
def test_inverse_dynamics_and_steady_ctrl0(name):
    """b2_inverse vs the oracle's mj_inverse; forward/inverse consistency; batched steady-state controls."""
    import torch
    import mujoco_template as mt
    from mujoco_template import _mj as mj
    from conftest import random_states

    model = load_model(name)
    n = 16
    qpos, qvel, ctrl = random_states(model, name, n, seed=41)
    env = mt.BatchedEnv(model, n)
    dev = env.data.qpos.device
    env.data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=dev)); env.data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=dev))
    env.data.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=dev))
    env.forward()                                     # qacc of the forward dynamics
    qacc = env.data.qacc.cpu().numpy().T.copy()
    mj.mj_inverse(model, env.data)                    # inverse dynamics for that qacc
    qfrc = env.data.qfrc_inverse.cpu().numpy().T
    moment = env.data.actuator_moment.cpu().numpy().reshape(model.nu, model.nv, n)
    om, od = oracle_for(model)
    for e in range(n):
        od.reset(); od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
        ref = od.inverse(qacc[e])
        assert _rel(qfrc[e], ref) <= 1e-9, (name, e)
        assert _rel(moment[:, :, e], od.actuator_moment) <= 1e-12
        od.forward()
        assert _rel(qfrc[e], np.array(od.qfrc_actuator)) <= 1e-6   # inverse(forward(u)) recovers the actuator force
    u = mt.batched_steady_ctrl0(env)
    assert tuple(u.shape) == (model.nu, n) and bool(torch.isfinite(u).all())
    if name == "pendulum":
        expect = float(model.body_mass[1]) * 9.81 * 0.25 * np.sin(qpos[:, 0]) 
        # qvel != 0 adds no bias for a single hinge about a fixed axis; damping is zero
        assert np.allclose(u.cpu().numpy()[0], expect, atol=1e-10)


def test_sensordata_matches_oracle_all_kinds_generic_kernels():
    """Every compiled sensor kind (two-body chain + free body, generic kernels), N=1 Env view and a batch, FP64."""
    import torch
    import mujoco_template as mt
    from mujoco_template import _mj as mj
    from test_oracle_analytic import SENSOR_XML

    model = _compile(SENSOR_XML.format(dt=0.002))
    n = 33
    rng = np.random.default_rng(11)
    qpos = np.tile(model.qpos0, (n, 1)); qpos[:, :3] += rng.uniform(-0.5, 0.5, (n, 3))
    quat = rng.normal(size=(n, 4)); qpos[:, 6:10] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    qvel = rng.normal(size=(n, model.nv)); ctrl = rng.uniform(-1, 1, (n, model.nu))
    data = mj.BatchData(model, n)
    dev = data.qpos.device
    data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=dev)); data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=dev))
    data.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=dev))
    om, od = oracle_for(model)
    for phase in ("forward", "step"):
        (mj.mj_forward if phase == "forward" else mj.mj_step)(model, data)
        got = data.sensordata.cpu().numpy().T
        assert got.shape == (n, model.nsensordata)
        for e in range(n):
            od.reset(); od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
            od.forward()   # a step's sensordata is that of its pre-integration forward pass
            assert _rel(got[e], od.sensordata) <= 1e-12, (phase, e)
    # N = 1 Env: the observation is a zero-copy view of data.sensordata
    env = mt.Env(mt.ModelHandle(model), obs_spec=mt.ObservationSpec(include_sensordata=True))
    env.reset()
    env.data.qvel[:] = qvel[0]
    res = env.step()
    assert np.shares_memory(res.obs["sensordata"], env.data.sensordata)
    od.reset(); od.qvel[:] = qvel[0]; od.forward()
    assert _rel(np.array(env.data.sensordata), od.sensordata) <= 1e-12


@pytest.mark.parametrize("spec", [True, False])
def test_drone_imu_sensors_specialised_and_generic(monkeypatch, spec):
    import torch
    import mujoco_template as mt
    from conftest import random_states

    if not spec:
        monkeypatch.setenv("B2_DISABLE_SPEC", "1")
    model = load_model("drone")
    n = 64
    qpos, qvel, ctrl = random_states(model, "drone", n, seed=4)
    benv = mt.BatchedEnv(model, n, obs_spec=mt.ObservationSpec(include_sensordata=True))
    assert benv.data.backend.batch.kernel_variant == ("drone" if spec else "generic")
    dev = benv.data.qpos.device
    benv.data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=dev)); benv.data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=dev))
    benv.data.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=dev))
    res = benv.step()
    got = res.obs["sensordata"].cpu().numpy().T
    om, od = oracle_for(model)
    for e in range(n):
        od.reset(); od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
        od.forward()
        assert _rel(got[e], od.sensordata) <= 1e-11, e
    # FP32 engine: same readings to single precision
    b32 = mt.BatchedEnv(model, n, precision=32, obs_spec=mt.ObservationSpec(include_sensordata=True))
    b32.data.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=dev, dtype=torch.float32))
    b32.data.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=dev, dtype=torch.float32))
    b32.data.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=dev, dtype=torch.float32))
    got32 = b32.step().obs["sensordata"].cpu().numpy().T.astype(float)
    assert np.max(np.abs(got32 - got)) <= 5e-4 * max(1.0, np.max(np.abs(got)))


def _variant_of(model, **changes):
    """A second model of the same size class: the compiled model with some constants changed (new blob, new handle)."""
    import copy

    from mujoco_template import _mj as mj

    c = copy.deepcopy(model._c)
    for k, v in changes.items():
        c[k] = np.asarray(v, dtype=np.asarray(c[k]).dtype) if isinstance(c[k], np.ndarray) else v
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel(c)


@pytest.mark.parametrize("family", ["tiny-generic", "large-warp", "large-lane"])
def test_two_models_of_one_size_class_interleaved_on_two_streams(family, monkeypatch):
    """No process-wide model state: batches of two different models of the same size class, launched alternately on two
    streams with no synchronisation in between, each follow their own oracle (round 1 kept one constant-memory image per
    size class and swapped it under a lock; a launch could pick up the other model's constants)."""
    import torch
    from mujoco_template import _mj as mj

    if family == "tiny-generic":
        monkeypatch.setenv("B2_DISABLE_SPEC", "1")
        a = _compile(ARM_XML)
        b = _variant_of(a, timestep=0.003, gravity=[0.0, 0.0, -3.0])
        n, nsteps, tol = 64, 20, 1e-9
    else:
        if family == "large-lane":
            monkeypatch.setenv("B2_DISABLE_WARP", "1")
        a = load_model("humanoid")
        b = _variant_of(a, timestep=0.004, gravity=[0.0, 0.0, -5.0])
        n, nsteps, tol = 8, 6, 1e-9
    rng = np.random.default_rng(5)
    models, datas, streams, states = (a, b), [], [torch.cuda.Stream(), torch.cuda.Stream()], []
    for m in models:
        if m.nq == 2:
            qpos = rng.uniform(-1.0, 1.0, (n, 2)); qvel = rng.uniform(-1, 1, (n, 2))
        else:
            qpos = np.tile(m.key_qpos[1], (n, 1)); qpos[:, 7:] += rng.normal(0, 0.02, (n, m.nq - 7)); qvel = rng.normal(0, 0.01, (n, m.nv))
        ctrl = rng.uniform(-0.3, 0.3, (n, m.nu))
        d = mj.BatchData(m, n)
        d.qpos.copy_(torch.as_tensor(qpos.T.copy(), device=d.qpos.device)); d.qvel.copy_(torch.as_tensor(qvel.T.copy(), device=d.qpos.device))
        d.ctrl.copy_(torch.as_tensor(ctrl.T.copy(), device=d.qpos.device))
        datas.append(d); states.append((qpos, qvel, ctrl))
    assert datas[0].backend.batch.size_class == datas[1].backend.batch.size_class
    torch.cuda.synchronize()
    for _ in range(nsteps):  # alternate launches, each batch on its own stream, nothing waits for anything
        for d, s in zip(datas, streams):
            with torch.cuda.stream(s):
                d.backend.stream = s.cuda_stream
                d.backend.step(1, derived=False)
    torch.cuda.synchronize()
    for m, d, (qpos, qvel, ctrl) in zip(models, datas, states):
        om, od = oracle_for(m)
        for e in range(n):
            od.reset(); od.qpos[:] = qpos[e]; od.qvel[:] = qvel[e]; od.ctrl[:] = ctrl[e]
            for _ in range(nsteps):
                od.step()
            qpos[e] = od.qpos; qvel[e] = od.qvel
        assert _rel(d.qpos.cpu().numpy().T, qpos) <= tol * nsteps
        assert _rel(d.qvel.cpu().numpy().T, qvel) <= tol * nsteps
    # the two models really differ (the test would pass trivially otherwise)
    assert float(models[0].opt.timestep) != float(models[1].opt.timestep)


def test_graph_replay_keeps_its_model_when_another_model_runs_in_between(monkeypatch):
    """A captured BatchedEnv step replays with the model it was captured with, whatever ran on the device since."""
    import torch
    import mujoco_template as mt

    monkeypatch.setenv("B2_DISABLE_SPEC", "1")
    a = _compile(ARM_XML)
    b = _variant_of(a, timestep=0.002, gravity=[0.0, 0.0, -1.0])
    n = 32
    envs = [mt.BatchedEnv(m, n, controller=mt.ZeroController()) for m in (a, b)]
    rng = np.random.default_rng(9)
    q0 = rng.uniform(-1, 1, (n, 2))
    for env in envs:
        env.reset()
        env.data.qpos.copy_(torch.as_tensor(q0.T.copy(), device=env.data.qpos.device))
        env.forward()
        env.enable_cuda_graph(True)
    for _ in range(6):
        for env in envs:
            env.step(return_obs=False)
    torch.cuda.synchronize()
    for m, env in zip((a, b), envs):
        om, od = oracle_for(m)
        ref = np.empty((n, 2))
        for e in range(n):
            od.reset(); od.qpos[:] = q0[e]
            for _ in range(6):
                od.step()
            ref[e] = od.qpos
        assert _rel(env.data.qpos.cpu().numpy().T, ref) <= 1e-8


def test_random_controller_kernel_statistics_reset_and_graph_replay():
    """b2_random_controls: U(lo, hi) per env, actuator and call (fresh numbers on every call, also when the call is replayed
    from a captured CUDA graph), reproducible per (seed, env), and the fused episode reset."""
    import torch
    import mujoco_template as mt
    from mujoco_template.batched_controllers import BatchedRandomController

    model = load_model("drone")
    n = 4096
    draws = {}
    for seed in (0, 0, 1):
        env = mt.BatchedEnv(model, n, controller=BatchedRandomController(0.0, 13.0, seed=seed, reset_below=(2, 0.5)))
        env.reset(0)
        z0 = env.data.qpos[2].clone()
        env.data.qpos[2, :7] = 0.2          # seven drones below the reset height, moving
        env.data.qvel[:, :7] = 1.0
        out = []
        for _ in range(3):
            env.controller(model, env.data, 0.0)
            out.append(env.data.ctrl.clone())
        torch.cuda.synchronize()
        assert torch.equal(env.data.qpos[2], z0) and float(env.data.qvel[:, :7].abs().max()) == 0.0  # reset to the prepare() state
        draws.setdefault(seed, []).append(torch.stack(out))
    a, b, c = draws[0][0], draws[0][1], draws[1][0]
    assert torch.equal(a, b) and not torch.equal(a, c)                       # per-seed reproducible, seeds differ
    assert not torch.equal(a[0], a[1]) and not torch.equal(a[1], a[2])       # fresh numbers every call
    u = a.flatten().cpu().numpy()
    assert u.min() >= 0.0 and u.max() < 13.0
    assert abs(u.mean() - 6.5) < 0.05 and abs(u.std() - 13.0 / np.sqrt(12.0)) < 0.05
    assert abs(np.corrcoef(a[0, 0].cpu().numpy(), a[0, 1].cpu().numpy())[0, 1]) < 0.05   # actuators independent
    assert abs(np.corrcoef(a[0, 0].cpu().numpy(), a[1, 0].cpu().numpy())[0, 1]) < 0.05   # calls independent
    # graph replay: the per-env draw counter lives on the device, so replays keep drawing fresh controls
    env = mt.BatchedEnv(model, n, controller=BatchedRandomController(0.0, 13.0, seed=3))
    env.reset(0)
    env.enable_cuda_graph(True)
    seen = []
    for _ in range(5):
        env.step(return_obs=False)
        seen.append(env.data.ctrl.clone())
    torch.cuda.synchronize()
    assert all(not torch.equal(seen[i], seen[i + 1]) for i in range(4))


def test_fp32_instantiations_of_the_auxiliary_kernels(tmp_path):
    """b2_dlqr, b2_random_controls and the recorder gather in FP32 mode (B2_F32 batches / float32 tensors)."""
    import torch
    from scipy.linalg import solve_discrete_are

    import mujoco_template as mt
    from mujoco_template.batched_controllers import BatchedRandomController, batched_dlqr_gain

    rng = np.random.default_rng(2)
    A = rng.normal(0, 0.5, (16, 4, 4)); B = rng.normal(0, 1.0, (16, 4, 2))
    K, P = batched_dlqr_gain(torch.as_tensor(A, device="cuda", dtype=torch.float32), torch.as_tensor(B, device="cuda", dtype=torch.float32),
                             np.eye(4), np.eye(2), tol=1e-6)
    assert K.dtype == torch.float32
    for e in range(16):
        Pe = solve_discrete_are(A[e], B[e], np.eye(4), np.eye(2))
        assert np.max(np.abs(P[e].cpu().numpy() - Pe)) <= 2e-3 * np.max(np.abs(Pe))
    model = load_model("drone")
    env = mt.BatchedEnv(model, 512, controller=BatchedRandomController(0.0, 13.0, seed=5), precision=32)
    env.reset(0)
    imu = 0
    with mt.BatchedStateControlRecorder(env, log_path=tmp_path / "f32.csv", env_indices=[3, 500], chunk_steps=4, store_rows=True,
                                        probes=[mt.ArrayProbe("imu_z_m", "site_xpos", 3 * imu + 2)]) as rec:
        mt.run_passive_headless(env, max_steps=6, hooks=rec, return_obs=False)
    assert env.data.ctrl.dtype == torch.float32 and 0.0 <= float(env.data.ctrl.min()) and float(env.data.ctrl.max()) < 13.0
    rows = np.array([r[1:] for r in rec.rows], dtype=float)
    assert rows.shape == (12, 1 + model.nq + model.nv + model.nu + 1) and np.isfinite(rows).all()
    assert abs(rows[-1, 0] - 6 * float(model.opt.timestep)) < 1e-6
    assert np.allclose(rows[-2:, 1 + model.nq + model.nv: 1 + model.nq + model.nv + model.nu], env.data.ctrl[:, [3, 500]].T.cpu().numpy())
