"""Host-side behaviour of the drop-in API, exercised on CPU through an oracle-backed MjData
(tests/oracle_backend.py).  Mirrors the reference's own tests for this path
(reference tests/test_mujoco_template.py:74-170, 364-588): shapes, call ordering, decimation,
info rules, observation contract, recorder schema, drivers."""
import csv
import warnings

import numpy as np
import pytest

import mujoco_template as mt
from mujoco_template import _mj as mj
from conftest import load_model
from oracle_backend import OracleBackend, make_env

BASE_XML = """
<mujoco model="template-test">
  <option timestep="0.005"/>
  <default>
    <joint limited="true" range="-1 1"/>
  </default>
  <worldbody>
    <body name="torso">
      <joint name="hinge" type="hinge" axis="0 0 1"/>
      <geom name="torso_geom" type="capsule" size="0.04 0.2" pos="0 0 0"/>
      <site name="tip" pos="0 0 0.2"/>
    </body>
  </worldbody>
  <actuator>
    <motor name="torque_act" joint="hinge" group="0" forcelimited="true" forcerange="-10 10"/>
    <position name="pos_act" joint="hinge" group="1" ctrllimited="true" ctrlrange="-0.5 0.5"/>
  </actuator>
  <sensor>
    <jointpos name="hinge_pos" joint="hinge"/>
  </sensor>
</mujoco>
"""


@pytest.fixture
def base_model():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mj.MjModel.from_xml_string(BASE_XML)


@pytest.fixture
def handle(base_model):
    data = mj.MjData(base_model, backend=OracleBackend(base_model))
    h = mt.ModelHandle(base_model, data)
    h.forward()
    return h


def test_model_handle_wraps_existing_data_and_rejects_mismatch(base_model):
    data = mj.MjData(base_model, backend=OracleBackend(base_model))
    data.qpos[0] = 0.25
    h = mt.ModelHandle.from_model_and_data(base_model, data)
    assert h.model is base_model and h.data is data
    env = mt.Env(h, obs_spec=mt.ObservationSpec(include_qpos=True))
    env.reset()
    assert env.data is data
    other = mj.MjModel.from_xml_string(BASE_XML.replace("<sensor>", "<!--").replace("</sensor>", "-->"))
    with pytest.raises(mt.ConfigError):
        mt.ModelHandle(other, data=data)


def test_step_can_skip_observation_but_still_runs_hooks(handle):
    calls = {"reward": 0, "done": 0, "info": 0}
    seen = []

    def reward_fn(m, d, obs): calls["reward"] += 1; seen.append(obs); return 0.0
    def done_fn(m, d, obs): calls["done"] += 1; seen.append(obs); return False
    def info_fn(m, d, obs): calls["info"] += 1; seen.append(obs); return {"value": float(d.time)}

    env = mt.Env(handle, obs_spec=mt.ObservationSpec(include_qpos=True), reward_fn=reward_fn, done_fn=done_fn, info_fn=info_fn)
    env.reset()
    res = env.step(return_obs=False)
    assert calls == {"reward": 1, "done": 1, "info": 1} and seen == [None, None, None]
    assert res.obs is None and res.reward == 0.0 and res.done is False
    assert res.info == {"value": pytest.approx(float(env.data.time))}
    with pytest.raises(mt.ConfigError):
        env.step(0)


def test_controller_ordering_precompute_and_info_shapes(handle):
    order = []
    backend = handle.data.backend

    class Lin:
        capabilities = mt.ControllerCapabilities(needs_linearization=True, needs_jacobians=("site:tip", "bodycom:torso"))
        prepared = 0
        def prepare(self, model, data): Lin.prepared += 1
        def __call__(self, model, data, t):
            order.append(("ctrl", backend.calls["linearize"], backend.calls["jacobian"], backend.calls["step"]))
            data.ctrl[:] = [0.3, 0.1]

    env = mt.Env(handle, obs_spec=mt.ObservationSpec(), controller=Lin())
    assert Lin.prepared == 1
    env.reset()
    assert Lin.prepared == 2                        # prepare runs at construction and again at reset
    res = env.step()
    nv, nu = env.model.nv, env.model.nu
    assert res.info["A"].shape == (2 * nv, 2 * nv) and res.info["B"].shape == (2 * nv, nu)
    assert np.all(np.isfinite(res.info["A"])) and np.all(np.isfinite(res.info["B"]))
    assert set(res.info["jacobians"]) == {"site:tip", "bodycom:torso"}
    assert res.info["jacobians"]["site:tip"]["jacp"].shape == (3, nv) and "jacr" not in res.info["jacobians"]["bodycom:torso"]
    assert "compat_warnings" in res.info              # once, in the first result after reset
    res2 = env.step(3)
    assert "compat_warnings" not in res2.info
    assert isinstance(res2.info["A"], list) and len(res2.info["A"]) == 3 and len(res2.info["jacobians"]) == 3
    # controller ran before the linearisation / jacobians / step of its own tick
    assert order[0] == ("ctrl", 0, 0, 0) and order[1] == ("ctrl", 1, 2, 1)


def test_control_decimation_and_info_key_collision(handle):
    ticks = []

    class Count:
        capabilities = mt.ControllerCapabilities(needs_linearization=True)
        def prepare(self, model, data): pass
        def __call__(self, model, data, t): ticks.append(t)

    env = mt.Env(handle, controller=Count(), control_decimation=2, info_fn=lambda m, d, o: {"A": 1})
    env.reset()
    with pytest.raises(mt.TemplateError, match="info key collision"):
        env.step(2)
    env = mt.Env(handle, controller=Count(), control_decimation=2)
    env.reset(); ticks.clear()
    res = env.step(2)
    assert len(ticks) == 1 and not isinstance(res.info["A"], list)   # one control tick in two substeps
    res = env.step(4)
    assert len(ticks) == 3 and len(res.info["A"]) == 2
    with pytest.raises(mt.ConfigError):
        mt.Env(handle, control_decimation=0)


def test_linearize_native_and_fallback_shapes(handle):
    nv, nu = handle.model.nv, handle.model.nu
    handle.data.qpos[0] = 0.004
    A, B = mt.linearize_discrete(handle.model, handle.data, use_native=True)
    A2, B2 = mt.linearize_discrete(handle.model, handle.data, use_native=False, horizon_steps=1)
    for M_, shape in ((A, (2 * nv, 2 * nv)), (B, (2 * nv, nu)), (A2, (2 * nv, 2 * nv)), (B2, (2 * nv, nu))):
        assert M_.shape == shape and np.all(np.isfinite(M_))
    # the fallback reproduces the reference's argument order for the position rows (base - new)
    assert np.allclose(A2[nv:], A[nv:], atol=1e-5) and np.allclose(A2[:nv], -A[:nv], atol=1e-5)
    assert handle.data.qpos[0] == 0.004 and handle.data.time == 0.0


def test_jacobian_tokens(handle):
    out = mt.compute_requested_jacobians(handle.model, handle.data, ["site:tip", "body:torso", "bodycom:torso", "subtreecom:torso"])
    assert out["site:tip"]["jacp"].shape == (3, 1) and out["body:torso"]["jacr"].shape == (3, 1)
    assert np.allclose(out["body:torso"]["jacr"][:, 0], [0, 0, 1])      # hinge about z
    with pytest.raises(mt.ConfigError):
        mt.compute_requested_jacobians(handle.model, handle.data, ["com"])
    with pytest.raises(mt.ConfigError):
        mt.compute_requested_jacobians(handle.model, handle.data, ["weird:thing"])
    with pytest.raises(mt.NameLookupError):
        mt.compute_requested_jacobians(handle.model, handle.data, ["site:nope"])


def test_observation_dict_flat_order_zero_copy_and_extras(handle):
    spec = mt.ObservationSpec(include_ctrl=True, include_time=True, include_sensordata=True, sites_pos=("tip",),
                              bodies_pos=("torso",), geoms_pos=("torso_geom",), subtree_com=("torso",),
                              extras={"twice": lambda m, d: 2 * np.array(d.qpos)})
    ext = mt.ObservationExtractor(handle.model, spec)
    handle.data.qpos[0] = 0.3
    handle.forward()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        obs = ext(handle.data)                            # the model has a jointpos sensor: no warning
    assert obs["sensordata"].shape == (handle.model.nsensordata,) == (1,)
    assert np.shares_memory(obs["sensordata"], handle.data.sensordata) and obs["sensordata"][0] == 0.3
    # reference tests/test_mujoco_template.py:291-314: a model without sensors warns once and returns an empty array
    bare_model = mj.MjModel.from_xml_string(BASE_XML.replace("<sensor>", "<!--").replace("</sensor>", "-->"))
    bare = mt.ModelHandle(bare_model, mj.MjData(bare_model, backend=OracleBackend(bare_model)))
    ext0 = mt.ObservationExtractor(bare.model, mt.ObservationSpec(include_sensordata=True))
    with pytest.warns(RuntimeWarning, match="sensordata"):
        assert ext0(bare.data)["sensordata"].shape == (0,)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ext0(bare.data)                                   # warns only once
    assert set(obs) == {"qpos", "qvel", "ctrl", "time", "sensordata", "sites_pos", "bodies_pos", "geoms_pos", "subtree_com", "twice"}
    assert np.shares_memory(obs["qpos"], handle.data.qpos) and obs["sites_pos"].shape == (1, 3) and obs["time"].shape == (1,)
    assert np.allclose(obs["sites_pos"][0], [0, 0, 0.2])
    flat = mt.ObservationExtractor(handle.model, mt.ObservationSpec(include_ctrl=True, include_time=True, sites_pos=("tip",), as_dict=False))(handle.data)
    assert flat.shape == (2 + 1 + 1 + 3 + 1,)              # sorted keys: ctrl, qpos, qvel, sites_pos, time
    assert np.allclose(flat[4:7], [0, 0, 0.2])
    copy = mt.ObservationExtractor(handle.model, mt.ObservationSpec(copy=True))(handle.data)
    assert not np.shares_memory(copy["qpos"], handle.data.qpos)
    with pytest.raises(mt.NameLookupError):
        mt.ObservationExtractor(handle.model, mt.ObservationSpec(sites_pos=("nope",)))
    with pytest.raises(ValueError):
        mt.ObservationExtractor(handle.model, mt.ObservationSpec(extras={"qpos": lambda m, d: [0.0]}))(handle.data)


def test_actuator_groups_mask(handle):
    assert handle.enabled_actuator_mask().tolist() == [True, True]
    handle.set_enabled_actuator_groups([1])
    assert handle.enabled_actuator_mask().tolist() == [False, True] and handle.model.opt.disableactuator == 1
    with pytest.raises(mt.CompatibilityError):
        handle.set_enabled_actuator_groups([])
    with pytest.raises(mt.ConfigError):
        handle.set_enabled_actuator_groups([40])
    with pytest.raises(mt.CompatibilityError):
        handle.set_enabled_actuator_groups([7])
    rep = mt.check_controller_compat(handle.model, mt.ControllerCapabilities(control_space=mt.ControlSpace.POSITION), None)
    assert rep.ok and any("lacks ctrlrange" in w for w in rep.warnings)


def test_passive_drivers_and_time_accumulation():
    """Loop exit uses accumulated float time: a 0.05 s run at dt=0.005 takes 11 steps, not 10."""
    model = load_model("pendulum")
    env = make_env(model, obs_spec=mt.ObservationSpec(include_time=True), controller=mt.ZeroController())
    env.reset()
    assert sum(1 for _ in env.passive(max_steps=3)) == 3
    env.reset()
    steps = mt.run_passive_headless(env, duration=0.05, max_steps=1000)
    t = 0.0
    k = 0
    while True:
        t += 0.005; k += 1
        if t >= 0.05:
            break
    assert steps == k and env.data.time == t
    with pytest.raises(mt.ConfigError):
        list(mt.iterate_passive(env, max_steps=0))
    hits = []
    list(mt.iterate_passive(env, max_steps=2, hooks=[hits.append, hits.append], return_obs=False))
    assert len(hits) == 4 and all(h.obs is None for h in hits)


def test_recorder_schema_and_csv(tmp_path):
    model = load_model("drone")
    env = make_env(model, controller=mt.ZeroController())
    env.reset("hover")
    probes = [mt.DataProbe("imu_z", lambda e, r: float(e.data.site_xpos[0][2])), mt.DataProbe("none", lambda e, r: None)]
    path = tmp_path / "log.csv"
    with mt.StateControlRecorder(env, log_path=path, probes=probes) as rec:
        mt.run_passive_headless(env, max_steps=3, hooks=rec)
    cols = rec.columns
    assert cols[0] == "time_s" and cols[1:8] == tuple(f"qpos[joint_0].{c}" for c in ("pos_x", "pos_y", "pos_z", "quat_w", "quat_x", "quat_y", "quat_z"))
    assert cols[8:14] == tuple(f"qvel[joint_0].{c}" for c in ("lin_x", "lin_y", "lin_z", "ang_x", "ang_y", "ang_z"))
    assert cols[14:18] == ("ctrl[thrust1]", "ctrl[thrust2]", "ctrl[thrust3]", "ctrl[thrust4]") and cols[18:] == ("imu_z", "none")
    rows = list(csv.reader(open(path)))
    assert rows[0] == list(cols) and len(rows) == 4 and rows[1][-1] == ""
    assert len(rec.rows) == 3 and rec.rows[0][0] == 0.01
    # derived quantities lag qpos by one step: the imu probe reads the pre-integration site position
    assert rec.rows[0][18] == pytest.approx(0.3 + 0.02) and rec.rows[1][18] < rec.rows[0][18]
    cart = make_env(load_model("cartpole"))
    rec2 = mt.StateControlRecorder(cart)
    assert rec2.columns == ("time_s", "qpos[slider]", "qvel[slider]", "qpos[hinge]", "qvel[hinge]", "ctrl[cart_force]")
    with pytest.raises(mt.ConfigError):
        mt.StateControlRecorder(cart, probes=[mt.DataProbe("a", lambda e, r: 0), mt.DataProbe("a", lambda e, r: 0)])


def test_reset_keyframe_and_errors():
    env = make_env(load_model("humanoid"))
    env.reset("squat")
    assert env.data.qpos[2] == 0.596
    env.reset(1)
    assert env.data.qpos[2] == 1.21948 and env.data.time == 0.0
    with pytest.raises(mt.NameLookupError):
        env.reset("nope")
    with pytest.raises(mt.ConfigError):
        env.reset(99)


def test_zero_controller_and_position_demo(handle):
    env = mt.Env(handle, controller=mt.ZeroController())
    env.reset()
    env.data.ctrl[:] = 1.0
    env.step()
    assert np.all(env.data.ctrl == 0)
    demo = mt.PositionTargetDemo(targets=np.array([0.1, 0.2]))
    env = mt.Env(handle, controller=demo)
    env.reset(); env.step()
    assert env.data.ctrl.tolist() == [0.1, 0.2]
    with pytest.raises(mt.ConfigError):
        mt.Env(handle, controller=mt.PositionTargetDemo(targets=np.zeros(3)))


def test_shard_range_partitions_exactly():
    for total, world in ((65536, 8), (10, 3), (7, 8)):
        spans = [mt.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(mt.ConfigError):
        mt.shard_range(8, 2, 2)


def test_steady_ctrl0_preserves_state_and_holds_the_pose(handle):
    model, data = handle.model, handle.data
    data.qpos[:] = 0.1
    data.qvel[:] = -0.05
    q0, v0 = np.copy(data.qpos), np.copy(data.qvel)
    u = mt.steady_ctrl0(model, data, qpos0=np.zeros(model.nq), qvel0=np.zeros(model.nv))
    assert u.shape == (model.nu,)
    np.testing.assert_allclose(data.qpos, q0)
    np.testing.assert_allclose(data.qvel, v0)
    with pytest.raises(mt.ConfigError):
        mt.steady_ctrl0(model, data, qpos0=np.zeros(model.nq + 1))
    # pendulum: the holding control equals the gravity torque m g l sin(theta) (gear 1)
    pend = load_model("pendulum")
    env = make_env(pend)
    u = mt.steady_ctrl0(env.model, env.data, qpos0=np.array([0.7]))
    assert abs(u[0] - float(pend.body_mass[1]) * 9.81 * 0.25 * np.sin(0.7)) < 1e-12
    # drone hover: four equal thrusts m g / 4 (site transmissions: state-dependent moment matrix)
    drone = load_model("drone")
    env = make_env(drone)
    u = mt.steady_ctrl0(env.model, env.data, qpos0=drone.key_qpos[0])
    assert np.allclose(u, 1.325 * 9.81 / 4, atol=1e-9)


def test_batched_dlqr_gain_needs_the_device():
    """LQR synthesis is a library kernel (b2_dlqr): CPU tensors are refused, there is no host fallback."""
    import torch

    import mujoco_template as mt
    from mujoco_template.batched_controllers import batched_dlqr_gain

    with pytest.raises(mt.TemplateError, match="CUDA"):
        batched_dlqr_gain(torch.eye(4)[None].double(), torch.ones(1, 4, 1).double(), np.eye(4), np.eye(1))

