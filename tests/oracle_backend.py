"""TEST INFRASTRUCTURE: a CPU backend for ``_mj.MjData`` built on the oracle, so the host-side
logic of the package (Env step semantics, observations, recorder, drivers) can be exercised
without a GPU.  The product never selects this backend; it exists only under tests/."""
import numpy as np

from conftest import oracle_for

_FIELDS = ("qpos", "qvel", "ctrl", "qacc_warmstart", "xpos", "xquat", "xipos", "geom_xpos", "site_xpos", "subtree_com",
           "qacc", "qfrc_bias")


class OracleBackend:
    host_mapped = True
    nenv = 1

    def __init__(self, model):
        self.model = model
        self.om, self.od = oracle_for(model)
        m = model
        dims = dict(qpos=m.nq, qvel=m.nv, ctrl=m.nu, qacc_warmstart=m.nv, xpos=3 * m.nbody, xquat=4 * m.nbody,
                    xipos=3 * m.nbody, geom_xpos=3 * m.ngeom, site_xpos=3 * m.nsite, subtree_com=3 * m.nbody, qacc=m.nv,
                    qfrc_bias=m.nv, qfrc_inverse=m.nv, actuator_moment=m.nu * m.nv, sensordata=m.nsensordata)
        self.buf = {k: np.zeros((v, 1)) for k, v in dims.items()}
        for k in ("flags", "ncon", "nefc", "solver_iter"):
            self.buf[k] = np.zeros((1, 1), dtype=np.int32)
        self.calls = dict(step=0, forward=0, linearize=0, jacobian=0)

    def array(self, name):
        return self.buf[name]

    def _push(self):
        od = self.od
        od.qpos[:] = self.buf["qpos"][:, 0]; od.qvel[:] = self.buf["qvel"][:, 0]
        if self.model.nu:
            od.ctrl[:] = self.buf["ctrl"][:, 0]
        od.qacc_warmstart[:] = self.buf["qacc_warmstart"][:, 0]

    def _pull(self, state=True):
        od, b = self.od, self.buf
        if state:
            b["qpos"][:, 0] = od.qpos; b["qvel"][:, 0] = od.qvel
        b["qacc_warmstart"][:, 0] = od.qacc_warmstart

    def _pull_derived(self):
        od, b = self.od, self.buf
        b["xpos"][:, 0] = od.xpos.ravel(); b["xquat"][:, 0] = od.xquat.ravel(); b["xipos"][:, 0] = od.xipos.ravel()
        b["geom_xpos"][:, 0] = od.geom_xpos.ravel(); b["site_xpos"][:, 0] = od.site_xpos.ravel()
        b["subtree_com"][:, 0] = od.subtree_com.ravel(); b["qacc"][:, 0] = od.qacc; b["qfrc_bias"][:, 0] = od.qfrc_bias
        b["sensordata"][:, 0] = od.sensordata
        b["ncon"][0, 0] = od.ncon; b["nefc"][0, 0] = od.nefc; b["solver_iter"][0, 0] = od.solver_iter

    def step(self, nsteps=1, derived=True):
        self.calls["step"] += nsteps
        self._push()
        for _ in range(nsteps):
            self.od.step()
        self._pull()
        self._pull_derived()  # oracle derived arrays are those of the pre-integration forward pass

    def forward(self):
        self.calls["forward"] += 1
        self._push()
        self.od.forward()
        self._pull(state=False)
        self._pull_derived()

    def linearize(self, eps, centered, out=None):
        import torch

        self.calls["linearize"] += 1
        self._push()
        A, B = self.od.transition_fd(eps, centered)
        return torch.as_tensor(A[:, :, None].copy()), torch.as_tensor(B[:, :, None].copy())

    def jacobian(self, kind, objid, want_rot):
        import torch

        self.calls["jacobian"] += 1
        self._push()
        self.od.forward()
        name = {0: "site", 1: "body", 2: "bodycom", 3: "subtreecom"}[kind]
        jp, jr = self.od.jac(name, objid)
        t = lambda a: torch.as_tensor(a[:, :, None].copy())  # noqa: E731
        return t(jp), (t(jr) if (want_rot and jr is not None) else None)

    def inverse(self):
        self._push()
        self.buf["qfrc_inverse"][:, 0] = self.od.inverse(self.buf["qacc"][:, 0])
        self.buf["actuator_moment"][:, 0] = self.od.actuator_moment.ravel()

    def integrate_pos_host(self, qpos, qvel, dt):
        qpos[:] = self.od.integrate_pos(qpos, qvel, dt)

    def differentiate_pos_host(self, out, dt, qpos1, qpos2):
        out[:] = self.od.differentiate_pos(dt, qpos1, qpos2)


def make_env(model, **kw):
    """Env whose MjData is backed by the oracle (CPU)."""
    import mujoco_template as mt
    from mujoco_template import _mj as mj

    data = mj.MjData(model, backend=OracleBackend(model))
    return mt.Env(mt.ModelHandle(model, data), **kw)
