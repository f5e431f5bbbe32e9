"""TEST INFRASTRUCTURE -- an independent NumPy derivation of the quantities the oracle must reproduce.

Written from the textbook / documented equations (MuJoCo "Computation" chapter: kinematic trees, soft-constraint
model, solver parameters), NOT from oracle/mjstep_oracle.c, and on purpose along different routes:

* kinematics by direct frame composition down the tree;
* mass matrix by the Jacobian sum  M = sum_b m_b Jp'Jp + Jr' I_b Jr  (no composite bodies);
* bias forces by d'Alembert / virtual power with the body accelerations obtained by differentiating the
  kinematics numerically along the zero-acceleration path q(t) = q (+) v t  (no recursive Newton-Euler);
* constraint rows from the contact geometry and the geometric point Jacobians;
* impedance, reference acceleration and regulariser from solref / solimp / invweight0 in closed form;
* the constrained acceleration by an active-set iteration on the convex problem, with the KKT residual as
  the acceptance test (no line search, no warm start).

Only the compiled model constants (tree topology, body frames, inertias, invweight0) are shared with the
product; those are cross-checked separately (tests/test_oracle_analytic.py::test_setconst_...).
"""
from __future__ import annotations

import numpy as np

FREE, BALL, SLIDE, HINGE = 0, 1, 2, 3
MINVAL = 1e-15


# ------------------------------------------------------------------ rotations
def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def quat_rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rodrigues(axis, angle):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def quat_exp(w):
    """unit quaternion of the rotation vector w"""
    ang = np.linalg.norm(w)
    if ang < 1e-300:
        return np.array([1.0, 0, 0, 0])
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])


# ------------------------------------------------------------------ kinematics
class Frames:
    """world poses of all bodies, joint anchors / axes, inertial frames"""

    def __init__(self, m, qpos):
        nb = m.nbody
        self.m = m
        self.R = [np.eye(3) for _ in range(nb)]
        self.p = [np.zeros(3) for _ in range(nb)]
        self.anchor = [None] * m.njnt
        self.axis = [None] * m.njnt
        for b in range(1, nb):
            par = int(m.body_parentid[b])
            R = self.R[par] @ quat_rot(m.body_quat[b])
            p = self.p[par] + self.R[par] @ m.body_pos[b]
            for j in range(int(m.body_jntadr[b]), int(m.body_jntadr[b]) + int(m.body_jntnum[b])):
                qa, t = int(m.jnt_qposadr[j]), int(m.jnt_type[j])
                if t == FREE:
                    p = np.array(qpos[qa:qa + 3], float)
                    R = quat_rot(np.array(qpos[qa + 3:qa + 7], float))
                    self.anchor[j], self.axis[j] = p.copy(), None
                elif t == SLIDE:
                    ax = R @ m.jnt_axis[j]
                    self.anchor[j], self.axis[j] = p + R @ m.jnt_pos[j], ax
                    p = p + ax * (qpos[qa] - m.qpos0[qa])
                elif t == HINGE:
                    ax = R @ m.jnt_axis[j]
                    anchor = p + R @ m.jnt_pos[j]
                    Rn = rodrigues(ax, qpos[qa] - m.qpos0[qa]) @ R   # rotate the body about the world-frame axis ...
                    p = anchor - Rn @ m.jnt_pos[j]                   # ... keeping the anchor where it is
                    R = Rn
                    self.anchor[j], self.axis[j] = anchor, ax
                else:
                    raise NotImplementedError("ball joints are outside the compiled subset")
            self.R[b], self.p[b] = R, p
        self.com = [self.p[b] + self.R[b] @ m.body_ipos[b] for b in range(nb)]
        self.Ri = [self.R[b] @ quat_rot(m.body_iquat[b]) for b in range(nb)]
        self.Iw = [self.Ri[b] @ np.diag(m.body_inertia[b]) @ self.Ri[b].T for b in range(nb)]

    def geom_pose(self, g):
        b = int(self.m.geom_bodyid[g])
        return self.p[b] + self.R[b] @ self.m.geom_pos[g], self.R[b] @ quat_rot(self.m.geom_quat[g])

    def jac(self, body, point):
        """geometric Jacobians (3 x nv each) of a world point rigidly attached to `body`, w.r.t. qvel"""
        m = self.m
        jp, jr = np.zeros((3, m.nv)), np.zeros((3, m.nv))
        b = int(body)
        while b > 0:
            for j in range(int(m.body_jntadr[b]), int(m.body_jntadr[b]) + int(m.body_jntnum[b])):
                d, t = int(m.jnt_dofadr[j]), int(m.jnt_type[j])
                if t == FREE:
                    jp[:, d:d + 3] = np.eye(3)                       # linear velocity of the body origin, world frame
                    for k in range(3):                                # angular velocity components in the BODY frame
                        a = self.R[b][:, k]
                        jr[:, d + 3 + k] = a
                        jp[:, d + 3 + k] = np.cross(a, point - self.p[b])
                elif t == SLIDE:
                    jp[:, d] = self.axis[j]
                else:
                    jr[:, d] = self.axis[j]
                    jp[:, d] = np.cross(self.axis[j], point - self.anchor[j])
            b = int(m.body_parentid[b])
        return jp, jr


def integrate_pos(m, qpos, qvel, t):
    """q (+) v t : the configuration reached after time t at constant generalised velocity"""
    q = np.array(qpos, float)
    for j in range(m.njnt):
        qa, d = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        if int(m.jnt_type[j]) == FREE:
            q[qa:qa + 3] += t * qvel[d:d + 3]
            quat = quat_mul(q[qa + 3:qa + 7], quat_exp(t * np.asarray(qvel[d + 3:d + 6])))  # body-frame angular velocity
            q[qa + 3:qa + 7] = quat / np.linalg.norm(quat)
        else:
            q[qa] += t * qvel[d]
    return q


def mass_matrix(m, qpos):
    f = Frames(m, qpos)
    M = np.diag(np.asarray(m.dof_armature, float).copy())
    for b in range(1, m.nbody):
        jp, jr = f.jac(b, f.com[b])
        M += m.body_mass[b] * jp.T @ jp + jr.T @ f.Iw[b] @ jr
    return M


def _body_twists(m, qpos, qvel):
    f = Frames(m, qpos)
    out = []
    for b in range(1, m.nbody):
        jp, jr = f.jac(b, f.com[b])
        out.append((jp @ qvel, jr @ qvel))
    return out


def bias_forces(m, qpos, qvel, dt=2e-4):
    """c(q, v) + gravity term, i.e. the generalised force with M qacc + bias = applied at qacc = 0 (d'Alembert):
    sum_b  Jp' m (a_com - g) + Jr' (I alpha + w x I w), accelerations differentiated along q(t) = q (+) v t."""
    def twists(t):
        return _body_twists(m, integrate_pos(m, qpos, qvel, t), qvel)

    # fourth-order central difference of the body twists in time
    tw = {k: twists(k * dt) for k in (-2, -1, 1, 2)}
    f = Frames(m, qpos)
    g = np.asarray(m.gravity, float)
    out = np.zeros(m.nv)
    for i, b in enumerate(range(1, m.nbody)):
        acc = [(-tw[2][i][c] + 8 * tw[1][i][c] - 8 * tw[-1][i][c] + tw[-2][i][c]) / (12 * dt) for c in (0, 1)]
        jp, jr = f.jac(b, f.com[b])
        w = jr @ qvel
        out += jp.T @ (m.body_mass[b] * (acc[0] - g)) + jr.T @ (f.Iw[b] @ acc[1] + np.cross(w, f.Iw[b] @ w))
    return out


# ------------------------------------------------------------------ soft-constraint parameters
def impedance(solimp, r):
    """d(r) of the documented five-parameter impedance curve; r = distance - margin"""
    d0, dw, width, mid, power = [float(x) for x in solimp]
    d0, dw, mid = np.clip(d0, 1e-4, 0.9999), np.clip(dw, 1e-4, 0.9999), np.clip(mid, 1e-4, 0.9999)
    power = max(1.0, power)
    if d0 == dw or width <= MINVAL:
        return 0.5 * (d0 + dw)
    x = abs(r) / width
    if x >= 1:
        return dw
    if x == 0:
        return d0
    if power == 1:
        y = x
    elif x <= mid:
        y = x ** power / mid ** (power - 1)
    else:
        y = 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return d0 + y * (dw - d0)


def stiffness_damping(solref, solimp, timestep):
    """(k / d(r), b) of the reference acceleration  aref = -b (J v) - k r,  k = d(r) K"""
    dmax = float(np.clip(solimp[1], 1e-4, 0.9999))
    if solref[0] > 0:
        tc, dr = max(float(solref[0]), 2 * timestep), float(solref[1])     # refsafe: time constant >= 2 h
        return 1.0 / max(MINVAL, dmax * dmax * tc * tc * dr * dr), 2.0 / max(MINVAL, dmax * tc)
    return -float(solref[0]) / max(MINVAL, dmax * dmax), -float(solref[1]) / max(MINVAL, dmax)


def row_parameters(solref, solimp, r, vel, diag_approx, timestep, pyramid_mu=None):
    """(D, aref, d) of one constraint row.  diag_approx: the invweight0-based estimate of the row's inverse inertia.
    Pyramidal edges: regulariser 2 mu^2 R of the edge estimate (1 + mu^2) * translational weight (impratio = 1)."""
    d = impedance(solimp, r)
    K, B = stiffness_damping(solref, solimp, timestep)
    R = max(MINVAL, (1 - d) / d * diag_approx)
    if pyramid_mu is not None:
        R = 2 * pyramid_mu * pyramid_mu * R
    return 1.0 / R, -B * vel - K * d * r, d


# ------------------------------------------------------------------ the convex problem
def solve_rows(M, qfrc_smooth, J, D, aref, max_sweeps=200):
    """argmin_a  1/2 (a - a0)' M (a - a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2   by iterating on the active set."""
    a0 = np.linalg.solve(M, qfrc_smooth)
    if len(D) == 0:
        return a0, np.zeros(0)
    active = (J @ a0 - aref) < 0
    a = a0
    seen = set()
    for _ in range(max_sweeps):
        Ja, Da = J[active], D[active]
        a = np.linalg.solve(M + Ja.T @ (Da[:, None] * Ja), qfrc_smooth + Ja.T @ (Da * aref[active]))
        new = (J @ a - aref) < 0
        if np.array_equal(new, active):
            break
        key = new.tobytes()
        if key in seen:       # cycling between two sets: fall back to enumeration over the rows in doubt
            raise RuntimeError("active-set iteration cycled")
        seen.add(key)
        active = new
    force = -D * np.minimum(0.0, J @ a - aref)
    return a, force


def kkt_residual(M, qfrc_smooth, J, D, aref, qacc):
    """stationarity residual of the convex problem at qacc (its only optimality condition: the cost is C1)"""
    force = -D * np.minimum(0.0, J @ qacc - aref) if len(D) else np.zeros(0)
    r = M @ qacc - qfrc_smooth - (J.T @ force if len(D) else 0.0)
    return float(np.max(np.abs(r))), force


# ------------------------------------------------------------------ contact geometry (plane against sphere / capsule)
def plane_sphere(plane_pos, plane_R, centre, radius):
    n = plane_R[:, 2]
    dist = float(n @ (centre - plane_pos)) - radius
    return dist, centre - n * (radius + 0.5 * dist), n


def plane_capsule(plane_pos, plane_R, centre, R, radius, halflen):
    out = []
    for s in (1.0, -1.0):
        out.append(plane_sphere(plane_pos, plane_R, centre + s * halflen * R[:, 2], radius))
    return out


def contact_rows(m, frames, body1, body2, pos, frame, dim, mu):
    """constraint Jacobian rows of one contact: relative velocity of body2 w.r.t. body1 at `pos`, in the contact frame
    (rows of `frame`: normal, tangent 1, tangent 2); condim 1: the normal row; condim 3: four pyramid edges."""
    jp2, _ = frames.jac(body2, pos)
    jp1, _ = frames.jac(body1, pos)
    rel = frame @ (jp2 - jp1)
    if dim == 1:
        return rel[:1]
    return np.stack([rel[0] + mu * rel[1], rel[0] - mu * rel[1], rel[0] + mu * rel[2], rel[0] - mu * rel[2]])


# ------------------------------------------------------------------ one semi-implicit Euler step
def euler_step(m, qpos, qvel, qfrc_smooth, J, force):
    """v+ = v + h (M + h diag(damping))^-1 (smooth + J' f);  q+ = q (+) h v+   (joint damping integrated implicitly)"""
    h = float(m.opt.timestep)
    M = mass_matrix(m, qpos)
    total = np.asarray(qfrc_smooth, float) + (J.T @ force if len(force) else 0.0)
    v = np.asarray(qvel, float) + h * np.linalg.solve(M + h * np.diag(np.asarray(m.dof_damping, float)), total)
    return integrate_pos(m, qpos, v, h), v
