// Per-env rigid-body engine ("lane engine": one env per thread).
//
// Computes what the reference obtains from mj.mj_step / mj.mj_forward
// (reference mujoco_template/model.py:53-57): forward kinematics, composite-rigid-body
// mass matrix, L'DL factorisation, bias forces (RNE), passive + actuator forces, primitive
// collisions, limit / contact constraint rows, a warm-started Newton solve, and
// semi-implicit Euler or RK4 integration.  All model reads are uniform constant-bank
// loads; all per-env scratch lives in this struct (registers / L1-resident local memory).
#pragma once
#include "b2_math.cuh"
#include "b2_model_dev.cuh"

namespace b2 {

template <typename T, class D>
struct LaneEnv {
  const DevModel<T, D>& m;
  // ---- state
  T qpos[D::NQ], qvel[D::NV], ctrl[D::NU], warm[D::NV];
  // ---- position-dependent
  T xpos[3 * D::NB], xquat[4 * D::NB], xmat[9 * D::NB], xipos[3 * D::NB], ximat[9 * D::NB];
  T xanchor[3 * D::NJ], xaxis[3 * D::NJ];
  T geom_xpos[3 * D::NG], geom_xmat[9 * D::NG], site_xpos[3 * D::NS], site_xmat[9 * D::NS];
  T com[3 * D::NB], cinert[10 * D::NB], cdof[6 * D::NV];
  T M[D::NV * D::NV], LD[D::NV * D::NV], dinv[D::NV];
  T ten_len[D::NT], ten_J[D::NT * D::NV], act_len[D::NU], act_moment[D::NU * D::NV];
  // ---- velocity-dependent
  T cvel[6 * D::NB], cdof_dot[6 * D::NV], spat[6 * D::NB], spat2[10 * D::NB];  // spat/spat2: cacc, cfrc / crb scratch
  T f_bias[D::NV], f_passive[D::NV], f_smooth[D::NV], f_con[D::NV], a_smooth[D::NV], qacc[D::NV];
  // ---- contacts + constraint rows
  int ncon, nefc, niter, flags;
  T con_dist[D::NCON], con_pos[3 * D::NCON], con_frame[9 * D::NCON];
  short con_pair[D::NCON];
  signed char row_type[D::NEFC];
  short row_id[D::NEFC];
  T J[D::NEFC * D::NV], row_pos[D::NEFC], row_margin[D::NEFC], row_D[D::NEFC], row_aref[D::NEFC];
  // ---- solver scratch
  T Jaref[D::NEFC], Jv[D::NEFC], Ma[D::NV], Mv[D::NV], grad[D::NV], Mgrad[D::NV], search[D::NV];
  T cost, gauss, qg0, qg1, qg2;
  int ls_iter;

  __device__ explicit LaneEnv(const DevModel<T, D>& model) : m(model), ncon(0), nefc(0), niter(0), flags(0) {}

  // ------------------------------------------------------------------ position stage
  __device__ void kinematics() {
    xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
    quat_to_mat(xmat, xquat);
    xipos[0] = xipos[1] = xipos[2] = 0; quat_to_mat(ximat, xquat);
    for (int i = 1; i < m.nbody; i++) {
      T p[3], q[4];
      const int ja = m.body_jntadr[i], jn = m.body_jntnum[i], pid = m.body_parentid[i];
      if (jn == 1 && m.jnt_type[ja] == JNT_FREE) {
        const int qa = m.jnt_qposadr[ja];
        for (int k = 0; k < 3; k++) p[k] = qpos[qa + k];
        for (int k = 0; k < 4; k++) q[k] = qpos[qa + 3 + k];
        normalize4(q);
        for (int k = 0; k < 3; k++) { xanchor[3 * ja + k] = p[k]; xaxis[3 * ja + k] = m.jnt_axis[3 * ja + k]; }
      } else {
        if (pid) {
          mat_vec(p, xmat + 9 * pid, m.body_pos + 3 * i);
          for (int k = 0; k < 3; k++) p[k] += xpos[3 * pid + k];
          quat_mul(q, xquat + 4 * pid, m.body_quat + 4 * i);
        } else {
          for (int k = 0; k < 3; k++) p[k] = m.body_pos[3 * i + k];
          for (int k = 0; k < 4; k++) q[k] = m.body_quat[4 * i + k];
        }
        for (int j = ja; j < ja + jn; j++) {
          T anchor[3], axis[3];
          const int qa = m.jnt_qposadr[j];
          quat_rot(axis, m.jnt_axis + 3 * j, q);
          quat_rot(anchor, m.jnt_pos + 3 * j, q);
          for (int k = 0; k < 3; k++) anchor[k] += p[k];
          const T disp = qpos[qa] - m.qpos0[qa];
          if (m.jnt_type[j] == JNT_SLIDE) {
            for (int k = 0; k < 3; k++) p[k] += axis[k] * disp;
          } else {
            T ql[4], off[3];
            quat_axis_angle(ql, m.jnt_axis + 3 * j, disp);
            quat_mul(q, q, ql);
            quat_rot(off, m.jnt_pos + 3 * j, q);
            for (int k = 0; k < 3; k++) p[k] = anchor[k] - off[k];
          }
          for (int k = 0; k < 3; k++) { xanchor[3 * j + k] = anchor[k]; xaxis[3 * j + k] = axis[k]; }
        }
      }
      normalize4(q);
      for (int k = 0; k < 3; k++) xpos[3 * i + k] = p[k];
      for (int k = 0; k < 4; k++) xquat[4 * i + k] = q[k];
      quat_to_mat(xmat + 9 * i, q);
      // inertial frame
      T qi[4];
      mat_vec(xipos + 3 * i, xmat + 9 * i, m.body_ipos + 3 * i);
      for (int k = 0; k < 3; k++) xipos[3 * i + k] += p[k];
      quat_mul(qi, q, m.body_iquat + 4 * i);
      quat_to_mat(ximat + 9 * i, qi);
    }
    for (int g = 0; g < m.ngeom; g++) {
      const int b = m.geom_bodyid[g];
      T q[4];
      mat_vec(geom_xpos + 3 * g, xmat + 9 * b, m.geom_pos + 3 * g);
      for (int k = 0; k < 3; k++) geom_xpos[3 * g + k] += xpos[3 * b + k];
      quat_mul(q, xquat + 4 * b, m.geom_quat + 4 * g);
      quat_to_mat(geom_xmat + 9 * g, q);
    }
    for (int s = 0; s < m.nsite; s++) {
      const int b = m.site_bodyid[s];
      T q[4];
      mat_vec(site_xpos + 3 * s, xmat + 9 * b, m.site_pos + 3 * s);
      for (int k = 0; k < 3; k++) site_xpos[3 * s + k] += xpos[3 * b + k];
      quat_mul(q, xquat + 4 * b, m.site_quat + 4 * s);
      quat_to_mat(site_xmat + 9 * s, q);
    }
  }

  // subtree centres of mass, com-frame inertias, motion axes
  __device__ void com_frame() {
    const int nb = m.nbody;
    for (int k = 0; k < 3 * nb; k++) com[k] = 0;
    for (int i = nb - 1; i >= 0; i--) {
      for (int k = 0; k < 3; k++) com[3 * i + k] += xipos[3 * i + k] * m.body_mass[i];
      if (i) { const int p = m.body_parentid[i]; for (int k = 0; k < 3; k++) com[3 * p + k] += com[3 * i + k]; }
      if (m.body_subtreemass[i] < Num<T>::minval()) { for (int k = 0; k < 3; k++) com[3 * i + k] = xipos[3 * i + k]; }
      else { const T inv = T(1) / tmax(Num<T>::minval(), m.body_subtreemass[i]); for (int k = 0; k < 3; k++) com[3 * i + k] *= inv; }
    }
    for (int k = 0; k < 10; k++) cinert[k] = 0;
    for (int i = 1; i < nb; i++) {
      T off[3];
      const int r = m.body_rootid[i];
      for (int k = 0; k < 3; k++) off[k] = xipos[3 * i + k] - com[3 * r + k];
      inert_about(cinert + 10 * i, m.body_inertia + 3 * i, ximat + 9 * i, off, m.body_mass[i]);
    }
    for (int j = 0; j < m.njnt; j++) {
      const int b = m.jnt_bodyid[j], r = m.body_rootid[b];
      T* cd = cdof + 6 * m.jnt_dofadr[j];
      T off[3];
      for (int k = 0; k < 3; k++) off[k] = com[3 * r + k] - xanchor[3 * j + k];
      const int t = m.jnt_type[j];
      if (t == JNT_FREE) {
        for (int k = 0; k < 18; k++) cd[k] = 0;
        cd[3] = 1; cd[10] = 1; cd[17] = 1;
        for (int a = 0; a < 3; a++) {
          T ax[3] = {xmat[9 * b + a], xmat[9 * b + a + 3], xmat[9 * b + a + 6]};
          T* c = cd + 18 + 6 * a;
          c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
          cross3(c + 3, ax, off);
        }
      } else if (t == JNT_SLIDE) {
        cd[0] = cd[1] = cd[2] = 0;
        for (int k = 0; k < 3; k++) cd[3 + k] = xaxis[3 * j + k];
      } else {
        for (int k = 0; k < 3; k++) cd[k] = xaxis[3 * j + k];
        cross3(cd + 3, xaxis + 3 * j, off);
      }
    }
  }

  __device__ void tendons() {
    const int nv = m.nv;
    for (int t = 0; t < m.ntendon; t++) {
      T L = 0;
      for (int k = 0; k < nv; k++) ten_J[t * nv + k] = 0;
      for (int w = m.tendon_adr[t]; w < m.tendon_adr[t] + m.tendon_num[t]; w++) {
        const int j = m.wrap_jntid[w];
        L += m.wrap_coef[w] * qpos[m.jnt_qposadr[j]];
        ten_J[t * nv + m.jnt_dofadr[j]] = m.wrap_coef[w];
      }
      ten_len[t] = L;
    }
  }

  // composite rigid body algorithm -> dense symmetric M
  __device__ void mass_matrix() {
    const int nb = m.nbody, nv = m.nv;
    T* crb = spat2;
    for (int k = 0; k < 10 * nb; k++) crb[k] = cinert[k];
    for (int i = nb - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      if (p > 0) for (int k = 0; k < 10; k++) crb[10 * p + k] += crb[10 * i + k];
    }
    for (int k = 0; k < nv * nv; k++) M[k] = 0;
    for (int i = 0; i < nv; i++) {
      T buf[6];
      M[i * nv + i] = m.dof_armature[i];
      inert_mul(buf, crb + 10 * m.dof_bodyid[i], cdof + 6 * i);
      for (int j = i; j >= 0; j = m.dof_parentid[j]) {
        T s = 0;
        for (int k = 0; k < 6; k++) s += cdof[6 * j + k] * buf[k];
        M[i * nv + j] += s;
        if (j != i) M[j * nv + i] = M[i * nv + j];
      }
    }
  }

  // in-place L'DL factorisation of the matrix held in LD (tree sparsity), dinv <- 1/D
  __device__ void factor_LD() {
    const int nv = m.nv;
    for (int k = nv - 1; k >= 0; k--) {
      const T dkk = LD[k * nv + k];
      for (int i = m.dof_parentid[k]; i >= 0; i = m.dof_parentid[i]) {
        const T t = LD[k * nv + i] / dkk;
        for (int j = i; j >= 0; j = m.dof_parentid[j]) LD[i * nv + j] -= t * LD[k * nv + j];
        LD[k * nv + i] = t;
      }
      dinv[k] = T(1) / dkk;
    }
  }
  __device__ void solve(T* x) const {
    const int nv = m.nv;
    for (int i = nv - 1; i >= 0; i--)
      for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) x[j] -= LD[i * nv + j] * x[i];
    for (int i = 0; i < nv; i++) x[i] *= dinv[i];
    for (int i = 0; i < nv; i++)
      for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) x[i] -= LD[i * nv + j] * x[j];
  }
  __device__ void mul_M(T* r, const T* v) const {
    const int nv = m.nv;
    for (int i = 0; i < nv; i++) r[i] = 0;
    for (int i = 0; i < nv; i++) {
      r[i] += M[i * nv + i] * v[i];
      for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) {
        r[i] += M[i * nv + j] * v[j];
        r[j] += M[i * nv + j] * v[i];
      }
    }
  }

  // point Jacobian columns: calls f(dof, jp[3], jr[3]) for every dof that moves `body`
  template <class F>
  __device__ void for_jac(int body, const T* point, F f) const {
    T off[3];
    const int r = m.body_rootid[body];
    for (int k = 0; k < 3; k++) off[k] = point[k] - com[3 * r + k];
    while (body && !m.body_dofnum[body]) body = m.body_parentid[body];
    if (!body) return;
    for (int i = m.body_dofadr[body] + m.body_dofnum[body] - 1; i >= 0; i = m.dof_parentid[i]) {
      const T* c = cdof + 6 * i;
      T jp[3];
      cross3(jp, c, off);
      jp[0] += c[3]; jp[1] += c[4]; jp[2] += c[5];
      f(i, jp, c);
    }
  }

  __device__ void transmission() {
    const int nv = m.nv;
    for (int a = 0; a < m.nu; a++) {
      T* mom = act_moment + a * nv;
      for (int k = 0; k < nv; k++) mom[k] = 0;
      const int id = m.actuator_trnid[a];
      const T* gear = m.actuator_gear + 6 * a;
      if (m.actuator_trntype[a] == TRN_JOINT) {
        act_len[a] = qpos[m.jnt_qposadr[id]] * gear[0];
        mom[m.jnt_dofadr[id]] = gear[0];
      } else {
        T wf[3], wt[3];
        mat_vec(wf, site_xmat + 9 * id, gear);
        mat_vec(wt, site_xmat + 9 * id, gear + 3);
        for_jac(m.site_bodyid[id], site_xpos + 3 * id, [&](int d, const T* jp, const T* jr) {
          mom[d] = jp[0] * wf[0] + jp[1] * wf[1] + jp[2] * wf[2] + jr[0] * wt[0] + jr[1] * wt[1] + jr[2] * wt[2];
        });
        act_len[a] = 0;
      }
    }
  }

  // ------------------------------------------------------------------ collision
  __device__ bool add_contact(int pair, T dist, const T* pos, const T* normal, const T* tangent_hint) {
    if (ncon >= D::NCON) { flags |= 8; return false; }
    const int c = ncon++;
    con_dist[c] = dist; con_pair[c] = (short)pair;
    for (int k = 0; k < 3; k++) { con_pos[3 * c + k] = pos[k]; con_frame[9 * c + k] = normal[k]; con_frame[9 * c + 3 + k] = tangent_hint ? tangent_hint[k] : T(0); }
    make_frame(con_frame + 9 * c);
    return true;
  }
  __device__ void plane_sphere(int pair, T margin, const T* ppos, const T* normal, const T* spos, T radius, const T* hint) {
    T d[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
    const T cd = dot3(d, normal);
    if (cd > margin + radius) return;
    const T dist = cd - radius;
    const T s = -dist / 2 - radius;
    T pos[3] = {spos[0] + normal[0] * s, spos[1] + normal[1] * s, spos[2] + normal[2] * s};
    add_contact(pair, dist, pos, normal, hint);
  }
  __device__ int sphere_sphere(int pair, T margin, const T* p1, const T* z1, T r1, const T* p2, const T* z2, T r2) {
    T n[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const T dsq = dot3(n, n), lim = margin + r1 + r2;
    if (dsq > lim * lim) return 0;
    const T cd = normalize3(n);
    const T dist = cd - r1 - r2;
    if (cd < Num<T>::minval()) { cross3(n, z1, z2); normalize3(n); }
    const T s = r1 + T(0.5) * dist;
    T pos[3] = {p1[0] + n[0] * s, p1[1] + n[1] * s, p1[2] + n[2] * s};
    return add_contact(pair, dist, pos, n, nullptr) ? 1 : 0;
  }
  __device__ void capsule_capsule(int pair, T margin, const T* p1, const T* z1, T r1, T l1, const T* p2, const T* z2, T r2, T l2) {
    T dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, v1[3], v2[3];
    const T ma = dot3(z1, z1), mb = -dot3(z1, z2), mc = dot3(z2, z2), u = -dot3(z1, dif), w = dot3(z2, dif);
    const T det = ma * mc - mb * mb;
    if (fabs(det) >= Num<T>::minval()) {
      T x1 = (mc * u - mb * w) / det, x2 = (ma * w - mb * u) / det;
      if (x1 > l1) { x1 = l1; x2 = (w - mb * l1) / mc; }
      else if (x1 < -l1) { x1 = -l1; x2 = (w + mb * l1) / mc; }
      if (x2 > l2) { x2 = l2; x1 = tclip((u - mb * l2) / ma, -l1, l1); }
      else if (x2 < -l2) { x2 = -l2; x1 = tclip((u + mb * l2) / ma, -l1, l1); }
      for (int k = 0; k < 3; k++) { v1[k] = p1[k] + z1[k] * x1; v2[k] = p2[k] + z2[k] * x2; }
      sphere_sphere(pair, margin, v1, z1, r1, v2, z2, r2);
      return;
    }
    int n = 0;
    T x;
    for (int side = 1; side >= -1 && n < 2; side -= 2) {
      x = tclip((w - side * mb * l1) / mc, -l2, l2);
      for (int k = 0; k < 3; k++) { v1[k] = p1[k] + z1[k] * (side * l1); v2[k] = p2[k] + z2[k] * x; }
      n += sphere_sphere(pair, margin, v1, z1, r1, v2, z2, r2);
    }
    for (int side = 1; side >= -1 && n < 2; side -= 2) {
      x = tclip((u - side * mb * l2) / ma, -l1, l1);
      for (int k = 0; k < 3; k++) { v2[k] = p2[k] + z2[k] * (side * l2); v1[k] = p1[k] + z1[k] * x; }
      n += sphere_sphere(pair, margin, v1, z1, r1, v2, z2, r2);
    }
  }

  __device__ void collide() {
    ncon = 0;
    for (int p = 0; p < m.npair; p++) {
      const int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p];
      const T margin = m.pair_margin[p];
      const T *p1 = geom_xpos + 3 * g1, *R1 = geom_xmat + 9 * g1, *p2 = geom_xpos + 3 * g2, *R2 = geom_xmat + 9 * g2;
      const T *s1 = m.geom_size + 3 * g1, *s2 = m.geom_size + 3 * g2;
      const int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
      T z1[3] = {R1[2], R1[5], R1[8]}, z2[3] = {R2[2], R2[5], R2[8]};
      T d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      if (t1 == GEOM_PLANE) {
        const T h = dot3(d, z1);
        if (h > margin + m.geom_rbound[g2]) continue;
        if (t2 == GEOM_SPHERE) plane_sphere(p, margin, p1, z1, p2, s2[0], nullptr);
        else if (t2 == GEOM_CAPSULE) {
          T e[3];
          for (int k = 0; k < 3; k++) e[k] = p2[k] + z2[k] * s2[1];
          plane_sphere(p, margin, p1, z1, e, s2[0], z2);
          for (int k = 0; k < 3; k++) e[k] = p2[k] - z2[k] * s2[1];
          plane_sphere(p, margin, p1, z1, e, s2[0], z2);
        } else if (t2 == GEOM_BOX) {
          int cnt = 0;
          for (int c = 0; c < 8 && cnt < 4; c++) {
            T loc[3] = {(c & 1) ? s2[0] : -s2[0], (c & 2) ? s2[1] : -s2[1], (c & 4) ? s2[2] : -s2[2]}, corner[3];
            mat_vec(corner, R2, loc);
            const T ld = dot3(z1, corner);
            if (h + ld > margin || ld > 0) continue;
            const T dist = h + ld;
            T pos[3];
            for (int k = 0; k < 3; k++) pos[k] = corner[k] + p2[k] - z1[k] * (dist / 2);
            add_contact(p, dist, pos, z1, nullptr);
            cnt++;
          }
        } else if (t2 == GEOM_ELLIPSOID) {
          T neg[3] = {-z1[0], -z1[1], -z1[2]}, dl[3], sup[3];
          matT_vec(dl, R2, neg);
          T sc[3] = {s2[0] * dl[0], s2[1] * dl[1], s2[2] * dl[2]};
          const T nr = sqrt(dot3(sc, sc));
          if (nr < Num<T>::minval()) continue;
          T loc[3] = {s2[0] * sc[0] / nr, s2[1] * sc[1] / nr, s2[2] * sc[2] / nr};
          mat_vec(sup, R2, loc);
          for (int k = 0; k < 3; k++) sup[k] += p2[k];
          T rel[3] = {sup[0] - p1[0], sup[1] - p1[1], sup[2] - p1[2]};
          const T dist = dot3(rel, z1);
          if (dist > margin) continue;
          T pos[3];
          for (int k = 0; k < 3; k++) pos[k] = sup[k] - z1[k] * (T(0.5) * dist);
          add_contact(p, dist, pos, z1, nullptr);
        }
      } else {
        const T bound = margin + m.geom_rbound[g1] + m.geom_rbound[g2];
        if (dot3(d, d) > bound * bound) continue;
        if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) sphere_sphere(p, margin, p1, z1, s1[0], p2, z2, s2[0]);
        else if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) {
          T v[3] = {-d[0], -d[1], -d[2]};
          const T x = tclip(dot3(z2, v), -s2[1], s2[1]);
          for (int k = 0; k < 3; k++) v[k] = p2[k] + z2[k] * x;
          sphere_sphere(p, margin, p1, z1, s1[0], v, z2, s2[0]);
        } else if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE)
          capsule_capsule(p, margin, p1, z1, s1[0], s1[1], p2, z2, s2[0], s2[1]);
      }
    }
  }

  // ------------------------------------------------------------------ constraint rows
  __device__ int new_row(int type, int id, T pos, T margin) {
    if (nefc >= D::NEFC) { flags |= 8; return -1; }
    const int r = nefc++;
    row_type[r] = (signed char)type; row_id[r] = (short)id; row_pos[r] = pos; row_margin[r] = margin;
    for (int k = 0; k < m.nv; k++) J[r * m.nv + k] = 0;
    return r;
  }
  __device__ void make_rows() {
    const int nv = m.nv;
    nefc = 0;
    for (int j = 0; j < m.njnt; j++) {
      if (!m.jnt_limited[j] || m.jnt_type[j] < JNT_SLIDE) continue;
      const T value = qpos[m.jnt_qposadr[j]], margin = m.jnt_margin[j];
      for (int side = -1; side <= 1; side += 2) {
        const T dist = side * (m.jnt_range[2 * j + (side + 1) / 2] - value);
        if (dist < margin) { const int r = new_row(ROW_LIMIT_JOINT, j, dist, margin); if (r >= 0) J[r * nv + m.jnt_dofadr[j]] = T(-side); }
      }
    }
    for (int t = 0; t < m.ntendon; t++) {
      if (!m.tendon_limited[t]) continue;
      const T value = ten_len[t], margin = m.tendon_margin[t];
      for (int side = -1; side <= 1; side += 2) {
        const T dist = side * (m.tendon_range[2 * t + (side + 1) / 2] - value);
        if (dist < margin) { const int r = new_row(ROW_LIMIT_TENDON, t, dist, margin); if (r >= 0) for (int k = 0; k < nv; k++) J[r * nv + k] = -side * ten_J[t * nv + k]; }
      }
    }
    for (int c = 0; c < ncon; c++) {
      const int p = con_pair[c];
      const T incl = m.pair_margin[p] - m.pair_gap[p];
      if (con_dist[c] >= incl) continue;
      const int b1 = m.geom_bodyid[m.pair_geom1[p]], b2 = m.geom_bodyid[m.pair_geom2[p]];
      const T* fr = con_frame + 9 * c;
      const int dim = m.pair_dim[p];
      const T mu = m.pair_friction[2 * p];
      int rows[4];
      const int nrow = dim == 1 ? 1 : 4;
      bool ok = true;
      for (int k = 0; k < nrow; k++) { rows[k] = new_row(dim == 1 ? ROW_CONTACT_1 : ROW_CONTACT_PYR, c, con_dist[c], incl); ok = ok && rows[k] >= 0; }
      if (!ok) break;
      // rows hold frame * (jac(b2) - jac(b1)); pyramid edges are normal +- mu * tangent
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        for_jac(sgn < 0 ? b1 : b2, con_pos + 3 * c, [&](int d, const T* jp, const T*) {
          const T jn = sgn * dot3(fr, jp);
          if (dim == 1) { J[rows[0] * nv + d] += jn; return; }
          const T j1 = sgn * dot3(fr + 3, jp), j2 = sgn * dot3(fr + 6, jp);
          J[rows[0] * nv + d] += jn + mu * j1; J[rows[1] * nv + d] += jn - mu * j1;
          J[rows[2] * nv + d] += jn + mu * j2; J[rows[3] * nv + d] += jn - mu * j2;
        });
      }
    }
  }

  static __device__ T impedance(const T* si, T pos, T margin) {
    const T lo = T(0.0001), hi = T(0.9999);
    const T dmin = tclip(si[0], lo, hi), dmax = tclip(si[1], lo, hi), width = tmax(Num<T>::minval(), si[2]);
    const T mid = tclip(si[3], lo, hi), power = tmax(T(1), si[4]);
    if (dmin == dmax || width <= Num<T>::minval()) return T(0.5) * (dmin + dmax);
    T x = (pos - margin) / width;
    if (x < 0) x = -x;
    if (x >= 1) return dmax;
    if (x == 0) return dmin;
    T y;
    if (power == 1) y = x;
    else if (x <= mid) y = (T(1) / pow(mid, power - 1)) * pow(x, power);
    else y = T(1) - (T(1) / pow(T(1) - mid, power - 1)) * pow(T(1) - x, power);
    return dmin + y * (dmax - dmin);
  }

  // impedance, regulariser D = 1/R and reference acceleration of every row
  __device__ void row_params() {
    const int nv = m.nv;
    int within = 0;  // row index inside the current pyramidal contact
    T Rpy = 0;
    for (int i = 0; i < nefc; i++) {
      const T *sr, *si;
      T diag;
      const int id = row_id[i], type = row_type[i];
      if (type == ROW_LIMIT_JOINT) { sr = m.jnt_solref + 2 * id; si = m.jnt_solimp + 5 * id; diag = m.dof_invweight0[m.jnt_dofadr[id]]; }
      else if (type == ROW_LIMIT_TENDON) { sr = m.tendon_solref + 2 * id; si = m.tendon_solimp + 5 * id; diag = m.tendon_invweight0[id]; }
      else {
        const int p = con_pair[id];
        sr = m.pair_solref + 2 * p; si = m.pair_solimp + 5 * p;
        const T tran = m.body_invweight0[2 * m.geom_bodyid[m.pair_geom1[p]]] + m.body_invweight0[2 * m.geom_bodyid[m.pair_geom2[p]]];
        const T mu = m.pair_friction[2 * p];
        diag = (type == ROW_CONTACT_1) ? tran : tran + mu * mu * tran;
      }
      const T pos = row_pos[i], margin = row_margin[i];
      const T imp = impedance(si, pos, margin);
      const T dmax = tclip(si[1], T(0.0001), T(0.9999));
      T K, B;
      if (sr[0] > 0) {
        const T tc = tmax(sr[0], 2 * m.timestep), dr = sr[1];
        K = T(1) / tmax(Num<T>::minval(), dmax * dmax * tc * tc * dr * dr);
        B = T(2) / tmax(Num<T>::minval(), dmax * tc);
      } else {
        K = -sr[0] / tmax(Num<T>::minval(), dmax * dmax);
        B = -sr[1] / tmax(Num<T>::minval(), dmax);
      }
      T R = tmax(Num<T>::minval(), (T(1) - imp) * diag / imp);
      if (type == ROW_CONTACT_PYR) {
        if (within == 0) { const T mu = m.pair_friction[2 * con_pair[id]]; Rpy = 2 * mu * mu * R; }
        R = Rpy;
        within = (within + 1) & 3;
      }
      row_D[i] = T(1) / R;
      T vel = 0;
      for (int k = 0; k < nv; k++) vel += J[i * nv + k] * qvel[k];
      row_aref[i] = -B * vel - K * imp * (pos - margin);
    }
  }

  // ------------------------------------------------------------------ velocity stage
  __device__ void velocities() {
    for (int k = 0; k < 6; k++) cvel[k] = 0;
    for (int i = 1; i < m.nbody; i++) {
      T v[6];
      const int da = m.body_dofadr[i], dn = m.body_dofnum[i], p = m.body_parentid[i];
      for (int k = 0; k < 6; k++) v[k] = cvel[6 * p + k];
      int j = 0;
      while (j < dn) {
        if (m.jnt_type[m.dof_jntid[da + j]] == JNT_FREE) {
          for (int k = 0; k < 18; k++) cdof_dot[6 * da + k] = 0;
          for (int k = 0; k < 6; k++) {
            T t = 0;
            for (int q = 0; q < 3; q++) t += cdof[6 * (da + q) + k] * qvel[da + q];
            v[k] += t;
          }
          for (int q = 3; q < 6; q++) cross_motion(cdof_dot + 6 * (da + q), v, cdof + 6 * (da + q));
          for (int k = 0; k < 6; k++) {
            T t = 0;
            for (int q = 3; q < 6; q++) t += cdof[6 * (da + q) + k] * qvel[da + q];
            v[k] += t;
          }
          j += 6;
        } else {
          cross_motion(cdof_dot + 6 * (da + j), v, cdof + 6 * (da + j));
          for (int k = 0; k < 6; k++) v[k] += cdof[6 * (da + j) + k] * qvel[da + j];
          j += 1;
        }
      }
      for (int k = 0; k < 6; k++) cvel[6 * i + k] = v[k];
    }
  }

  __device__ void fluid_body(int i) {
    const T* I = m.body_inertia + 3 * i;
    const T mass = m.body_mass[i];
    T box[3], lv[6], lf[6] = {0, 0, 0, 0, 0, 0}, wf[6];
    box[0] = sqrt(tmax(Num<T>::minval(), I[1] + I[2] - I[0]) / mass * T(6));
    box[1] = sqrt(tmax(Num<T>::minval(), I[0] + I[2] - I[1]) / mass * T(6));
    box[2] = sqrt(tmax(Num<T>::minval(), I[0] + I[1] - I[2]) / mass * T(6));
    // com-frame velocity -> local velocity at the inertial frame, minus wind
    const int r = m.body_rootid[i];
    T dif[3], tmp[3], lin[3];
    for (int k = 0; k < 3; k++) dif[k] = xipos[3 * i + k] - com[3 * r + k];
    cross3(tmp, dif, cvel + 6 * i);
    for (int k = 0; k < 3; k++) lin[k] = cvel[6 * i + 3 + k] - tmp[k];
    matT_vec(lv, ximat + 9 * i, cvel + 6 * i);
    matT_vec(lv + 3, ximat + 9 * i, lin);
    T lw[3];
    matT_vec(lw, ximat + 9 * i, m.wind);
    for (int k = 0; k < 3; k++) lv[3 + k] -= lw[k];
    if (m.viscosity > 0) {
      const T diam = (box[0] + box[1] + box[2]) / T(3);
      const T ca = -Num<T>::pi() * diam * diam * diam * m.viscosity, cl = T(-3) * Num<T>::pi() * diam * m.viscosity;
      for (int k = 0; k < 3; k++) { lf[k] = lv[k] * ca; lf[3 + k] = lv[3 + k] * cl; }
    }
    if (m.density > 0) {
      lf[3] -= T(0.5) * m.density * box[1] * box[2] * fabs(lv[3]) * lv[3];
      lf[4] -= T(0.5) * m.density * box[0] * box[2] * fabs(lv[4]) * lv[4];
      lf[5] -= T(0.5) * m.density * box[0] * box[1] * fabs(lv[5]) * lv[5];
      const T b0 = box[0] * box[0], b1 = box[1] * box[1], b2 = box[2] * box[2];
      lf[0] -= m.density * box[0] * (b1 * b1 + b2 * b2) * fabs(lv[0]) * lv[0] / T(64);
      lf[1] -= m.density * box[1] * (b0 * b0 + b2 * b2) * fabs(lv[1]) * lv[1] / T(64);
      lf[2] -= m.density * box[2] * (b0 * b0 + b1 * b1) * fabs(lv[2]) * lv[2] / T(64);
    }
    mat_vec(wf, ximat + 9 * i, lf);
    mat_vec(wf + 3, ximat + 9 * i, lf + 3);
    for_jac(i, xipos + 3 * i, [&](int d, const T* jp, const T* jr) {
      f_passive[d] += jp[0] * wf[3] + jp[1] * wf[4] + jp[2] * wf[5] + jr[0] * wf[0] + jr[1] * wf[1] + jr[2] * wf[2];
    });
  }

  __device__ void passive_forces() {
    const int nv = m.nv;
    for (int k = 0; k < nv; k++) f_passive[k] = 0;
    for (int j = 0; j < m.njnt; j++) {
      const T st = m.jnt_stiffness[j];
      if (st == 0) continue;
      const int pa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
      if (m.jnt_type[j] >= JNT_SLIDE) f_passive[da] -= st * (qpos[pa] - m.qpos_spring[pa]);
      else if (m.jnt_type[j] == JNT_FREE) {
        T q[4], qs[4] = {m.qpos_spring[pa + 3], -m.qpos_spring[pa + 4], -m.qpos_spring[pa + 5], -m.qpos_spring[pa + 6]}, qd[4], dv[3];
        for (int k = 0; k < 3; k++) f_passive[da + k] -= st * (qpos[pa + k] - m.qpos_spring[pa + k]);
        for (int k = 0; k < 4; k++) q[k] = qpos[pa + 3 + k];
        normalize4(q);
        quat_mul(qd, qs, q);
        quat_to_vel(dv, qd, T(1));
        for (int k = 0; k < 3; k++) f_passive[da + 3 + k] -= st * dv[k];
      }
    }
    for (int k = 0; k < nv; k++) f_passive[k] -= m.dof_damping[k] * qvel[k];
    for (int t = 0; t < m.ntendon; t++) {
      const T st = m.tendon_stiffness[t], dm = m.tendon_damping[t];
      if (st == 0 && dm == 0) continue;
      T frc = 0, vel = 0;
      const T lo = m.tendon_lengthspring[2 * t], hi = m.tendon_lengthspring[2 * t + 1], L = ten_len[t];
      if (L > hi) frc = st * (hi - L); else if (L < lo) frc = st * (lo - L);
      for (int k = 0; k < nv; k++) vel += ten_J[t * nv + k] * qvel[k];
      frc -= dm * vel;
      for (int k = 0; k < nv; k++) f_passive[k] += ten_J[t * nv + k] * frc;
    }
    if (m.has_fluid)
      for (int i = 1; i < m.nbody; i++)
        if (m.body_mass[i] >= Num<T>::minval()) fluid_body(i);
  }

  // recursive Newton-Euler without accelerations: Coriolis, centrifugal, gravity
  __device__ void bias_forces() {
    const int nb = m.nbody, nv = m.nv;
    T* cacc = spat;      // 6 per body, overwritten by cfrc as we go
    T* cfrc = spat2;     // 6 per body (spat2 holds 10 per body)
    cacc[0] = cacc[1] = cacc[2] = 0;
    cacc[3] = -m.gravity[0]; cacc[4] = -m.gravity[1]; cacc[5] = -m.gravity[2];
    for (int k = 0; k < 6; k++) cfrc[k] = 0;
    for (int i = 1; i < nb; i++) {
      const int da = m.body_dofadr[i], p = m.body_parentid[i];
      T t[6] = {0, 0, 0, 0, 0, 0}, t1[6];
      for (int j = 0; j < m.body_dofnum[i]; j++)
        for (int k = 0; k < 6; k++) t[k] += cdof_dot[6 * (da + j) + k] * qvel[da + j];
      for (int k = 0; k < 6; k++) cacc[6 * i + k] = cacc[6 * p + k] + t[k];
      inert_mul(cfrc + 6 * i, cinert + 10 * i, cacc + 6 * i);
      inert_mul(t, cinert + 10 * i, cvel + 6 * i);
      cross_force(t1, cvel + 6 * i, t);
      for (int k = 0; k < 6; k++) cfrc[6 * i + k] += t1[k];
    }
    for (int i = nb - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      if (p) for (int k = 0; k < 6; k++) cfrc[6 * p + k] += cfrc[6 * i + k];
    }
    for (int d = 0; d < nv; d++) {
      T s = 0;
      const int b = m.dof_bodyid[d];
      for (int k = 0; k < 6; k++) s += cdof[6 * d + k] * cfrc[6 * b + k];
      f_bias[d] = s;
    }
  }

  // ------------------------------------------------------------------ smooth acceleration
  __device__ void smooth_dynamics() {
    const int nv = m.nv;
    T* f_act = grad;  // scratch: generalized actuator force
    for (int k = 0; k < nv; k++) f_act[k] = 0;
    for (int a = 0; a < m.nu; a++) {
      T u = ctrl[a];
      if (m.actuator_ctrllimited[a]) u = tclip(u, m.actuator_ctrlrange[2 * a], m.actuator_ctrlrange[2 * a + 1]);
      const T* bp = m.actuator_biasprm + 3 * a;
      T vel = 0;
      for (int k = 0; k < nv; k++) vel += act_moment[a * nv + k] * qvel[k];
      T force = m.actuator_gainprm[a] * u + bp[0] + bp[1] * act_len[a] + bp[2] * vel;
      if (m.actuator_forcelimited[a]) force = tclip(force, m.actuator_forcerange[2 * a], m.actuator_forcerange[2 * a + 1]);
      if (m.actuator_disabled[a]) force = 0;
      for (int k = 0; k < nv; k++) f_act[k] += act_moment[a * nv + k] * force;
    }
    for (int k = 0; k < nv; k++) {
      f_smooth[k] = f_passive[k] - f_bias[k];
      f_smooth[k] += f_act[k];
      a_smooth[k] = f_smooth[k];
    }
    solve(a_smooth);
  }

  // ------------------------------------------------------------------ Newton solver
  // constraint cost at Jaref; optionally refresh f_con = J' * force
  __device__ T row_cost(const T* jar, bool write_force) {
    const int nv = m.nv;
    T c = 0;
    if (write_force) for (int k = 0; k < nv; k++) f_con[k] = 0;
    for (int i = 0; i < nefc; i++) {
      if (jar[i] >= 0) continue;
      c += T(0.5) * row_D[i] * jar[i] * jar[i];
      if (write_force) { const T f = -row_D[i] * jar[i]; for (int k = 0; k < nv; k++) f_con[k] += J[i * nv + k] * f; }
    }
    return c;
  }
  struct LsPoint { T alpha, cost, d1, d2; };
  __device__ void ls_eval(T alpha, LsPoint& p) {
    ls_iter++;
    T q0 = qg0, q1 = qg1, q2 = qg2;
    for (int i = 0; i < nefc; i++) {
      if (Jaref[i] + alpha * Jv[i] < 0) {
        const T dj = row_D[i] * Jaref[i];
        q0 += T(0.5) * Jaref[i] * dj; q1 += Jv[i] * dj; q2 += T(0.5) * Jv[i] * row_D[i] * Jv[i];
      }
    }
    p.alpha = alpha; p.cost = alpha * alpha * q2 + alpha * q1 + q0; p.d1 = 2 * alpha * q2 + q1; p.d2 = 2 * q2;
    if (p.d2 <= 0) p.d2 = Num<T>::minval();
  }
  __device__ int ls_bracket(LsPoint& p, const LsPoint* cand, LsPoint& pnext) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = 1; }
      else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = 2; }
    }
    if (flag) ls_eval(p.alpha - p.d1 / p.d2, pnext);
    return flag;
  }
  // exact 1-D minimisation of the piecewise-quadratic cost along `search`
  __device__ T line_search() {
    const int nv = m.nv;
    LsPoint p0, p1, p2, pmid, p1n, p2n;
    ls_iter = 0;
    T sn = 0;
    for (int k = 0; k < nv; k++) sn += search[k] * search[k];
    sn = sqrt(sn);
    if (sn < Num<T>::minval()) return 0;
    const T scale = T(1) / (m.meaninertia * T(nv > 1 ? nv : 1));
    const T gtol = m.tolerance * m.ls_tolerance * sn / scale;
    mul_M(Mv, search);
    for (int i = 0; i < nefc; i++) { T s = 0; for (int k = 0; k < nv; k++) s += J[i * nv + k] * search[k]; Jv[i] = s; }
    T a = 0, b = 0, e = 0;
    for (int k = 0; k < nv; k++) { a += search[k] * Ma[k]; b += f_smooth[k] * search[k]; e += search[k] * Mv[k]; }
    qg0 = gauss; qg1 = a - b; qg2 = T(0.5) * e;
    ls_eval(0, p0);
    ls_eval(p0.alpha - p0.d1 / p0.d2, p1);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.d1) < gtol) return p1.alpha;
    const int dir = p1.d1 < 0 ? 1 : -1;
    bool p2up = false;
    const int maxls = m.ls_iterations;
    while (p1.d1 * dir <= -gtol && ls_iter < maxls) {
      p2 = p1; p2up = true;
      ls_eval(p1.alpha - p1.d1 / p1.d2, p1);
      if (fabs(p1.d1) < gtol) return p1.alpha;
    }
    if (ls_iter >= maxls || !p2up) return p1.alpha;
    p2n = p1;
    ls_eval(p1.alpha - p1.d1 / p1.d2, p1n);
    while (ls_iter < maxls) {
      ls_eval(T(0.5) * (p1.alpha + p2.alpha), pmid);
      LsPoint cand[3] = {p1n, p2n, pmid};
      T best = 0; int bi = -1;
      for (int i = 0; i < 3; i++)
        if (fabs(cand[i].d1) < gtol && (bi == -1 || cand[i].cost < best)) { best = cand[i].cost; bi = i; }
      if (bi >= 0) return cand[bi].alpha;
      const int b1 = ls_bracket(p1, cand, p1n), b2 = ls_bracket(p2, cand, p2n);
      if (!b1 && !b2) return pmid.alpha;
    }
    if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
    if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
    return 0;
  }
  __device__ void newton_refresh() {  // cost, forces, gradient, Newton direction at the current qacc
    const int nv = m.nv;
    cost = row_cost(Jaref, true);
    T g = 0;
    for (int k = 0; k < nv; k++) g += (Ma[k] - f_smooth[k]) * (qacc[k] - a_smooth[k]);
    gauss = T(0.5) * g;
    cost += gauss;
    for (int k = 0; k < nv; k++) grad[k] = Ma[k] - f_smooth[k] - f_con[k];
    // H = M + J' D_active J  (lower triangle in LD), dense Cholesky, Mgrad = H^-1 grad
    T* H = LD;
    for (int k = 0; k < nv * nv; k++) H[k] = M[k];
    for (int i = 0; i < nefc; i++) {
      if (Jaref[i] >= 0) continue;
      const T* Ji = J + i * nv;
      for (int r = 0; r < nv; r++) {
        if (Ji[r] == 0) continue;
        const T s = row_D[i] * Ji[r];
        for (int q = 0; q <= r; q++) H[r * nv + q] += s * Ji[q];
      }
    }
    for (int j = 0; j < nv; j++) {
      T t = H[j * nv + j];
      for (int k = 0; k < j; k++) t -= H[j * nv + k] * H[j * nv + k];
      if (t < Num<T>::minval()) t = Num<T>::minval();
      const T djj = sqrt(t);
      H[j * nv + j] = djj;
      const T inv = T(1) / djj;
      for (int i = j + 1; i < nv; i++) {
        T s = H[i * nv + j];
        for (int k = 0; k < j; k++) s -= H[i * nv + k] * H[j * nv + k];
        H[i * nv + j] = s * inv;
      }
    }
    for (int i = 0; i < nv; i++) {
      T s = grad[i];
      for (int k = 0; k < i; k++) s -= H[i * nv + k] * Mgrad[k];
      Mgrad[i] = s / H[i * nv + i];
    }
    for (int i = nv - 1; i >= 0; i--) {
      T s = Mgrad[i];
      for (int k = i + 1; k < nv; k++) s -= H[k * nv + i] * Mgrad[k];
      Mgrad[i] = s / H[i * nv + i];
    }
  }
  __device__ void constrained_acceleration() {
    const int nv = m.nv;
    niter = 0;
    if (!nefc) {
      for (int k = 0; k < nv; k++) { qacc[k] = a_smooth[k]; warm[k] = a_smooth[k]; f_con[k] = 0; }
      return;
    }
    // warm start: keep qacc_warmstart only if it is cheaper than the unconstrained acceleration
    for (int k = 0; k < nv; k++) qacc[k] = warm[k];
    for (int i = 0; i < nefc; i++) { T s = 0; for (int k = 0; k < nv; k++) s += J[i * nv + k] * qacc[k]; Jaref[i] = s - row_aref[i]; }
    T cw = row_cost(Jaref, false);
    mul_M(Ma, qacc);
    for (int k = 0; k < nv; k++) cw += T(0.5) * (Ma[k] - f_smooth[k]) * (qacc[k] - a_smooth[k]);
    for (int i = 0; i < nefc; i++) { T s = 0; for (int k = 0; k < nv; k++) s += J[i * nv + k] * a_smooth[k]; Jv[i] = s - row_aref[i]; }
    const T cs = row_cost(Jv, false);
    if (cw > cs) {
      for (int k = 0; k < nv; k++) qacc[k] = a_smooth[k];
      for (int i = 0; i < nefc; i++) Jaref[i] = Jv[i];
      mul_M(Ma, qacc);
    }
    newton_refresh();
    for (int k = 0; k < nv; k++) search[k] = -Mgrad[k];
    const T scale = T(1) / (m.meaninertia * T(nv > 1 ? nv : 1));
    while (niter < m.iterations) {
      const T alpha = line_search();
      if (alpha == 0) break;
      for (int k = 0; k < nv; k++) { qacc[k] += alpha * search[k]; Ma[k] += alpha * Mv[k]; }
      for (int i = 0; i < nefc; i++) Jaref[i] += alpha * Jv[i];
      const T old = cost;
      newton_refresh();
      T gn = 0;
      for (int k = 0; k < nv; k++) gn += grad[k] * grad[k];
      niter++;
      if (scale * (old - cost) < m.tolerance || scale * sqrt(gn) < m.tolerance) break;
      for (int k = 0; k < nv; k++) search[k] = -Mgrad[k];
    }
    for (int k = 0; k < nv; k++) warm[k] = qacc[k];
  }

  // ------------------------------------------------------------------ forward + integrators
  __device__ void forward() {
    kinematics();
    com_frame();
    tendons();
    mass_matrix();
    for (int k = 0; k < m.nv * m.nv; k++) LD[k] = M[k];
    factor_LD();
    collide();
    make_rows();
    transmission();
    velocities();
    passive_forces();
    row_params();
    bias_forces();
    smooth_dynamics();
    constrained_acceleration();
  }
  __device__ void integrate_pos(T* q, const T* v, T dt) const {
    for (int j = 0; j < m.njnt; j++) {
      const int pa = m.jnt_qposadr[j], va = m.jnt_dofadr[j];
      if (m.jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; k++) q[pa + k] += dt * v[va + k];
        quat_integrate(q + pa + 3, v + va + 3, dt);
      } else q[pa] += dt * v[va];
    }
  }
  __device__ void differentiate_pos(T* out, T dt, const T* q1, const T* q2) const {
    for (int j = 0; j < m.njnt; j++) {
      const int pa = m.jnt_qposadr[j], va = m.jnt_dofadr[j];
      if (m.jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; k++) out[va + k] = (q2[pa + k] - q1[pa + k]) / dt;
        T neg[4] = {q1[pa + 3], -q1[pa + 4], -q1[pa + 5], -q1[pa + 6]}, dq[4];
        quat_mul(dq, neg, q2 + pa + 3);
        quat_to_vel(out + va + 3, dq, dt);
      } else out[va] = (q2[pa] - q1[pa]) / dt;
    }
  }
  __device__ void check_state() {
    for (int k = 0; k < m.nq; k++) if (!(fabs(qpos[k]) <= T(1e10))) flags |= 1;
    for (int k = 0; k < m.nv; k++) if (!(fabs(qvel[k]) <= T(1e10))) flags |= 2;
  }
  __device__ void euler() {
    const int nv = m.nv;
    const T h = m.timestep;
    T* acc = grad;
    if (!m.has_dofdamping) { for (int k = 0; k < nv; k++) acc[k] = qacc[k]; }
    else {
      // (M + h*diag(damping)) acc = f_smooth + f_con   (implicit in joint damping)
      for (int k = 0; k < nv * nv; k++) LD[k] = M[k];
      for (int k = 0; k < nv; k++) LD[k * nv + k] += h * m.dof_damping[k];
      factor_LD();
      for (int k = 0; k < nv; k++) acc[k] = f_smooth[k] + f_con[k];
      solve(acc);
    }
    for (int k = 0; k < nv; k++) qvel[k] += acc[k] * h;
    integrate_pos(qpos, qvel, h);
  }
  __device__ void rk4() {
    const int nq = m.nq, nv = m.nv;
    const T h = m.timestep;
    const T A[9] = {T(0.5), 0, 0, 0, T(0.5), 0, 0, 0, 1}, B[4] = {T(1) / 6, T(1) / 3, T(1) / 3, T(1) / 6};
    T X[4][D::NQ + D::NV], F[4][D::NV], dX[2 * D::NV];
    for (int k = 0; k < nq; k++) X[0][k] = qpos[k];
    for (int k = 0; k < nv; k++) { X[0][nq + k] = qvel[k]; F[0][k] = qacc[k]; }
    for (int i = 1; i < 4; i++) {
      for (int k = 0; k < 2 * nv; k++) dX[k] = 0;
      for (int j = 0; j < i; j++) {
        const T a = A[(i - 1) * 3 + j];
        for (int k = 0; k < nv; k++) { dX[k] += a * X[j][nq + k]; dX[nv + k] += a * F[j][k]; }
      }
      for (int k = 0; k < nq + nv; k++) X[i][k] = X[0][k];
      integrate_pos(X[i], dX, h);
      for (int k = 0; k < nv; k++) X[i][nq + k] += h * dX[nv + k];
      for (int k = 0; k < nq; k++) qpos[k] = X[i][k];
      for (int k = 0; k < nv; k++) qvel[k] = X[i][nq + k];
      forward();
      for (int k = 0; k < nv; k++) F[i][k] = qacc[k];
    }
    for (int k = 0; k < 2 * nv; k++) dX[k] = 0;
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < nv; k++) { dX[k] += B[j] * X[j][nq + k]; dX[nv + k] += B[j] * F[j][k]; }
    for (int k = 0; k < nq; k++) qpos[k] = X[0][k];
    for (int k = 0; k < nv; k++) qvel[k] = X[0][nq + k] + dX[nv + k] * h;
    integrate_pos(qpos, dX, h);
  }
  // one mj_step; when `snapshot` is set it is invoked after the pre-integration forward pass
  template <class Snap>
  __device__ void step(Snap snapshot) {
    check_state();
    forward();
    for (int k = 0; k < m.nv; k++) if (!(fabs(qacc[k]) <= T(1e10))) flags |= 4;
    snapshot();
    if (m.integrator == 1) rk4(); else euler();
  }
  __device__ void step() { step([] {}); }
};

}  // namespace b2
