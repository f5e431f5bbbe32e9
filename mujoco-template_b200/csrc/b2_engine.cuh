// Per-env rigid-body engine ("lane engine": one env per thread).
//
// Computes what the reference obtains from mj.mj_step / mj.mj_forward
// (reference mujoco_template/model.py:53-57): forward kinematics, composite-rigid-body
// mass matrix, L'DL factorisation, bias forces (RNE), passive + actuator forces, primitive
// collisions, limit / contact constraint rows, a warm-started Newton solve, and
// semi-implicit Euler or RK4 integration.
//
// The model is read only through the provider policy `M` (see b2_model_dev.cuh).  With a
// generated static provider and B2_STATIC_MODEL defined, every loop below has a compile-time
// trip count and is fully unrolled, all indices fold, and the smooth-dynamics state
// (poses, inertias, motion axes, M, forces) is register-resident; only the data-dependent
// contact / constraint-row arrays (touched only when a limit or contact is active) remain in
// local memory.  With the runtime provider the same code runs out of L1-resident local
// memory for any model of the size class.
#pragma once
#include "b2_math.cuh"
#include "b2_model_dev.cuh"

#ifdef B2_STATIC_MODEL
#define B2_UNROLL _Pragma("unroll")
#else
#define B2_UNROLL
#endif
#define B2_NOUNROLL _Pragma("unroll 1")
// stage-sized member functions: inlined under a static provider (registers), real calls in the
// generic build (the env lives in local memory there anyway; keeps compile time and code size down)
#ifdef B2_STATIC_MODEL
#define B2_STAGE __device__ __forceinline__
#else
#define B2_STAGE __device__ __noinline__
#endif

namespace b2 {

#define B2_LD3(dst, fn, base) T dst[3] = {M::fn((base)), M::fn((base) + 1), M::fn((base) + 2)}
#define B2_LD4(dst, fn, base) T dst[4] = {M::fn((base)), M::fn((base) + 1), M::fn((base) + 2), M::fn((base) + 3)}

// Data-dependent contact / constraint-row arrays.  They are indexed with runtime row and
// contact numbers, so they must live in (L1-resident) local memory; keeping them in their own
// object lets the compiler promote everything in LaneEnv to registers under a static provider.
template <typename T, class D>
struct RowStore {
  T con_dist[D::NCON], con_pos[3 * D::NCON], con_frame[9 * D::NCON];
  short con_pair[D::NCON];
  signed char row_type[D::NEFC];
  short row_id[D::NEFC];
  T J[D::NEFC * D::NV], row_pos[D::NEFC], row_margin[D::NEFC], row_D[D::NEFC], row_aref[D::NEFC];
  T Jaref[D::NEFC], Jv[D::NEFC];
};

template <typename T, class D, class M>
struct LaneEnv {
  // ---- state
  T qpos[D::NQ], qvel[D::NV], ctrl[D::NU], warm[D::NV];
  // ---- position-dependent
  T xpos[3 * D::NB], xquat[4 * D::NB], xmat[9 * D::NB], xipos[3 * D::NB], ximat[9 * D::NB];
  T xanchor[3 * D::NJ], xaxis[3 * D::NJ];
  T geom_xpos[3 * D::NG], geom_xmat[9 * D::NG], site_xpos[3 * D::NS], site_xmat[9 * D::NS];
  T com[3 * D::NB], cinert[10 * D::NB], cdof[6 * D::NV];
  T Mm[D::NV * D::NV], LD[D::NV * D::NV], dinv[D::NV];
  // Small static models keep both factorisations of the position stage -- L'DL of M and of M + h*diag(damping), the
  // Euler update's matrix -- alive across the rollouts of an FD thread that share qpos (the Newton Hessian then gets
  // its own array instead of borrowing LD).  Larger / generic models refactor per rollout: registers (local memory) are
  // the scarcer resource there.
#ifdef B2_STATIC_MODEL
  static constexpr bool kKeepFactors = D::NV <= 4;
#else
  static constexpr bool kKeepFactors = false;
#endif
  T LDe[kKeepFactors ? D::NV * D::NV : 1], dinve[kKeepFactors ? D::NV : 1], Hs[kKeepFactors ? D::NV * D::NV : 1];
  // The static models that do not keep factors (4 < nv <= 8: the drone) are short of registers instead.  k_step runs
  // them with LATE = true: the mass matrix and the actuator moments -- functions of qpos that nothing reads before the
  // acceleration stage -- are computed there, so that neither is live across collision, velocities and bias forces
  // (same inputs, same arithmetic: bit-identical).  The FD kernel keeps them in the position stage its rollouts share.
#if defined(B2_STATIC_MODEL) && !defined(B2_EARLY_POSITION)
  static constexpr bool kLateMass = !kKeepFactors;
#else
  static constexpr bool kLateMass = false;
#endif
  T ten_len[D::NT], ten_J[D::NT * D::NV], act_len[D::NU], act_moment[D::NU * D::NV];
  // ---- velocity-dependent
  T cvel[6 * D::NB], cdof_dot[6 * D::NV], spat[6 * D::NB], spat2[10 * D::NB];  // scratch: cacc / crb, cfrc
  T f_bias[D::NV], f_passive[D::NV], f_smooth[D::NV], f_con[D::NV], a_smooth[D::NV], qacc[D::NV];
  // ---- contact / constraint-row counters; the rows themselves live in RowStore
  int ncon, nefc, niter, flags;
  RowStore<T, D>& R;
  // ---- solver scratch
  T Ma[D::NV], Mv[D::NV], grad[D::NV], Mgrad[D::NV], search[D::NV];
  T cost, gauss, qg0, qg1, qg2;
  int ls_iter;

  __device__ explicit LaneEnv(RowStore<T, D>& rows) : ncon(0), nefc(0), niter(0), flags(0), R(rows) {}

  static B2_DEV bool is_anc(int i, int j) { return (((unsigned)M::dof_anc(i)) >> j) & 1u; }

  // ------------------------------------------------------------------ position stage
  B2_STAGE void kinematics() {
    xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
    quat_to_mat(xmat, xquat);
    xipos[0] = xipos[1] = xipos[2] = 0; quat_to_mat(ximat, xquat);
    B2_UNROLL
    for (int i = 1; i < M::nbody(); i++) {
      T p[3], q[4];
      const int ja = M::body_jntadr(i), jn = M::body_jntnum(i), pid = M::body_parentid(i);
      if (jn == 1 && M::jnt_type(ja) == JNT_FREE) {
        const int qa = M::jnt_qposadr(ja);
        for (int k = 0; k < 3; k++) p[k] = qpos[qa + k];
        for (int k = 0; k < 4; k++) q[k] = qpos[qa + 3 + k];
        normalize4(q);
        for (int k = 0; k < 3; k++) { xanchor[3 * ja + k] = p[k]; xaxis[3 * ja + k] = M::jnt_axis(3 * ja + k); }
      } else {
        B2_LD3(bpos, body_pos, 3 * i);
        B2_LD4(bquat, body_quat, 4 * i);
        if (pid) {
          mat_vec(p, xmat + 9 * pid, bpos);
          for (int k = 0; k < 3; k++) p[k] += xpos[3 * pid + k];
          quat_mul(q, xquat + 4 * pid, bquat);
        } else {
          for (int k = 0; k < 3; k++) p[k] = bpos[k];
          for (int k = 0; k < 4; k++) q[k] = bquat[k];
        }
        B2_UNROLL
        for (int j = ja; j < ja + jn; j++) {
          T anchor[3], axis[3];
          const int qa = M::jnt_qposadr(j);
          B2_LD3(jaxis, jnt_axis, 3 * j);
          B2_LD3(jpos, jnt_pos, 3 * j);
          quat_rot(axis, jaxis, q);
          quat_rot(anchor, jpos, q);
          for (int k = 0; k < 3; k++) anchor[k] += p[k];
          const T disp = qpos[qa] - M::qpos0(qa);
          if (M::jnt_type(j) == JNT_SLIDE) {
            for (int k = 0; k < 3; k++) p[k] += axis[k] * disp;
          } else {
            T ql[4], off[3];
            quat_axis_angle(ql, jaxis, disp);
            quat_mul(q, q, ql);
            quat_rot(off, jpos, q);
            for (int k = 0; k < 3; k++) p[k] = anchor[k] - off[k];
          }
          for (int k = 0; k < 3; k++) { xanchor[3 * j + k] = anchor[k]; xaxis[3 * j + k] = axis[k]; }
        }
      }
      normalize4(q);
      for (int k = 0; k < 3; k++) xpos[3 * i + k] = p[k];
      for (int k = 0; k < 4; k++) xquat[4 * i + k] = q[k];
      quat_to_mat(xmat + 9 * i, q);
      T qi[4];
      B2_LD3(ipos, body_ipos, 3 * i);
      B2_LD4(iquat, body_iquat, 4 * i);
      mat_vec(xipos + 3 * i, xmat + 9 * i, ipos);
      for (int k = 0; k < 3; k++) xipos[3 * i + k] += p[k];
      quat_mul(qi, q, iquat);
      quat_to_mat(ximat + 9 * i, qi);
    }
#ifndef B2_STATIC_MODEL
    // generic build: geom / site poses are materialised once per forward pass
    for (int g = 0; g < M::ngeom(); g++) compute_geom_pose(g, geom_xpos + 3 * g, geom_xmat + 9 * g);
    for (int s = 0; s < M::nsite(); s++) compute_site_pose(s, site_xpos + 3 * s, site_xmat + 9 * s);
#endif
  }
  B2_DEV void compute_geom_pose(int g, T* p, T* R) const {
    const int b = M::geom_bodyid(g);
    T q[4];
    B2_LD3(gp, geom_pos, 3 * g);
    B2_LD4(gq, geom_quat, 4 * g);
    mat_vec(p, xmat + 9 * b, gp);
    for (int k = 0; k < 3; k++) p[k] += xpos[3 * b + k];
    quat_mul(q, xquat + 4 * b, gq);
    quat_to_mat(R, q);
  }
  B2_DEV void compute_site_pose(int s, T* p, T* R) const {
    const int b = M::site_bodyid(s);
    T q[4];
    B2_LD3(sp, site_pos, 3 * s);
    B2_LD4(sq, site_quat, 4 * s);
    mat_vec(p, xmat + 9 * b, sp);
    for (int k = 0; k < 3; k++) p[k] += xpos[3 * b + k];
    quat_mul(q, xquat + 4 * b, sq);
    quat_to_mat(R, q);
  }
  // Pose of a geom / site.  Under a static provider poses are recomputed from the body pose where
  // they are consumed (a handful of FMAs) instead of staying live in registers across the stages
  // in between; the generic build reads the arrays filled by kinematics().  Same arithmetic.
  B2_DEV void geom_pose(int g, T* p, T* R) const {
#ifdef B2_STATIC_MODEL
    compute_geom_pose(g, p, R);
#else
    for (int k = 0; k < 3; k++) p[k] = geom_xpos[3 * g + k];
    for (int k = 0; k < 9; k++) R[k] = geom_xmat[9 * g + k];
#endif
  }
  B2_DEV void site_pose(int s, T* p, T* R) const {
#ifdef B2_STATIC_MODEL
    compute_site_pose(s, p, R);
#else
    for (int k = 0; k < 3; k++) p[k] = site_xpos[3 * s + k];
    for (int k = 0; k < 9; k++) R[k] = site_xmat[9 * s + k];
#endif
  }

  // subtree centres of mass, com-frame inertias, motion axes
  B2_STAGE void com_frame() {
    const int nb = M::nbody();
    B2_UNROLL
    for (int k = 0; k < 3 * nb; k++) com[k] = 0;
    B2_UNROLL
    for (int i = nb - 1; i >= 0; i--) {
      for (int k = 0; k < 3; k++) com[3 * i + k] += xipos[3 * i + k] * M::body_mass(i);
      if (i) { const int p = M::body_parentid(i); for (int k = 0; k < 3; k++) com[3 * p + k] += com[3 * i + k]; }
      if (M::body_subtreemass(i) < Num<T>::minval()) { for (int k = 0; k < 3; k++) com[3 * i + k] = xipos[3 * i + k]; }
      else { const T inv = T(1) / tmax(Num<T>::minval(), M::body_subtreemass(i)); for (int k = 0; k < 3; k++) com[3 * i + k] *= inv; }
    }
    for (int k = 0; k < 10; k++) cinert[k] = 0;
    B2_UNROLL
    for (int i = 1; i < nb; i++) {
      T off[3];
      const int r = M::body_rootid(i);
      B2_LD3(inertia, body_inertia, 3 * i);
      for (int k = 0; k < 3; k++) off[k] = xipos[3 * i + k] - com[3 * r + k];
      inert_about(cinert + 10 * i, inertia, ximat + 9 * i, off, M::body_mass(i));
    }
    B2_UNROLL
    for (int j = 0; j < M::njnt(); j++) {
      const int b = M::jnt_bodyid(j), r = M::body_rootid(b);
      T* cd = cdof + 6 * M::jnt_dofadr(j);
      T off[3];
      for (int k = 0; k < 3; k++) off[k] = com[3 * r + k] - xanchor[3 * j + k];
      const int t = M::jnt_type(j);
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

  B2_STAGE void tendons() {
    const int nv = M::nv();
    B2_UNROLL
    for (int t = 0; t < M::ntendon(); t++) {
      T L = 0;
      B2_UNROLL
      for (int k = 0; k < nv; k++) ten_J[t * nv + k] = 0;
      B2_UNROLL
      for (int w = M::tendon_adr(t); w < M::tendon_adr(t) + M::tendon_num(t); w++) {
        const int j = M::wrap_jntid(w);
        L += M::wrap_coef(w) * qpos[M::jnt_qposadr(j)];
        ten_J[t * nv + M::jnt_dofadr(j)] = M::wrap_coef(w);
      }
      ten_len[t] = L;
    }
  }

  // composite rigid body algorithm -> dense symmetric mass matrix
  B2_STAGE void mass_matrix() {
    const int nb = M::nbody(), nv = M::nv();
    T* crb = spat2;
    B2_UNROLL
    for (int k = 0; k < 10 * nb; k++) crb[k] = cinert[k];
    B2_UNROLL
    for (int i = nb - 1; i > 0; i--) {
      const int p = M::body_parentid(i);
      if (p > 0) for (int k = 0; k < 10; k++) crb[10 * p + k] += crb[10 * i + k];
    }
    B2_UNROLL
    for (int k = 0; k < nv * nv; k++) Mm[k] = 0;
    B2_UNROLL
    for (int i = 0; i < nv; i++) {
      T buf[6];
      Mm[i * nv + i] = M::dof_armature(i);
      inert_mul(buf, crb + 10 * M::dof_bodyid(i), cdof + 6 * i);
      B2_UNROLL
      for (int j = i; j >= 0; j--) {
        if (!is_anc(i, j)) continue;
        T s = 0;
        for (int k = 0; k < 6; k++) s += cdof[6 * j + k] * buf[k];
        Mm[i * nv + j] += s;
        if (j != i) Mm[j * nv + i] = Mm[i * nv + j];
      }
    }
  }

  // in-place L'DL factorisation of the matrix held in A (tree sparsity), di <- 1/D
  B2_STAGE void factor_into(T* A, T* di) {
    const int nv = M::nv();
    B2_UNROLL
    for (int k = nv - 1; k >= 0; k--) {
      const T dkk = A[k * nv + k];
      B2_UNROLL
      for (int i = k - 1; i >= 0; i--) {
        if (!is_anc(k, i)) continue;
        const T t = A[k * nv + i] / dkk;
        B2_UNROLL
        for (int j = i; j >= 0; j--) if (is_anc(i, j)) A[i * nv + j] -= t * A[k * nv + j];
        A[k * nv + i] = t;
      }
      di[k] = T(1) / dkk;
    }
  }
  B2_STAGE void factor_LD() { factor_into(LD, dinv); }
  // both factorisations of the position stage: M, and M + h*diag(damping) for the implicit-in-damping Euler update
  B2_STAGE void factor_mass() {
    const int nv = M::nv();
    B2_UNROLL
    for (int k = 0; k < nv * nv; k++) LD[k] = Mm[k];
    factor_into(LD, dinv);
    if (kKeepFactors && M::has_dofdamping()) {
      B2_UNROLL
      for (int k = 0; k < nv * nv; k++) LDe[k] = Mm[k];
      B2_UNROLL
      for (int k = 0; k < nv; k++) LDe[k * nv + k] += M::timestep() * M::dof_damping(k);
      factor_into(LDe, dinve);
    }
  }
  B2_DEV void solve_with(const T* A, const T* di, T* x) const {
    const int nv = M::nv();
    B2_UNROLL
    for (int i = nv - 1; i >= 0; i--) {
      B2_UNROLL
      for (int j = i - 1; j >= 0; j--) if (is_anc(i, j)) x[j] -= A[i * nv + j] * x[i];
    }
    B2_UNROLL
    for (int i = 0; i < nv; i++) x[i] *= di[i];
    B2_UNROLL
    for (int i = 0; i < nv; i++) {
      B2_UNROLL
      for (int j = i - 1; j >= 0; j--) if (is_anc(i, j)) x[i] -= A[i * nv + j] * x[j];
    }
  }
  B2_DEV void solve(T* x) const { solve_with(LD, dinv, x); }
  B2_DEV void mul_M(T* r, const T* v) const {
    const int nv = M::nv();
    B2_UNROLL
    for (int i = 0; i < nv; i++) r[i] = 0;
    B2_UNROLL
    for (int i = 0; i < nv; i++) {
      r[i] += Mm[i * nv + i] * v[i];
      B2_UNROLL
      for (int j = i - 1; j >= 0; j--) {
        if (!is_anc(i, j)) continue;
        r[i] += Mm[i * nv + j] * v[j];
        r[j] += Mm[i * nv + j] * v[i];
      }
    }
  }

  // point Jacobian columns: calls f(dof, jp[3], jr[3]) for every dof that moves `body`
  // (descending dof order).  `body` may be a runtime value: the dof loop stays static.
  template <class F>
  B2_DEV void for_jac(int body, const T* point, F f) const {
    T off[3];
    const int r = M::body_rootid(body);
    for (int k = 0; k < 3; k++) off[k] = point[k] - com[3 * r + k];
    while (body && !M::body_dofnum(body)) body = M::body_parentid(body);
    if (!body) return;
    const int last = M::body_dofadr(body) + M::body_dofnum(body) - 1;
    B2_UNROLL
    for (int i = M::nv() - 1; i >= 0; i--) {
      if (i > last || !is_anc(last, i)) continue;
      const T* c = cdof + 6 * i;
      T jp[3];
      cross3(jp, c, off);
      jp[0] += c[3]; jp[1] += c[4]; jp[2] += c[5];
      f(i, jp, c);
    }
  }

  B2_STAGE void transmission() {
    const int nv = M::nv();
    B2_UNROLL
    for (int a = 0; a < M::nu(); a++) {
      T* mom = act_moment + a * nv;
      B2_UNROLL
      for (int k = 0; k < nv; k++) mom[k] = 0;
      const int id = M::actuator_trnid(a);
      if (M::actuator_trntype(a) == TRN_JOINT) {
        act_len[a] = qpos[M::jnt_qposadr(id)] * M::actuator_gear(6 * a);
        mom[M::jnt_dofadr(id)] = M::actuator_gear(6 * a);
      } else {
        T wf[3], wt[3];
        B2_LD3(gf, actuator_gear, 6 * a);
        B2_LD3(gt, actuator_gear, 6 * a + 3);
        T sp[3], sR[9];
        site_pose(id, sp, sR);
        mat_vec(wf, sR, gf);
        mat_vec(wt, sR, gt);
        for_jac(M::site_bodyid(id), sp, [&](int d, const T* jp, const T* jr) {
          mom[d] = jp[0] * wf[0] + jp[1] * wf[1] + jp[2] * wf[2] + jr[0] * wt[0] + jr[1] * wt[1] + jr[2] * wt[2];
        });
        act_len[a] = 0;
      }
    }
  }

  // ------------------------------------------------------------------ constraint rows
  B2_DEV int new_row(int type, int id, T pos, T margin) {
    if (nefc >= D::NEFC) { flags |= 8; return -1; }
    const int r = nefc++;
    R.row_type[r] = (signed char)type; R.row_id[r] = (short)id; R.row_pos[r] = pos; R.row_margin[r] = margin;
    B2_UNROLL
    for (int k = 0; k < M::nv(); k++) R.J[r * M::nv() + k] = 0;
    return r;
  }
  // joint and tendon limit rows (they precede contact rows, as in mj_makeConstraint)
  B2_STAGE void limit_rows() {
    const int nv = M::nv();
    nefc = 0;
    B2_UNROLL
    for (int j = 0; j < M::njnt(); j++) {
      if (!M::jnt_limited(j) || M::jnt_type(j) < JNT_SLIDE) continue;
      const T value = qpos[M::jnt_qposadr(j)], margin = M::jnt_margin(j);
      B2_UNROLL
      for (int side = -1; side <= 1; side += 2) {
        const T dist = side * (M::jnt_range(2 * j + (side + 1) / 2) - value);
        if (dist < margin) { const int r = new_row(ROW_LIMIT_JOINT, j, dist, margin); if (r >= 0) R.J[r * nv + M::jnt_dofadr(j)] = T(-side); }
      }
    }
    B2_UNROLL
    for (int t = 0; t < M::ntendon(); t++) {
      if (!M::tendon_limited(t)) continue;
      const T value = ten_len[t], margin = M::tendon_margin(t);
      B2_UNROLL
      for (int side = -1; side <= 1; side += 2) {
        const T dist = side * (M::tendon_range(2 * t + (side + 1) / 2) - value);
        if (dist < margin) {
          const int r = new_row(ROW_LIMIT_TENDON, t, dist, margin);
          if (r >= 0) { B2_UNROLL for (int k = 0; k < nv; k++) R.J[r * nv + k] = -side * ten_J[t * nv + k]; }
        }
      }
    }
  }
  // record one contact of candidate pair `pair` and append its rows: frame * (jac(b2) - jac(b1)),
  // one row for condim 1, four pyramid edges (normal +- mu * tangent) for condim 3
  B2_STAGE bool add_contact(int pair, T dist, const T* pos, const T* normal, const T* tangent_hint) {
    if (ncon >= D::NCON) { flags |= 8; return false; }
    const int c = ncon++;
    R.con_dist[c] = dist; R.con_pair[c] = (short)pair;
    T fr[9];
    for (int k = 0; k < 3; k++) { fr[k] = normal[k]; fr[3 + k] = tangent_hint ? tangent_hint[k] : T(0); }
    make_frame(fr);
    for (int k = 0; k < 3; k++) R.con_pos[3 * c + k] = pos[k];
    for (int k = 0; k < 9; k++) R.con_frame[9 * c + k] = fr[k];
    const T incl = M::pair_margin(pair) - M::pair_gap(pair);
    if (dist >= incl) return true;  // inside the gap: reported but not constrained
    const int nv = M::nv();
    const int b1 = M::geom_bodyid(M::pair_geom1(pair)), b2 = M::geom_bodyid(M::pair_geom2(pair));
    const int dim = M::pair_dim(pair);
    const T mu = M::pair_friction(2 * pair);
    int rows[4];
    const int nrow = dim == 1 ? 1 : 4;
    for (int k = 0; k < nrow; k++) {
      rows[k] = new_row(dim == 1 ? ROW_CONTACT_1 : ROW_CONTACT_PYR, c, dist, incl);
      if (rows[k] < 0) return false;
    }
    // body 1 enters with a minus sign, then body 2 with a plus sign (two explicit passes keep
    // the body index a compile-time constant under a static provider)
    auto accumulate = [&](int body, T sgn) {
      for_jac(body, pos, [&](int d, const T* jp, const T*) {
        const T jn = sgn * dot3(fr, jp);
        if (dim == 1) { R.J[rows[0] * nv + d] += jn; return; }
        const T j1 = sgn * dot3(fr + 3, jp), j2 = sgn * dot3(fr + 6, jp);
        R.J[rows[0] * nv + d] += jn + mu * j1; R.J[rows[1] * nv + d] += jn - mu * j1;
        R.J[rows[2] * nv + d] += jn + mu * j2; R.J[rows[3] * nv + d] += jn - mu * j2;
      });
    };
    accumulate(b1, T(-1));
    accumulate(b2, T(1));
    return true;
  }

  // ------------------------------------------------------------------ collision
  B2_DEV void plane_sphere(int pair, T margin, const T* ppos, const T* normal, const T* spos, T radius, const T* hint) {
    T d[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
    const T cd = dot3(d, normal);
    if (cd > margin + radius) return;
    const T dist = cd - radius;
    const T s = -dist / 2 - radius;
    T pos[3] = {spos[0] + normal[0] * s, spos[1] + normal[1] * s, spos[2] + normal[2] * s};
    add_contact(pair, dist, pos, normal, hint);
  }
  B2_DEV int sphere_sphere(int pair, T margin, const T* p1, const T* z1, T r1, const T* p2, const T* z2, T r2) {
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
  B2_DEV void capsule_capsule(int pair, T margin, const T* p1, const T* z1, T r1, T l1, const T* p2, const T* z2, T r2, T l2) {
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
    B2_NOUNROLL
    for (int side = 1; side >= -1 && n < 2; side -= 2) {
      x = tclip((w - side * mb * l1) / mc, -l2, l2);
      for (int k = 0; k < 3; k++) { v1[k] = p1[k] + z1[k] * (side * l1); v2[k] = p2[k] + z2[k] * x; }
      n += sphere_sphere(pair, margin, v1, z1, r1, v2, z2, r2);
    }
    B2_NOUNROLL
    for (int side = 1; side >= -1 && n < 2; side -= 2) {
      x = tclip((u - side * mb * l2) / ma, -l1, l1);
      for (int k = 0; k < 3; k++) { v2[k] = p2[k] + z2[k] * (side * l2); v1[k] = p1[k] + z1[k] * x; }
      n += sphere_sphere(pair, margin, v1, z1, r1, v2, z2, r2);
    }
  }

  // narrow phase over the statically filtered candidate pairs; contacts append their rows
  B2_STAGE void collide() {
    ncon = 0;
    B2_UNROLL
    for (int p = 0; p < M::npair(); p++) {
      const int g1 = M::pair_geom1(p), g2 = M::pair_geom2(p);
      const T margin = M::pair_margin(p);
      T p1[3], R1[9], p2[3], R2[9];
      geom_pose(g1, p1, R1);
      geom_pose(g2, p2, R2);
      B2_LD3(s1, geom_size, 3 * g1);
      B2_LD3(s2, geom_size, 3 * g2);
      const int t1 = M::geom_type(g1), t2 = M::geom_type(g2);
      T z1[3] = {R1[2], R1[5], R1[8]}, z2[3] = {R2[2], R2[5], R2[8]};
      T d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      if (t1 == GEOM_PLANE) {
        const T h = dot3(d, z1);
        if (h > margin + M::geom_rbound(g2)) continue;
        if (t2 == GEOM_SPHERE) plane_sphere(p, margin, p1, z1, p2, s2[0], nullptr);
        else if (t2 == GEOM_CAPSULE) {
          T e[3];
          for (int k = 0; k < 3; k++) e[k] = p2[k] + z2[k] * s2[1];
          plane_sphere(p, margin, p1, z1, e, s2[0], z2);
          for (int k = 0; k < 3; k++) e[k] = p2[k] - z2[k] * s2[1];
          plane_sphere(p, margin, p1, z1, e, s2[0], z2);
        } else if (t2 == GEOM_BOX) {
          // exact cull: the lowest corner sits reach = sum_k |n . axis_k| s_k below the centre along the plane normal
          // (folds to a constant test when the box never rotates, e.g. a cart on a rail)
          T reach = 0;
          for (int k = 0; k < 3; k++) reach += fabs(z1[0] * R2[k] + z1[1] * R2[3 + k] + z1[2] * R2[6 + k]) * s2[k];
          if (h - reach > margin) continue;
          int cnt = 0;
          B2_NOUNROLL
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
        const T bound = margin + M::geom_rbound(g1) + M::geom_rbound(g2);
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

  static B2_DEV T impedance(const T* si, T pos, T margin) {
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

  // impedance, regulariser D = 1/R and reference acceleration of every row (runtime row loop)
  B2_STAGE void row_params() {
    const int nv = M::nv();
    int within = 0;  // row index inside the current pyramidal contact
    T Rpy = 0;
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) {
      T sr[2], si[5], diag;
      const int id = R.row_id[i], type = R.row_type[i];
      if (type == ROW_LIMIT_JOINT) {
        for (int k = 0; k < 2; k++) sr[k] = M::jnt_solref(2 * id + k);
        for (int k = 0; k < 5; k++) si[k] = M::jnt_solimp(5 * id + k);
        diag = M::dof_invweight0(M::jnt_dofadr(id));
      } else if (type == ROW_LIMIT_TENDON) {
        for (int k = 0; k < 2; k++) sr[k] = M::tendon_solref(2 * id + k);
        for (int k = 0; k < 5; k++) si[k] = M::tendon_solimp(5 * id + k);
        diag = M::tendon_invweight0(id);
      } else {
        const int p = R.con_pair[id];
        for (int k = 0; k < 2; k++) sr[k] = M::pair_solref(2 * p + k);
        for (int k = 0; k < 5; k++) si[k] = M::pair_solimp(5 * p + k);
        const T tran = M::body_invweight0(2 * M::geom_bodyid(M::pair_geom1(p))) + M::body_invweight0(2 * M::geom_bodyid(M::pair_geom2(p)));
        const T mu = M::pair_friction(2 * p);
        diag = (type == ROW_CONTACT_1) ? tran : tran + mu * mu * tran;
      }
      const T pos = R.row_pos[i], margin = R.row_margin[i];
      const T imp = impedance(si, pos, margin);
      const T dmax = tclip(si[1], T(0.0001), T(0.9999));
      T K, B;
      if (sr[0] > 0) {
        const T tc = tmax(sr[0], 2 * M::timestep()), dr = sr[1];
        K = T(1) / tmax(Num<T>::minval(), dmax * dmax * tc * tc * dr * dr);
        B = T(2) / tmax(Num<T>::minval(), dmax * tc);
      } else {
        K = -sr[0] / tmax(Num<T>::minval(), dmax * dmax);
        B = -sr[1] / tmax(Num<T>::minval(), dmax);
      }
      T reg = tmax(Num<T>::minval(), (T(1) - imp) * diag / imp);
      if (type == ROW_CONTACT_PYR) {
        if (within == 0) { const T mu = M::pair_friction(2 * R.con_pair[id]); Rpy = 2 * mu * mu * reg; }
        reg = Rpy;
        within = (within + 1) & 3;
      }
      R.row_D[i] = T(1) / reg;
      T vel = 0;
      B2_UNROLL
      for (int k = 0; k < nv; k++) vel += R.J[i * nv + k] * qvel[k];
      R.row_aref[i] = -B * vel - K * imp * (pos - margin);
    }
  }

  // ------------------------------------------------------------------ velocity stage
  B2_STAGE void velocities() {
    for (int k = 0; k < 6; k++) cvel[k] = 0;
    B2_UNROLL
    for (int i = 1; i < M::nbody(); i++) {
      T v[6];
      const int da = M::body_dofadr(i), dn = M::body_dofnum(i), p = M::body_parentid(i);
      for (int k = 0; k < 6; k++) v[k] = cvel[6 * p + k];
      if (dn == 6 && M::jnt_type(M::dof_jntid(da)) == JNT_FREE) {
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
      } else {
        B2_UNROLL
        for (int j = 0; j < dn; j++) {
          cross_motion(cdof_dot + 6 * (da + j), v, cdof + 6 * (da + j));
          for (int k = 0; k < 6; k++) v[k] += cdof[6 * (da + j) + k] * qvel[da + j];
        }
      }
      for (int k = 0; k < 6; k++) cvel[6 * i + k] = v[k];
    }
  }

  B2_DEV void fluid_body(int i) {
    B2_LD3(I, body_inertia, 3 * i);
    const T mass = M::body_mass(i);
    T box[3], lv[6], lf[6] = {0, 0, 0, 0, 0, 0}, wf[6];
    box[0] = sqrt(tmax(Num<T>::minval(), I[1] + I[2] - I[0]) / mass * T(6));
    box[1] = sqrt(tmax(Num<T>::minval(), I[0] + I[2] - I[1]) / mass * T(6));
    box[2] = sqrt(tmax(Num<T>::minval(), I[0] + I[1] - I[2]) / mass * T(6));
    // com-frame velocity -> local velocity at the inertial frame, minus wind
    const int r = M::body_rootid(i);
    T dif[3], tmp[3], lin[3];
    for (int k = 0; k < 3; k++) dif[k] = xipos[3 * i + k] - com[3 * r + k];
    cross3(tmp, dif, cvel + 6 * i);
    for (int k = 0; k < 3; k++) lin[k] = cvel[6 * i + 3 + k] - tmp[k];
    matT_vec(lv, ximat + 9 * i, cvel + 6 * i);
    matT_vec(lv + 3, ximat + 9 * i, lin);
    T lw[3];
    B2_LD3(wind, wind, 0);
    matT_vec(lw, ximat + 9 * i, wind);
    for (int k = 0; k < 3; k++) lv[3 + k] -= lw[k];
    if (M::viscosity() > 0) {
      const T diam = (box[0] + box[1] + box[2]) / T(3);
      const T ca = -Num<T>::pi() * diam * diam * diam * M::viscosity(), cl = T(-3) * Num<T>::pi() * diam * M::viscosity();
      for (int k = 0; k < 3; k++) { lf[k] = lv[k] * ca; lf[3 + k] = lv[3 + k] * cl; }
    }
    if (M::density() > 0) {
      const T rho = M::density();
      lf[3] -= T(0.5) * rho * box[1] * box[2] * fabs(lv[3]) * lv[3];
      lf[4] -= T(0.5) * rho * box[0] * box[2] * fabs(lv[4]) * lv[4];
      lf[5] -= T(0.5) * rho * box[0] * box[1] * fabs(lv[5]) * lv[5];
      const T b0 = box[0] * box[0], b1 = box[1] * box[1], b2 = box[2] * box[2];
      lf[0] -= rho * box[0] * (b1 * b1 + b2 * b2) * fabs(lv[0]) * lv[0] / T(64);
      lf[1] -= rho * box[1] * (b0 * b0 + b2 * b2) * fabs(lv[1]) * lv[1] / T(64);
      lf[2] -= rho * box[2] * (b0 * b0 + b1 * b1) * fabs(lv[2]) * lv[2] / T(64);
    }
    mat_vec(wf, ximat + 9 * i, lf);
    mat_vec(wf + 3, ximat + 9 * i, lf + 3);
    for_jac(i, xipos + 3 * i, [&](int d, const T* jp, const T* jr) {
      f_passive[d] += jp[0] * wf[3] + jp[1] * wf[4] + jp[2] * wf[5] + jr[0] * wf[0] + jr[1] * wf[1] + jr[2] * wf[2];
    });
  }

  B2_STAGE void passive_forces() {
    const int nv = M::nv();
    B2_UNROLL
    for (int k = 0; k < nv; k++) f_passive[k] = 0;
    B2_UNROLL
    for (int j = 0; j < M::njnt(); j++) {
      const T st = M::jnt_stiffness(j);
      if (st == 0) continue;
      const int pa = M::jnt_qposadr(j), da = M::jnt_dofadr(j);
      if (M::jnt_type(j) >= JNT_SLIDE) f_passive[da] -= st * (qpos[pa] - M::qpos_spring(pa));
      else if (M::jnt_type(j) == JNT_FREE) {
        T q[4], qs[4] = {M::qpos_spring(pa + 3), -M::qpos_spring(pa + 4), -M::qpos_spring(pa + 5), -M::qpos_spring(pa + 6)}, qd[4], dv[3];
        for (int k = 0; k < 3; k++) f_passive[da + k] -= st * (qpos[pa + k] - M::qpos_spring(pa + k));
        for (int k = 0; k < 4; k++) q[k] = qpos[pa + 3 + k];
        normalize4(q);
        quat_mul(qd, qs, q);
        quat_to_vel(dv, qd, T(1));
        for (int k = 0; k < 3; k++) f_passive[da + 3 + k] -= st * dv[k];
      }
    }
    B2_UNROLL
    for (int k = 0; k < nv; k++) f_passive[k] -= M::dof_damping(k) * qvel[k];
    B2_UNROLL
    for (int t = 0; t < M::ntendon(); t++) {
      const T st = M::tendon_stiffness(t), dm = M::tendon_damping(t);
      if (st == 0 && dm == 0) continue;
      T frc = 0, vel = 0;
      const T lo = M::tendon_lengthspring(2 * t), hi = M::tendon_lengthspring(2 * t + 1), L = ten_len[t];
      if (L > hi) frc = st * (hi - L); else if (L < lo) frc = st * (lo - L);
      B2_UNROLL
      for (int k = 0; k < nv; k++) vel += ten_J[t * nv + k] * qvel[k];
      frc -= dm * vel;
      B2_UNROLL
      for (int k = 0; k < nv; k++) f_passive[k] += ten_J[t * nv + k] * frc;
    }
    if (M::has_fluid()) {
      B2_UNROLL
      for (int i = 1; i < M::nbody(); i++)
        if (M::body_mass(i) >= Num<T>::minval()) fluid_body(i);
    }
  }

  // recursive Newton-Euler without accelerations: Coriolis, centrifugal, gravity
  B2_STAGE void bias_forces() {
    const int nb = M::nbody(), nv = M::nv();
    T* cacc = spat;
    T* cfrc = spat2;
    cacc[0] = cacc[1] = cacc[2] = 0;
    cacc[3] = -M::gravity(0); cacc[4] = -M::gravity(1); cacc[5] = -M::gravity(2);
    for (int k = 0; k < 6; k++) cfrc[k] = 0;
    B2_UNROLL
    for (int i = 1; i < nb; i++) {
      const int da = M::body_dofadr(i), p = M::body_parentid(i);
      T t[6] = {0, 0, 0, 0, 0, 0}, t1[6];
      B2_UNROLL
      for (int j = 0; j < M::body_dofnum(i); j++)
        for (int k = 0; k < 6; k++) t[k] += cdof_dot[6 * (da + j) + k] * qvel[da + j];
      for (int k = 0; k < 6; k++) cacc[6 * i + k] = cacc[6 * p + k] + t[k];
      inert_mul(cfrc + 6 * i, cinert + 10 * i, cacc + 6 * i);
      inert_mul(t, cinert + 10 * i, cvel + 6 * i);
      cross_force(t1, cvel + 6 * i, t);
      for (int k = 0; k < 6; k++) cfrc[6 * i + k] += t1[k];
    }
    B2_UNROLL
    for (int i = nb - 1; i > 0; i--) {
      const int p = M::body_parentid(i);
      if (p) for (int k = 0; k < 6; k++) cfrc[6 * p + k] += cfrc[6 * i + k];
    }
    B2_UNROLL
    for (int d = 0; d < nv; d++) {
      T s = 0;
      const int b = M::dof_bodyid(d);
      for (int k = 0; k < 6; k++) s += cdof[6 * d + k] * cfrc[6 * b + k];
      f_bias[d] = s;
    }
  }

  // ------------------------------------------------------------------ sensors
  // mj_sensorPos / mj_sensorVel / mj_sensorAcc for the compiled subset (type codes SENS_*); call after forward().
  // Feeds data.sensordata (reference mujoco_template/observations.py:117-127, examples/drone/x2.xml:83-87).
  B2_DEV void site_motion(int s, const T* vec6, T* res) const {  // com-based spatial vector -> site-local frame
    T p[3], R[9];
    site_pose(s, p, R);
    const int b = M::site_bodyid(s), root = M::body_rootid(b);
    const T dif[3] = {p[0] - com[3 * root], p[1] - com[3 * root + 1], p[2] - com[3 * root + 2]};
    T cr[3], lin[3];
    cross3(cr, dif, vec6);
    for (int k = 0; k < 3; k++) lin[k] = vec6[3 + k] - cr[k];
    matT_vec(res, R, vec6);
    matT_vec(res + 3, R, lin);
  }
  B2_DEV void sensors(T* out) {
    bool need_acc = false;
    B2_UNROLL
    for (int s = 0; s < M::nsensor(); s++) need_acc |= M::sensor_type(s) == SENS_ACCELEROMETER;
    T* cacc = spat;
    if (need_acc) {  // acceleration pass of mj_rnePostConstraint: cacc = cacc_parent + cdof_dot*qvel + cdof*qacc
      cacc[0] = cacc[1] = cacc[2] = 0;
      cacc[3] = -M::gravity(0); cacc[4] = -M::gravity(1); cacc[5] = -M::gravity(2);
      B2_UNROLL
      for (int i = 1; i < M::nbody(); i++) {
        const int da = M::body_dofadr(i), p = M::body_parentid(i);
        T a[6];
        for (int k = 0; k < 6; k++) a[k] = cacc[6 * p + k];
        B2_UNROLL
        for (int j = 0; j < M::body_dofnum(i); j++)
          for (int k = 0; k < 6; k++) a[k] += cdof_dot[6 * (da + j) + k] * qvel[da + j] + cdof[6 * (da + j) + k] * qacc[da + j];
        for (int k = 0; k < 6; k++) cacc[6 * i + k] = a[k];
      }
    }
    B2_UNROLL
    for (int s = 0; s < M::nsensor(); s++) {
      const int type = M::sensor_type(s), id = M::sensor_objid(s), ot = M::sensor_objtype(s), adr = M::sensor_adr(s);
      T v[6] = {0, 0, 0, 0, 0, 0};
      int dim = 3;
      if (type == SENS_JOINTPOS) { v[0] = qpos[M::jnt_qposadr(id)]; dim = 1; }
      else if (type == SENS_JOINTVEL) { v[0] = qvel[M::jnt_dofadr(id)]; dim = 1; }
      else if (type == SENS_FRAMEPOS) {
        if (ot == 1) { for (int k = 0; k < 3; k++) v[k] = xipos[3 * id + k]; }
        else if (ot == 2) { for (int k = 0; k < 3; k++) v[k] = xpos[3 * id + k]; }
        else { T R[9]; if (ot == 5) geom_pose(id, v, R); else site_pose(id, v, R); }
      } else if (type == SENS_FRAMEQUAT) {
        dim = 4;
        if (ot == 2) { for (int k = 0; k < 4; k++) v[k] = xquat[4 * id + k]; }
        else {
          T lq[4];
          int b = id;
          if (ot == 1) { for (int k = 0; k < 4; k++) lq[k] = M::body_iquat(4 * id + k); }
          else if (ot == 5) { b = M::geom_bodyid(id); for (int k = 0; k < 4; k++) lq[k] = M::geom_quat(4 * id + k); }
          else { b = M::site_bodyid(id); for (int k = 0; k < 4; k++) lq[k] = M::site_quat(4 * id + k); }
          quat_mul(v, xquat + 4 * b, lq);
        }
        normalize4(v);
      } else if (type == SENS_GYRO || type == SENS_VELOCIMETER) {
        T loc[6];
        site_motion(id, cvel + 6 * M::site_bodyid(id), loc);
        for (int k = 0; k < 3; k++) v[k] = loc[(type == SENS_GYRO ? 0 : 3) + k];
      } else if (type == SENS_ACCELEROMETER) {
        T vel[6], acc[6], corr[3];
        site_motion(id, cvel + 6 * M::site_bodyid(id), vel);
        site_motion(id, cacc + 6 * M::site_bodyid(id), acc);
        cross3(corr, vel, vel + 3);
        for (int k = 0; k < 3; k++) v[k] = acc[3 + k] + corr[k];
      }
      const T cut = M::sensor_cutoff(s);
      B2_UNROLL
      for (int k = 0; k < 4; k++)
        if (k < dim) out[adr + k] = (cut > 0 && type != SENS_FRAMEQUAT) ? tclip(v[k], -cut, cut) : v[k];
    }
  }

  // ------------------------------------------------------------------ smooth acceleration
  B2_STAGE void smooth_dynamics() {
    const int nv = M::nv();
    T* f_act = grad;  // scratch: generalized actuator force
    B2_UNROLL
    for (int k = 0; k < nv; k++) f_act[k] = 0;
    B2_UNROLL
    for (int a = 0; a < M::nu(); a++) {
      T u = ctrl[a];
      if (M::actuator_ctrllimited(a)) u = tclip(u, M::actuator_ctrlrange(2 * a), M::actuator_ctrlrange(2 * a + 1));
      T vel = 0;
      B2_UNROLL
      for (int k = 0; k < nv; k++) vel += act_moment[a * nv + k] * qvel[k];
      T force = M::actuator_gainprm(a) * u + M::actuator_biasprm(3 * a) + M::actuator_biasprm(3 * a + 1) * act_len[a] +
                M::actuator_biasprm(3 * a + 2) * vel;
      if (M::actuator_forcelimited(a)) force = tclip(force, M::actuator_forcerange(2 * a), M::actuator_forcerange(2 * a + 1));
      if (M::actuator_disabled(a)) force = 0;
      B2_UNROLL
      for (int k = 0; k < nv; k++) f_act[k] += act_moment[a * nv + k] * force;
    }
    B2_UNROLL
    for (int k = 0; k < nv; k++) {
      f_smooth[k] = f_passive[k] - f_bias[k];
      f_smooth[k] += f_act[k];
      a_smooth[k] = f_smooth[k];
    }
    solve(a_smooth);
  }

  // ------------------------------------------------------------------ Newton solver
  // constraint cost at jar; optionally refresh f_con = R.J' * force
  B2_DEV T row_cost(const T* jar, bool write_force) {
    const int nv = M::nv();
    T c = 0;
    if (write_force) { B2_UNROLL for (int k = 0; k < nv; k++) f_con[k] = 0; }
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) {
      if (jar[i] >= 0) continue;
      c += T(0.5) * R.row_D[i] * jar[i] * jar[i];
      if (write_force) { const T f = -R.row_D[i] * jar[i]; B2_UNROLL for (int k = 0; k < nv; k++) f_con[k] += R.J[i * nv + k] * f; }
    }
    return c;
  }
  struct LsPoint { T alpha, cost, d1, d2; };
  B2_DEV void ls_eval(T alpha, LsPoint& p) {
    ls_iter++;
    T q0 = qg0, q1 = qg1, q2 = qg2;
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) {
      if (R.Jaref[i] + alpha * R.Jv[i] < 0) {
        const T dj = R.row_D[i] * R.Jaref[i];
        q0 += T(0.5) * R.Jaref[i] * dj; q1 += R.Jv[i] * dj; q2 += T(0.5) * R.Jv[i] * R.row_D[i] * R.Jv[i];
      }
    }
    p.alpha = alpha; p.cost = alpha * alpha * q2 + alpha * q1 + q0; p.d1 = 2 * alpha * q2 + q1; p.d2 = 2 * q2;
    if (p.d2 <= 0) p.d2 = Num<T>::minval();
  }
  B2_DEV int ls_bracket(LsPoint& p, const LsPoint* cand, LsPoint& pnext) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = 1; }
      else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = 2; }
    }
    if (flag) ls_eval(p.alpha - p.d1 / p.d2, pnext);
    return flag;
  }
  // exact 1-D minimisation of the piecewise-quadratic cost along `search`
  B2_STAGE T line_search() {
    const int nv = M::nv();
    LsPoint p0, p1, p2, pmid, p1n, p2n;
    ls_iter = 0;
    T sn = 0;
    B2_UNROLL
    for (int k = 0; k < nv; k++) sn += search[k] * search[k];
    sn = sqrt(sn);
    if (sn < Num<T>::minval()) return 0;
    const T scale = T(1) / (M::meaninertia() * T(nv > 1 ? nv : 1));
    const T gtol = M::tolerance() * M::ls_tolerance() * sn / scale;
    mul_M(Mv, search);
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) { T s = 0; B2_UNROLL for (int k = 0; k < nv; k++) s += R.J[i * nv + k] * search[k]; R.Jv[i] = s; }
    T a = 0, b = 0, e = 0;
    B2_UNROLL
    for (int k = 0; k < nv; k++) { a += search[k] * Ma[k]; b += f_smooth[k] * search[k]; e += search[k] * Mv[k]; }
    qg0 = gauss; qg1 = a - b; qg2 = T(0.5) * e;
    ls_eval(0, p0);
    ls_eval(p0.alpha - p0.d1 / p0.d2, p1);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.d1) < gtol) return p1.alpha;
    const int dir = p1.d1 < 0 ? 1 : -1;
    bool p2up = false;
    const int maxls = M::ls_iterations();
    B2_NOUNROLL
    while (p1.d1 * dir <= -gtol && ls_iter < maxls) {
      p2 = p1; p2up = true;
      ls_eval(p1.alpha - p1.d1 / p1.d2, p1);
      if (fabs(p1.d1) < gtol) return p1.alpha;
    }
    if (ls_iter >= maxls || !p2up) return p1.alpha;
    p2n = p1;
    ls_eval(p1.alpha - p1.d1 / p1.d2, p1n);
    B2_NOUNROLL
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
  // cost, forces and gradient at the current qacc ...
  B2_STAGE void newton_gradient() {
    const int nv = M::nv();
    cost = row_cost(R.Jaref, true);
    T g = 0;
    B2_UNROLL
    for (int k = 0; k < nv; k++) g += (Ma[k] - f_smooth[k]) * (qacc[k] - a_smooth[k]);
    gauss = T(0.5) * g;
    cost += gauss;
    B2_UNROLL
    for (int k = 0; k < nv; k++) grad[k] = Ma[k] - f_smooth[k] - f_con[k];
  }
  // ... and the Newton direction (not needed by the round that detects convergence: one factorisation fewer per solve)
  B2_STAGE void newton_direction() {
    const int nv = M::nv();
    // H = M + R.J' D_active R.J  (lower triangle, held in LD), dense Cholesky, Mgrad = H^-1 grad
    T* H = kKeepFactors ? Hs : LD;
    B2_UNROLL
    for (int k = 0; k < nv * nv; k++) H[k] = Mm[k];
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) {
      if (R.Jaref[i] >= 0) continue;
      const T* Ji = R.J + i * nv;
      B2_UNROLL
      for (int r = 0; r < nv; r++) {
        if (Ji[r] == 0) continue;
        const T s = R.row_D[i] * Ji[r];
        B2_UNROLL
        for (int q = 0; q <= r; q++) H[r * nv + q] += s * Ji[q];
      }
    }
    B2_UNROLL
    for (int j = 0; j < nv; j++) {
      T t = H[j * nv + j];
      B2_UNROLL
      for (int k = 0; k < j; k++) t -= H[j * nv + k] * H[j * nv + k];
      if (t < Num<T>::minval()) t = Num<T>::minval();
      const T djj = sqrt(t);
      H[j * nv + j] = djj;
      const T inv = T(1) / djj;
      B2_UNROLL
      for (int i = 0; i < nv; i++) {  // constant bounds + predicate: a triangular bound is not unrolled, and one runtime
        if (i <= j) continue;         // index into a LaneEnv member keeps the whole struct in local memory
        T s = H[i * nv + j];
        B2_UNROLL
        for (int k = 0; k < nv; k++) if (k < j) s -= H[i * nv + k] * H[j * nv + k];
        H[i * nv + j] = s * inv;
      }
    }
    B2_UNROLL
    for (int i = 0; i < nv; i++) {
      T s = grad[i];
      B2_UNROLL
      for (int k = 0; k < i; k++) s -= H[i * nv + k] * Mgrad[k];
      Mgrad[i] = s / H[i * nv + i];
    }
    B2_UNROLL
    for (int i = nv - 1; i >= 0; i--) {
      T s = Mgrad[i];
      B2_UNROLL
      for (int k = i + 1; k < nv; k++) s -= H[k * nv + i] * Mgrad[k];
      Mgrad[i] = s / H[i * nv + i];
    }
  }
  // warm-started Newton solve (only entered when at least one row exists)
  B2_STAGE void newton_solve() {
    const int nv = M::nv();
    // warm start: keep qacc_warmstart only if it is cheaper than the unconstrained acceleration
    B2_UNROLL
    for (int k = 0; k < nv; k++) qacc[k] = warm[k];
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) { T s = 0; B2_UNROLL for (int k = 0; k < nv; k++) s += R.J[i * nv + k] * qacc[k]; R.Jaref[i] = s - R.row_aref[i]; }
    T cw = row_cost(R.Jaref, false);
    mul_M(Ma, qacc);
    B2_UNROLL
    for (int k = 0; k < nv; k++) cw += T(0.5) * (Ma[k] - f_smooth[k]) * (qacc[k] - a_smooth[k]);
    B2_NOUNROLL
    for (int i = 0; i < nefc; i++) { T s = 0; B2_UNROLL for (int k = 0; k < nv; k++) s += R.J[i * nv + k] * a_smooth[k]; R.Jv[i] = s - R.row_aref[i]; }
    const T cs = row_cost(R.Jv, false);
    if (cw > cs) {
      B2_UNROLL
      for (int k = 0; k < nv; k++) qacc[k] = a_smooth[k];
      B2_NOUNROLL
      for (int i = 0; i < nefc; i++) R.Jaref[i] = R.Jv[i];
      mul_M(Ma, qacc);
    }
    const T scale = T(1) / (M::meaninertia() * T(nv > 1 ? nv : 1));
    T old = 0;
    bool first = true;
    // one refresh / one line-search call site: the loop is the upstream iteration unrolled by half a turn
    B2_NOUNROLL
    while (true) {
      newton_gradient();
      if (!first) {
        T gn = 0;
        B2_UNROLL
        for (int k = 0; k < nv; k++) gn += grad[k] * grad[k];
        niter++;
        if (scale * (old - cost) < M::tolerance() || scale * sqrt(gn) < M::tolerance()) break;
      }
      first = false;
      if (niter >= M::iterations()) break;
      newton_direction();
      B2_UNROLL
      for (int k = 0; k < nv; k++) search[k] = -Mgrad[k];
      const T alpha = line_search();
      if (alpha == 0) break;
      B2_UNROLL
      for (int k = 0; k < nv; k++) { qacc[k] += alpha * search[k]; Ma[k] += alpha * Mv[k]; }
      B2_NOUNROLL
      for (int i = 0; i < nefc; i++) R.Jaref[i] += alpha * R.Jv[i];
      old = cost;
    }
    B2_UNROLL
    for (int k = 0; k < nv; k++) warm[k] = qacc[k];
  }
  B2_DEV void constrained_acceleration() {
    const int nv = M::nv();
    niter = 0;
    if (!nefc) {
      B2_UNROLL
      for (int k = 0; k < nv; k++) { qacc[k] = a_smooth[k]; warm[k] = a_smooth[k]; f_con[k] = 0; }
      return;
    }
    row_params();
    newton_solve();
  }

  // ------------------------------------------------------------------ forward + integrators
  // position-dependent part of mj_forward: poses, inertias, mass matrix, limit and contact rows,
  // actuator moments.  Depends on qpos only, so the FD kernel reuses it across rollouts that
  // perturb velocities or controls (upstream's mjSTAGE_POS skip, bit-identical by construction).
  template <bool LATE = false>
  B2_STAGE void forward_position() {
    kinematics();
    com_frame();
    tendons();
    if (!LATE) mass_matrix();
    if (kKeepFactors) factor_mass();
    limit_rows();
    collide();
    if (!LATE) transmission();
  }
  // velocity-dependent part (depends on qpos and qvel only): reused by FD rollouts that perturb a control
  // (upstream's mjSTAGE_VEL skip)
  B2_STAGE void forward_velocity() {
    velocities();
    passive_forces();
    bias_forces();
  }
  template <bool LATE = false>
  B2_STAGE void forward_acc() {
    if (LATE) { mass_matrix(); transmission(); }
    if (!kKeepFactors) factor_mass();
    smooth_dynamics();
    constrained_acceleration();
  }
  B2_STAGE void forward_rest() {
    forward_velocity();
    forward_acc();
  }
  B2_DEV void forward() {
    forward_position();
    forward_rest();
  }
  // inverse dynamics (mj_inverse, continuous time) at (qpos, qvel) for a prescribed acceleration:
  // out = M acc + bias - passive - J' f, with the row forces the prescribed acceleration implies
  B2_DEV void inverse(const T* acc, T* out) {
    const int nv = M::nv();
    forward_position();
    velocities();
    passive_forces();
    bias_forces();
    B2_UNROLL
    for (int k = 0; k < nv; k++) f_con[k] = 0;
    if (nefc) {
      row_params();
      B2_NOUNROLL
      for (int i = 0; i < nefc; i++) {
        T jar = -R.row_aref[i];
        B2_UNROLL
        for (int k = 0; k < nv; k++) jar += R.J[i * nv + k] * acc[k];
        if (jar >= 0) continue;
        const T f = -R.row_D[i] * jar;
        B2_UNROLL
        for (int k = 0; k < nv; k++) f_con[k] += R.J[i * nv + k] * f;
      }
    }
    mul_M(Ma, acc);
    B2_UNROLL
    for (int k = 0; k < nv; k++) out[k] = Ma[k] + f_bias[k] - f_passive[k] - f_con[k];
  }
  B2_DEV void integrate_pos(T* q, const T* v, T dt) const {
    B2_UNROLL
    for (int j = 0; j < M::njnt(); j++) {
      const int pa = M::jnt_qposadr(j), va = M::jnt_dofadr(j);
      if (M::jnt_type(j) == JNT_FREE) {
        for (int k = 0; k < 3; k++) q[pa + k] += dt * v[va + k];
        quat_integrate(q + pa + 3, v + va + 3, dt);
      } else q[pa] += dt * v[va];
    }
  }
  B2_DEV void differentiate_pos(T* out, T dt, const T* q1, const T* q2) const {
    B2_UNROLL
    for (int j = 0; j < M::njnt(); j++) {
      const int pa = M::jnt_qposadr(j), va = M::jnt_dofadr(j);
      if (M::jnt_type(j) == JNT_FREE) {
        for (int k = 0; k < 3; k++) out[va + k] = (q2[pa + k] - q1[pa + k]) / dt;
        T neg[4] = {q1[pa + 3], -q1[pa + 4], -q1[pa + 5], -q1[pa + 6]}, dq[4];
        quat_mul(dq, neg, q2 + pa + 3);
        quat_to_vel(out + va + 3, dq, dt);
      } else out[va] = (q2[pa] - q1[pa]) / dt;
    }
  }
  B2_DEV void check_state() {
    B2_UNROLL
    for (int k = 0; k < M::nq(); k++) if (!(fabs(qpos[k]) <= T(1e10))) flags |= 1;
    B2_UNROLL
    for (int k = 0; k < M::nv(); k++) if (!(fabs(qvel[k]) <= T(1e10))) flags |= 2;
  }
  B2_STAGE void euler() {
    const int nv = M::nv();
    const T h = M::timestep();
    T* acc = grad;
    if (!M::has_dofdamping()) { B2_UNROLL for (int k = 0; k < nv; k++) acc[k] = qacc[k]; }
    else {
      // (M + h*diag(damping)) acc = f_smooth + f_con   (implicit in joint damping)
      if (!kKeepFactors) {
        B2_UNROLL
        for (int k = 0; k < nv * nv; k++) LD[k] = Mm[k];
        B2_UNROLL
        for (int k = 0; k < nv; k++) LD[k * nv + k] += h * M::dof_damping(k);
        factor_LD();
      }
      B2_UNROLL
      for (int k = 0; k < nv; k++) acc[k] = f_smooth[k] + f_con[k];
      solve_with(kKeepFactors ? LDe : LD, kKeepFactors ? dinve : dinv, acc);
    }
    B2_UNROLL
    for (int k = 0; k < nv; k++) qvel[k] += acc[k] * h;
    integrate_pos(qpos, qvel, h);
  }
  // classical RK4 over (qpos, qvel) with tangent-space position updates; the stage loop is
  // kept rolled so forward() is instantiated once here
  B2_STAGE void rk4() {
    const int nq = M::nq(), nv = M::nv();
    const T h = M::timestep();
    const T A[9] = {T(0.5), 0, 0, 0, T(0.5), 0, 0, 0, 1}, B[4] = {T(1) / 6, T(1) / 3, T(1) / 3, T(1) / 6};
    T X[4][D::NQ + D::NV], F[4][D::NV], dX[2 * D::NV];
    B2_UNROLL
    for (int k = 0; k < nq; k++) X[0][k] = qpos[k];
    B2_UNROLL
    for (int k = 0; k < nv; k++) { X[0][nq + k] = qvel[k]; F[0][k] = qacc[k]; }
    B2_NOUNROLL
    for (int i = 1; i < 4; i++) {
      B2_UNROLL
      for (int k = 0; k < 2 * nv; k++) dX[k] = 0;
      for (int j = 0; j < i; j++) {
        const T a = A[(i - 1) * 3 + j];
        B2_UNROLL
        for (int k = 0; k < nv; k++) { dX[k] += a * X[j][nq + k]; dX[nv + k] += a * F[j][k]; }
      }
      B2_UNROLL
      for (int k = 0; k < nq; k++) qpos[k] = X[0][k];
      integrate_pos(qpos, dX, h);
      B2_UNROLL
      for (int k = 0; k < nv; k++) qvel[k] = X[0][nq + k] + h * dX[nv + k];
      B2_UNROLL
      for (int k = 0; k < nq; k++) X[i][k] = qpos[k];
      B2_UNROLL
      for (int k = 0; k < nv; k++) X[i][nq + k] = qvel[k];
      forward();
      B2_UNROLL
      for (int k = 0; k < nv; k++) F[i][k] = qacc[k];
    }
    B2_UNROLL
    for (int k = 0; k < 2 * nv; k++) dX[k] = 0;
    for (int j = 0; j < 4; j++) {
      B2_UNROLL
      for (int k = 0; k < nv; k++) { dX[k] += B[j] * X[j][nq + k]; dX[nv + k] += B[j] * F[j][k]; }
    }
    B2_UNROLL
    for (int k = 0; k < nq; k++) qpos[k] = X[0][k];
    B2_UNROLL
    for (int k = 0; k < nv; k++) qvel[k] = X[0][nq + k] + dX[nv + k] * h;
    integrate_pos(qpos, dX, h);
  }
  // one mj_step; `snapshot` is invoked after the pre-integration forward pass
  template <class Snap>
  B2_DEV void step(Snap snapshot) {
    check_state();
    forward();
    B2_UNROLL
    for (int k = 0; k < M::nv(); k++) if (!(fabs(qacc[k]) <= T(1e10))) flags |= 4;
    snapshot();
    if (M::integrator() == 1) rk4(); else euler();
  }
  B2_DEV void step() { step([] {}); }
};

}  // namespace b2
