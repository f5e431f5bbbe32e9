/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * CPU oracle: scalar FP64 restatement of what the reference reaches through
 *   mj.mj_step            (reference mujoco_template/model.py:56-57)
 *   mj.mj_forward         (reference mujoco_template/model.py:53-54)
 *   mj.mj_resetData / mj_resetDataKeyframe (model.py:59-71)
 *   mj.mjd_transitionFD   (reference mujoco_template/linearization.py:16-35)
 *   mj.mj_integratePos / mj_differentiatePos (linearization.py:10-13,55-70)
 *   mj.mj_jacSite/Body/BodyCom/SubtreeCom (reference mujoco_template/jacobians.py:44-79)
 *
 * The arithmetic itself lives in the third-party dependency `mujoco` (pinned only as
 * `mujoco>=3.1`, reference pyproject.toml:10-14; not vendored, not installable offline).
 * Each function below names the upstream MuJoCo routine whose published algorithm it
 * restates (target: 3.1.x semantics).
 *
 * PARITY UNPINNED against a live MuJoCo for mj_step, the Newton solver and mjd_transitionFD (no golden
 * vectors exist in the reference and mujoco cannot be imported here).  One path IS pinned against the real
 * engine: mj_inverse of the humanoid at the LQR tutorial's keyframe reproduces the output the upstream
 * notebook publishes, digit for digit (tests/test_published_values.py: compile, kinematics, bias forces,
 * foot contacts and constraint rows).  Everything else is pinned by analytic known answers, independent
 * derivations and invariants: tests/test_oracle_analytic.py, tests/test_oracle_independent.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared).  Loaded via ctypes by oracle/oracle.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/b2_model_layout.h"
#include "orc_math.h"

/* The op-counting build (make liborc_count.so, orc_count.h) compiles this file as C++ with `double` replaced by a
 * counting class; the exported names stay C names. */
#ifdef __cplusplus
extern "C" {
#endif

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4, GEOM_BOX = 6 };
enum { TRN_JOINT = 0, TRN_SITE = 4 };
enum { EFC_LIMIT_JOINT = 0, EFC_LIMIT_TENDON = 1, EFC_CONTACT_FRICTIONLESS = 2, EFC_CONTACT_PYRAMIDAL = 3 };

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5];
  int dim, geom1, geom2, exclude, pair;
} orc_contact;

typedef struct orc_model {
  void* blob;
  size_t nbytes;
  b2m_view v;
  int maxcon, maxefc;
} orc_model;

typedef struct orc_data {
  /* state */
  double time, *qpos, *qvel, *ctrl, *qacc, *qacc_warmstart;
  /* position stage */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos, *site_xmat;
  double *subtree_com, *cinert, *crb, *cdof, *qM, *qLD, *qLDiagInv, *qH, *qHDiagInv;
  double *ten_length, *ten_J, *actuator_length, *actuator_moment;
  /* velocity stage */
  double *cvel, *cdof_dot, *ten_velocity, *actuator_velocity, *qfrc_bias, *qfrc_passive;
  /* acceleration stage */
  double *actuator_force, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qfrc_constraint;
  /* sensors */
  double *cacc, *sensordata;
  /* constraints */
  int ncon, nefc;
  orc_contact* contact;
  int *efc_type, *efc_id;
  double *efc_J, *efc_pos, *efc_margin, *efc_D, *efc_R, *efc_aref, *efc_vel, *efc_force, *efc_diagApprox;
  /* solver stats / flags */
  int solver_iter, warn_bad_qpos, warn_bad_qvel, warn_bad_qacc, warn_overflow;
  /* bump-allocated scratch for per-call temporaries (no malloc on the hot path) */
  double* arena; size_t arena_top, arena_cap;
} orc_data;

#define ALLOC(n) ((double*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(double)))
/* scratch from the per-data arena: zero-filled, released by restoring the saved top */
static double* orc_scratch(struct orc_data* d, size_t n);
#define SCRATCH(n) orc_scratch(d, (size_t)((n) > 0 ? (n) : 1))
#define ARENA_MARK size_t arena_mark__ = d->arena_top
#define ARENA_RELEASE d->arena_top = arena_mark__

static double* orc_scratch(struct orc_data* d, size_t n) {
  if (d->arena_top + n > d->arena_cap) { fprintf(stderr, "oracle scratch arena exhausted\n"); abort(); }
  double* p = d->arena + d->arena_top;
  d->arena_top += n;
  memset(p, 0, n * sizeof(double));
  return p;
}

/* ------------------------------------------------------------------ lifecycle */
orc_model* orc_model_create(const void* blob, size_t nbytes) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  m->blob = malloc(nbytes);
  memcpy(m->blob, blob, nbytes);
  m->nbytes = nbytes;
  if (b2m_view_init(&m->v, m->blob, nbytes)) { free(m->blob); free(m); return NULL; }
  m->maxcon = 4 * m->v.npair + 1;
  m->maxefc = 2 * m->v.njnt + 2 * m->v.ntendon + 4 * m->maxcon + 1;
  return m;
}
void orc_model_free(orc_model* m) { if (m) { free(m->blob); free(m); } }
const b2m_view* orc_model_view(const orc_model* m) { return &m->v; }
int orc_model_nsensordata(const orc_model* m) { return m->v.nsensordata; }

orc_data* orc_data_create(const orc_model* m) {
  const b2m_view* v = &m->v;
  orc_data* d = (orc_data*)calloc(1, sizeof(orc_data));
  int nq = v->nq, nv = v->nv, nu = v->nu, nb = v->nbody, nj = v->njnt, ng = v->ngeom, ns = v->nsite, nt = v->ntendon;
  d->qpos = ALLOC(nq); d->qvel = ALLOC(nv); d->ctrl = ALLOC(nu); d->qacc = ALLOC(nv); d->qacc_warmstart = ALLOC(nv);
  d->xpos = ALLOC(3 * nb); d->xquat = ALLOC(4 * nb); d->xmat = ALLOC(9 * nb); d->xipos = ALLOC(3 * nb); d->ximat = ALLOC(9 * nb);
  d->xanchor = ALLOC(3 * nj); d->xaxis = ALLOC(3 * nj);
  d->geom_xpos = ALLOC(3 * ng); d->geom_xmat = ALLOC(9 * ng); d->site_xpos = ALLOC(3 * ns); d->site_xmat = ALLOC(9 * ns);
  d->subtree_com = ALLOC(3 * nb); d->cinert = ALLOC(10 * nb); d->crb = ALLOC(10 * nb); d->cdof = ALLOC(6 * nv);
  d->qM = ALLOC(nv * nv); d->qLD = ALLOC(nv * nv); d->qLDiagInv = ALLOC(nv); d->qH = ALLOC(nv * nv); d->qHDiagInv = ALLOC(nv);
  d->ten_length = ALLOC(nt); d->ten_J = ALLOC(nt * nv); d->actuator_length = ALLOC(nu); d->actuator_moment = ALLOC(nu * nv);
  d->cvel = ALLOC(6 * nb); d->cdof_dot = ALLOC(6 * nv); d->ten_velocity = ALLOC(nt); d->actuator_velocity = ALLOC(nu);
  d->qfrc_bias = ALLOC(nv); d->qfrc_passive = ALLOC(nv);
  d->actuator_force = ALLOC(nu); d->qfrc_actuator = ALLOC(nv); d->qfrc_smooth = ALLOC(nv); d->qacc_smooth = ALLOC(nv);
  d->qfrc_constraint = ALLOC(nv);
  d->cacc = ALLOC(6 * nb); d->sensordata = ALLOC(v->nsensordata);
  d->contact = (orc_contact*)calloc((size_t)m->maxcon, sizeof(orc_contact));
  int me = m->maxefc;
  d->efc_type = (int*)calloc((size_t)me, sizeof(int)); d->efc_id = (int*)calloc((size_t)me, sizeof(int));
  d->efc_J = ALLOC(me * nv); d->efc_pos = ALLOC(me); d->efc_margin = ALLOC(me); d->efc_D = ALLOC(me); d->efc_R = ALLOC(me);
  d->efc_aref = ALLOC(me); d->efc_vel = ALLOC(me); d->efc_force = ALLOC(me); d->efc_diagApprox = ALLOC(me);
  d->arena_cap = (size_t)(v->nsensordata + 64 * (nq + nv + nu + 8) + 40 * nb + 8 * nv * nv + 12 * me + 16 * nv * 6 + 4096);
  d->arena = ALLOC(d->arena_cap); d->arena_top = 0;
  return d;
}
void orc_data_free(orc_data* d) {
  if (!d) return;
  double** p[] = {&d->qpos, &d->qvel, &d->ctrl, &d->qacc, &d->qacc_warmstart, &d->xpos, &d->xquat, &d->xmat, &d->xipos,
                  &d->ximat, &d->xanchor, &d->xaxis, &d->geom_xpos, &d->geom_xmat, &d->site_xpos, &d->site_xmat,
                  &d->subtree_com, &d->cinert, &d->crb, &d->cdof, &d->qM, &d->qLD, &d->qLDiagInv, &d->qH, &d->qHDiagInv,
                  &d->ten_length, &d->ten_J, &d->actuator_length, &d->actuator_moment, &d->cvel, &d->cdof_dot,
                  &d->ten_velocity, &d->actuator_velocity, &d->qfrc_bias, &d->qfrc_passive, &d->actuator_force,
                  &d->qfrc_actuator, &d->qfrc_smooth, &d->qacc_smooth, &d->qfrc_constraint, &d->efc_J, &d->efc_pos,
                  &d->efc_margin, &d->efc_D, &d->efc_R, &d->efc_aref, &d->efc_vel, &d->efc_force, &d->efc_diagApprox,
                  &d->cacc, &d->sensordata};
  for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(*p[i]);
  free(d->arena); free(d->contact); free(d->efc_type); free(d->efc_id); free(d);
}

/* mj_resetData / mj_resetDataKeyframe */
void orc_reset(const orc_model* m, orc_data* d, int key) {
  const b2m_view* v = &m->v;
  memcpy(d->qpos, key >= 0 ? v->key_qpos + (size_t)key * v->nq : v->qpos0, sizeof(double) * v->nq);
  for (int i = 0; i < v->nv; i++) { d->qvel[i] = key >= 0 ? v->key_qvel[(size_t)key * v->nv + i] : 0; d->qacc[i] = 0; d->qacc_warmstart[i] = 0; }
  for (int i = 0; i < v->nu; i++) d->ctrl[i] = key >= 0 ? v->key_ctrl[(size_t)key * v->nu + i] : 0;
  d->time = key >= 0 ? v->key_time[key] : 0;
  d->ncon = d->nefc = 0;
  d->warn_bad_qpos = d->warn_bad_qvel = d->warn_bad_qacc = d->warn_overflow = 0;
}

/* ------------------------------------------------------------------ position stage */
/* mj_kinematics (engine_core_smooth.c) */
static void orc_kinematics(const b2m_view* v, orc_data* d) {
  v3_zero(d->xpos); d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  q_tomat(d->xmat, d->xquat); v3_zero(d->xipos); q_tomat(d->ximat, d->xquat);
  for (int i = 1; i < v->nbody; i++) {
    double xpos[3], xquat[4];
    int jntadr = v->body_jntadr[i], jntnum = v->body_jntnum[i], pid = v->body_parentid[i];
    if (jntnum == 1 && v->jnt_type[jntadr] == JNT_FREE) {
      int qadr = v->jnt_qposadr[jntadr];
      v3_copy(xpos, d->qpos + qadr);
      memcpy(xquat, d->qpos + qadr + 3, 4 * sizeof(double));
      q_normalize(xquat);
      v3_copy(d->xanchor + 3 * jntadr, xpos);
      v3_copy(d->xaxis + 3 * jntadr, v->jnt_axis + 3 * jntadr);
    } else {
      if (pid) {
        m3_mulvec(xpos, d->xmat + 9 * pid, v->body_pos + 3 * i);
        v3_addto(xpos, d->xpos + 3 * pid);
        q_mul(xquat, d->xquat + 4 * pid, v->body_quat + 4 * i);
      } else {
        v3_copy(xpos, v->body_pos + 3 * i);
        memcpy(xquat, v->body_quat + 4 * i, 4 * sizeof(double));
      }
      for (int j = 0; j < jntnum; j++) {
        int jid = jntadr + j, qadr = v->jnt_qposadr[jid];
        double xanchor[3], xaxis[3];
        q_rotvec(xaxis, v->jnt_axis + 3 * jid, xquat);
        q_rotvec(xanchor, v->jnt_pos + 3 * jid, xquat);
        v3_addto(xanchor, xpos);
        if (v->jnt_type[jid] == JNT_SLIDE) {
          v3_addtoscl(xpos, xaxis, d->qpos[qadr] - v->qpos0[qadr]);
        } else if (v->jnt_type[jid] == JNT_HINGE) {
          double qloc[4], vec[3];
          q_axisangle(qloc, v->jnt_axis + 3 * jid, d->qpos[qadr] - v->qpos0[qadr]);
          q_mul(xquat, xquat, qloc);
          q_rotvec(vec, v->jnt_pos + 3 * jid, xquat);
          v3_sub(xpos, xanchor, vec);
        }
        v3_copy(d->xanchor + 3 * jid, xanchor);
        v3_copy(d->xaxis + 3 * jid, xaxis);
      }
    }
    q_normalize(xquat);
    memcpy(d->xquat + 4 * i, xquat, 4 * sizeof(double));
    v3_copy(d->xpos + 3 * i, xpos);
    q_tomat(d->xmat + 9 * i, xquat);
  }
  /* mj_local2Global for inertial frames, geoms, sites */
  for (int i = 1; i < v->nbody; i++) {
    double q[4];
    m3_mulvec(d->xipos + 3 * i, d->xmat + 9 * i, v->body_ipos + 3 * i);
    v3_addto(d->xipos + 3 * i, d->xpos + 3 * i);
    q_mul(q, d->xquat + 4 * i, v->body_iquat + 4 * i);
    q_tomat(d->ximat + 9 * i, q);
  }
  for (int g = 0; g < v->ngeom; g++) {
    int b = v->geom_bodyid[g];
    double q[4];
    m3_mulvec(d->geom_xpos + 3 * g, d->xmat + 9 * b, v->geom_pos + 3 * g);
    v3_addto(d->geom_xpos + 3 * g, d->xpos + 3 * b);
    q_mul(q, d->xquat + 4 * b, v->geom_quat + 4 * g);
    q_tomat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < v->nsite; s++) {
    int b = v->site_bodyid[s];
    double q[4];
    m3_mulvec(d->site_xpos + 3 * s, d->xmat + 9 * b, v->site_pos + 3 * s);
    v3_addto(d->site_xpos + 3 * s, d->xpos + 3 * b);
    q_mul(q, d->xquat + 4 * b, v->site_quat + 4 * s);
    q_tomat(d->site_xmat + 9 * s, q);
  }
}

/* mj_comPos */
static void orc_compos(const b2m_view* v, orc_data* d) {
  int nb = v->nbody;
  memset(d->subtree_com, 0, sizeof(double) * 3 * nb);
  for (int i = nb - 1; i >= 0; i--) {
    v3_addtoscl(d->subtree_com + 3 * i, d->xipos + 3 * i, v->body_mass[i]);
    if (i) v3_addto(d->subtree_com + 3 * v->body_parentid[i], d->subtree_com + 3 * i);
    if (v->body_subtreemass[i] < ORC_MINVAL) v3_copy(d->subtree_com + 3 * i, d->xipos + 3 * i);
    else v3_scl(d->subtree_com + 3 * i, d->subtree_com + 3 * i, 1.0 / orc_max(ORC_MINVAL, v->body_subtreemass[i]));
  }
  memset(d->cinert, 0, sizeof(double) * 10);
  for (int i = 1; i < nb; i++) {
    double offset[3];
    v3_sub(offset, d->xipos + 3 * i, d->subtree_com + 3 * v->body_rootid[i]);
    sp_inert_com(d->cinert + 10 * i, v->body_inertia + 3 * i, d->ximat + 9 * i, offset, v->body_mass[i]);
  }
  for (int j = 0; j < v->njnt; j++) {
    int bi = v->jnt_bodyid[j], da = 6 * v->jnt_dofadr[j];
    double offset[3], axis[3];
    v3_sub(offset, d->subtree_com + 3 * v->body_rootid[bi], d->xanchor + 3 * j);
    switch (v->jnt_type[j]) {
      case JNT_FREE:
        memset(d->cdof + da, 0, sizeof(double) * 18);
        for (int i = 0; i < 3; i++) d->cdof[da + 3 + 7 * i] = 1;
        for (int i = 0; i < 3; i++) {
          axis[0] = d->xmat[9 * bi + i]; axis[1] = d->xmat[9 * bi + i + 3]; axis[2] = d->xmat[9 * bi + i + 6];
          sp_dof_com(d->cdof + da + 18 + 6 * i, axis, offset);
        }
        break;
      case JNT_SLIDE: sp_dof_com(d->cdof + da, d->xaxis + 3 * j, NULL); break;
      case JNT_HINGE: sp_dof_com(d->cdof + da, d->xaxis + 3 * j, offset); break;
      default: break;
    }
  }
}

/* mj_tendon (fixed tendons only) */
static void orc_tendon(const b2m_view* v, orc_data* d) {
  int nv = v->nv;
  for (int t = 0; t < v->ntendon; t++) {
    double L = 0;
    for (int k = 0; k < nv; k++) d->ten_J[t * nv + k] = 0;
    for (int w = v->tendon_adr[t]; w < v->tendon_adr[t] + v->tendon_num[t]; w++) {
      int j = v->wrap_jntid[w];
      L += v->wrap_coef[w] * d->qpos[v->jnt_qposadr[j]];
      d->ten_J[t * nv + v->jnt_dofadr[j]] = v->wrap_coef[w];
    }
    d->ten_length[t] = L;
  }
}

/* mj_crb: composite rigid body inertia -> dense symmetric qM */
static void orc_crb(const b2m_view* v, orc_data* d) {
  int nv = v->nv, nb = v->nbody;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * nb);
  for (int i = nb - 1; i > 0; i--)
    if (v->body_parentid[i] > 0)
      for (int k = 0; k < 10; k++) d->crb[10 * v->body_parentid[i] + k] += d->crb[10 * i + k];
  memset(d->qM, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    d->qM[i * nv + i] = v->dof_armature[i];
    sp_mul_inert(buf, d->crb + 10 * v->dof_bodyid[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = v->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->qM[i * nv + j] += s;
      if (j != i) d->qM[j * nv + i] = d->qM[i * nv + j];
    }
  }
}

/* mj_factorI-style L'DL on the ancestor sparsity (engine_core_smooth.c mj_factorM) */
static void orc_factor(const b2m_view* v, const double* M, double* LD, double* diaginv) {
  int nv = v->nv;
  memcpy(LD, M, sizeof(double) * nv * nv);
  for (int k = nv - 1; k >= 0; k--) {
    double Mkk = LD[k * nv + k];
    for (int i = v->dof_parentid[k]; i >= 0; i = v->dof_parentid[i]) {
      double tmp = LD[k * nv + i] / Mkk;
      for (int j = i; j >= 0; j = v->dof_parentid[j]) LD[i * nv + j] -= tmp * LD[k * nv + j];
      LD[k * nv + i] = tmp;
    }
    diaginv[k] = 1.0 / Mkk;
  }
}
/* mj_solveLD */
static void orc_solve_ld(const b2m_view* v, const double* LD, const double* diaginv, double* x) {
  int nv = v->nv;
  for (int i = nv - 1; i >= 0; i--)
    for (int j = v->dof_parentid[i]; j >= 0; j = v->dof_parentid[j]) x[j] -= LD[i * nv + j] * x[i];
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++)
    for (int j = v->dof_parentid[i]; j >= 0; j = v->dof_parentid[j]) x[i] -= LD[i * nv + j] * x[j];
}
/* mj_mulM on the ancestor sparsity */
static void orc_mul_m(const b2m_view* v, const orc_data* d, double* res, const double* vec) {
  int nv = v->nv;
  for (int i = 0; i < nv; i++) res[i] = 0;
  for (int i = 0; i < nv; i++) {
    res[i] += d->qM[i * nv + i] * vec[i];
    for (int j = v->dof_parentid[i]; j >= 0; j = v->dof_parentid[j]) {
      res[i] += d->qM[i * nv + j] * vec[j];
      res[j] += d->qM[i * nv + j] * vec[i];
    }
  }
}

/* mj_jac: point Jacobian of `body` at world `point` (3 x nv row-major each) */
void orc_jac(const orc_model* m, const orc_data* d, double* jacp, double* jacr, const double* point, int body) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  double offset[3];
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  v3_sub(offset, point, d->subtree_com + 3 * v->body_rootid[body]);
  while (body && !v->body_dofnum[body]) body = v->body_parentid[body];
  if (!body) return;
  for (int i = v->body_dofadr[body] + v->body_dofnum[body] - 1; i >= 0; i = v->dof_parentid[i]) {
    const double* cdof = d->cdof + 6 * i;
    if (jacr) { jacr[i] = cdof[0]; jacr[nv + i] = cdof[1]; jacr[2 * nv + i] = cdof[2]; }
    if (jacp) {
      double tmp[3];
      v3_cross(tmp, cdof, offset);
      jacp[i] = cdof[3] + tmp[0]; jacp[nv + i] = cdof[4] + tmp[1]; jacp[2 * nv + i] = cdof[5] + tmp[2];
    }
  }
}
void orc_jac_site(const orc_model* m, const orc_data* d, double* jacp, double* jacr, int site) {
  orc_jac(m, d, jacp, jacr, d->site_xpos + 3 * site, m->v.site_bodyid[site]);
}
void orc_jac_body(const orc_model* m, const orc_data* d, double* jacp, double* jacr, int body) {
  orc_jac(m, d, jacp, jacr, d->xpos + 3 * body, body);
}
void orc_jac_bodycom(const orc_model* m, const orc_data* d, double* jacp, double* jacr, int body) {
  orc_jac(m, d, jacp, jacr, d->xipos + 3 * body, body);
}
/* mj_jacSubtreeCom */
void orc_jac_subtreecom(const orc_model* m, const orc_data* d, double* jacp, int body) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  double* tmp = ALLOC(3 * nv);
  memset(jacp, 0, sizeof(double) * 3 * nv);
  for (int b = body; b < v->nbody; b++) {
    if (b > body && v->body_parentid[b] < body) break;
    orc_jac_bodycom(m, d, tmp, NULL, b);
    for (int k = 0; k < 3 * nv; k++) jacp[k] += tmp[k] * v->body_mass[b];
  }
  for (int k = 0; k < 3 * nv; k++) jacp[k] *= 1.0 / v->body_subtreemass[body];
  free(tmp);
}

/* mj_transmission (joint and site transmissions) */
static void orc_transmission(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  double* jac = SCRATCH(6 * nv);
  for (int i = 0; i < v->nu; i++) {
    double* moment = d->actuator_moment + i * nv;
    const double* gear = v->actuator_gear + 6 * i;
    for (int k = 0; k < nv; k++) moment[k] = 0;
    int id = v->actuator_trnid[i];
    if (v->actuator_trntype[i] == TRN_JOINT) {
      d->actuator_length[i] = d->qpos[v->jnt_qposadr[id]] * gear[0];
      moment[v->jnt_dofadr[id]] = gear[0];
    } else {
      double wrench[6];
      m3_mulvec(wrench, d->site_xmat + 9 * id, gear);
      m3_mulvec(wrench + 3, d->site_xmat + 9 * id, gear + 3);
      orc_jac_site(m, d, jac, jac + 3 * nv, id);
      for (int k = 0; k < nv; k++) {
        double s = 0;
        for (int r = 0; r < 6; r++) s += jac[r * nv + k] * wrench[r];
        moment[k] = s;
      }
      d->actuator_length[i] = 0;
    }
  }
}

/* ------------------------------------------------------------------ collision */
static int raw_plane_sphere(orc_contact* con, double margin, const double* pos1, const double* mat1,
                            const double* pos2, double radius) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, tmp[3];
  v3_sub(tmp, pos2, pos1);
  double cdist = v3_dot(tmp, normal);
  if (cdist > margin + radius) return 0;
  con->dist = cdist - radius;
  v3_addscl(con->pos, pos2, normal, -con->dist / 2 - radius);
  v3_copy(con->frame, normal);
  v3_zero(con->frame + 3);
  return 1;
}
static int raw_sphere_sphere(orc_contact* con, double margin, const double* pos1, const double* mat1, double r1,
                             const double* pos2, const double* mat2, double r2) {
  double dif[3];
  v3_sub(dif, pos2, pos1);
  double cdist_sq = v3_dot(dif, dif), mind = margin + r1 + r2;
  if (cdist_sq > mind * mind) return 0;
  v3_copy(con->frame, dif);
  double cdist = v3_normalize(con->frame);
  con->dist = cdist - r1 - r2;
  if (cdist < ORC_MINVAL) {
    double a1[3] = {mat1[2], mat1[5], mat1[8]}, a2[3] = {mat2[2], mat2[5], mat2[8]};
    v3_cross(con->frame, a1, a2);
    v3_normalize(con->frame);
  }
  v3_addscl(con->pos, pos1, con->frame, r1 + 0.5 * con->dist);
  v3_zero(con->frame + 3);
  return 1;
}
static int raw_capsule_capsule(orc_contact* con, double margin, const double* pos1, const double* mat1, const double* size1,
                               const double* pos2, const double* mat2, const double* size2) {
  double axis1[3] = {mat1[2], mat1[5], mat1[8]}, axis2[3] = {mat2[2], mat2[5], mat2[8]}, dif[3], vec1[3], vec2[3];
  v3_sub(dif, pos1, pos2);
  double ma = v3_dot(axis1, axis1), mb = -v3_dot(axis1, axis2), mc = v3_dot(axis2, axis2);
  double u = -v3_dot(axis1, dif), w = v3_dot(axis2, dif);
  double det = ma * mc - mb * mb;
  double l1 = size1[1], l2 = size2[1];
  if (fabs(det) >= ORC_MINVAL) {
    double x1 = (mc * u - mb * w) / det, x2 = (ma * w - mb * u) / det;
    if (x1 > l1) { x1 = l1; x2 = (w - mb * l1) / mc; }
    else if (x1 < -l1) { x1 = -l1; x2 = (w + mb * l1) / mc; }
    if (x2 > l2) { x2 = l2; x1 = orc_clip((u - mb * l2) / ma, -l1, l1); }
    else if (x2 < -l2) { x2 = -l2; x1 = orc_clip((u + mb * l2) / ma, -l1, l1); }
    v3_addscl(vec1, pos1, axis1, x1);
    v3_addscl(vec2, pos2, axis2, x2);
    return raw_sphere_sphere(con, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  }
  /* parallel axes: test the four end-point configurations, keep at most two */
  int n = 0;
  double x;
  v3_addscl(vec1, pos1, axis1, l1);
  x = orc_clip((w - mb * l1) / mc, -l2, l2);
  v3_addscl(vec2, pos2, axis2, x);
  n += raw_sphere_sphere(con + n, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  v3_addscl(vec1, pos1, axis1, -l1);
  x = orc_clip((w + mb * l1) / mc, -l2, l2);
  v3_addscl(vec2, pos2, axis2, x);
  n += raw_sphere_sphere(con + n, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  if (n == 2) return n;
  v3_addscl(vec2, pos2, axis2, l2);
  x = orc_clip((u - mb * l2) / ma, -l1, l1);
  v3_addscl(vec1, pos1, axis1, x);
  n += raw_sphere_sphere(con + n, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  if (n == 2) return n;
  v3_addscl(vec2, pos2, axis2, -l2);
  x = orc_clip((u + mb * l2) / ma, -l1, l1);
  v3_addscl(vec1, pos1, axis1, x);
  n += raw_sphere_sphere(con + n, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  return n;
}

/* narrow phase for one candidate pair (engine_collision_primitive.c routines) */
static int orc_collide_pair(const b2m_view* v, const orc_data* d, int g1, int g2, double margin, orc_contact* con) {
  const double *pos1 = d->geom_xpos + 3 * g1, *mat1 = d->geom_xmat + 9 * g1, *size1 = v->geom_size + 3 * g1;
  const double *pos2 = d->geom_xpos + 3 * g2, *mat2 = d->geom_xmat + 9 * g2, *size2 = v->geom_size + 3 * g2;
  int t1 = v->geom_type[g1], t2 = v->geom_type[g2];
  if (t1 == GEOM_PLANE && t2 == GEOM_SPHERE) return raw_plane_sphere(con, margin, pos1, mat1, pos2, size2[0]);
  if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) { /* mjc_PlaneCapsule */
    double axis[3] = {mat2[2], mat2[5], mat2[8]}, seg[3], p[3];
    v3_scl(seg, axis, size2[1]);
    v3_add(p, pos2, seg);
    int n1 = raw_plane_sphere(con, margin, pos1, mat1, p, size2[0]);
    v3_sub(p, pos2, seg);
    int n2 = raw_plane_sphere(con + n1, margin, pos1, mat1, p, size2[0]);
    if (n1) v3_copy(con[0].frame + 3, axis);
    if (n2) v3_copy(con[n1].frame + 3, axis);
    return n1 + n2;
  }
  if (t1 == GEOM_PLANE && t2 == GEOM_BOX) { /* mjc_PlaneBox */
    double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
    v3_sub(dif, pos2, pos1);
    double dist = v3_dot(dif, normal);
    int cnt = 0;
    for (int i = 0; i < 8; i++) {
      double vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])}, corner[3];
      m3_mulvec(corner, mat2, vec);
      double ldist = v3_dot(normal, corner);
      if (dist + ldist > margin || ldist > 0) continue;
      con[cnt].dist = dist + ldist;
      v3_copy(con[cnt].frame, normal);
      v3_zero(con[cnt].frame + 3);
      v3_addto(corner, pos2);
      v3_addscl(con[cnt].pos, corner, normal, -con[cnt].dist / 2);
      if (++cnt >= 4) return 4;
    }
    return cnt;
  }
  if (t1 == GEOM_PLANE && t2 == GEOM_ELLIPSOID) { /* mjc_PlaneConvex with the ellipsoid support map */
    double normal[3] = {mat1[2], mat1[5], mat1[8]}, dirl[3], neg[3] = {-normal[0], -normal[1], -normal[2]}, sup[3], tmp[3];
    m3_multvec(dirl, mat2, neg);
    double s[3] = {size2[0] * dirl[0], size2[1] * dirl[1], size2[2] * dirl[2]};
    double nrm = v3_norm(s);
    if (nrm < ORC_MINVAL) return 0;
    double loc[3] = {size2[0] * s[0] / nrm, size2[1] * s[1] / nrm, size2[2] * s[2] / nrm};
    m3_mulvec(sup, mat2, loc);
    v3_addto(sup, pos2);
    v3_sub(tmp, sup, pos1);
    double dist = v3_dot(tmp, normal);
    if (dist > margin) return 0;
    con->dist = dist;
    v3_addscl(con->pos, sup, normal, -0.5 * dist);
    v3_copy(con->frame, normal);
    v3_zero(con->frame + 3);
    return 1;
  }
  if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) return raw_sphere_sphere(con, margin, pos1, mat1, size1[0], pos2, mat2, size2[0]);
  if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) { /* mjc_SphereCapsule */
    double axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3];
    v3_sub(vec, pos1, pos2);
    double x = orc_clip(v3_dot(axis, vec), -size2[1], size2[1]);
    v3_addscl(vec, pos2, axis, x);
    return raw_sphere_sphere(con, margin, pos1, mat1, size1[0], vec, mat2, size2[0]);
  }
  if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) return raw_capsule_capsule(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  return 0;
}

/* mj_collision over the statically filtered pair list */
static void orc_collision(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  d->ncon = 0;
  for (int p = 0; p < v->npair; p++) {
    int g1 = v->pair_geom1[p], g2 = v->pair_geom2[p];
    double margin = v->pair_margin[p];
    /* bounding-sphere / plane-distance prefilter (mj_collideGeoms) */
    if (v->geom_type[g1] == GEOM_PLANE) {
      const double* mat1 = d->geom_xmat + 9 * g1;
      double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
      v3_sub(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (v3_dot(dif, normal) > margin + v->geom_rbound[g2]) continue;
    } else {
      double dif[3], bound = margin + v->geom_rbound[g1] + v->geom_rbound[g2];
      v3_sub(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (v3_dot(dif, dif) > bound * bound) continue;
    }
    orc_contact con[4];
    memset(con, 0, sizeof(con));
    int num = orc_collide_pair(v, d, g1, g2, margin, con);
    for (int i = 0; i < num; i++) {
      if (d->ncon >= m->maxcon) { d->warn_overflow = 1; return; }
      orc_contact* c = &con[i];
      c->includemargin = margin - v->pair_gap[p];
      c->dim = v->pair_dim[p];
      memcpy(c->friction, v->pair_friction + 5 * p, 5 * sizeof(double));
      memcpy(c->solref, v->pair_solref + 2 * p, 2 * sizeof(double));
      memcpy(c->solimp, v->pair_solimp + 5 * p, 5 * sizeof(double));
      c->geom1 = g1; c->geom2 = g2; c->pair = p;
      c->exclude = (c->dist >= c->includemargin);
      make_frame(c->frame);
      d->contact[d->ncon++] = *c;
    }
  }
}

/* ------------------------------------------------------------------ constraints */
static int orc_add_row(const orc_model* m, orc_data* d, int type, int id, double pos, double margin) {
  if (d->nefc >= m->maxefc) { d->warn_overflow = 1; return -1; }
  int r = d->nefc++;
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin;
  memset(d->efc_J + (size_t)r * m->v.nv, 0, sizeof(double) * m->v.nv);
  return r;
}

/* mj_makeConstraint: limits (joint, tendon) then contacts (pyramidal / frictionless) */
static void orc_make_constraint(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  d->nefc = 0;
  for (int i = 0; i < v->njnt; i++) {
    if (!v->jnt_limited[i] || (v->jnt_type[i] != JNT_SLIDE && v->jnt_type[i] != JNT_HINGE)) continue;
    double value = d->qpos[v->jnt_qposadr[i]], margin = v->jnt_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (v->jnt_range[2 * i + (side + 1) / 2] - value);
      if (dist < margin) {
        int r = orc_add_row(m, d, EFC_LIMIT_JOINT, i, dist, margin);
        if (r >= 0) d->efc_J[(size_t)r * nv + v->jnt_dofadr[i]] = -(double)side;
      }
    }
  }
  for (int i = 0; i < v->ntendon; i++) {
    if (!v->tendon_limited[i]) continue;
    double value = d->ten_length[i], margin = v->tendon_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (v->tendon_range[2 * i + (side + 1) / 2] - value);
      if (dist < margin) {
        int r = orc_add_row(m, d, EFC_LIMIT_TENDON, i, dist, margin);
        if (r >= 0) for (int k = 0; k < nv; k++) d->efc_J[(size_t)r * nv + k] = -side * d->ten_J[i * nv + k];
      }
    }
  }
  double* j1 = SCRATCH(3 * nv);
  double* j2 = SCRATCH(3 * nv);
  double* jc = SCRATCH(3 * nv);
  for (int c = 0; c < d->ncon; c++) {
    orc_contact* con = d->contact + c;
    if (con->exclude) continue;
    int b1 = v->geom_bodyid[con->geom1], b2 = v->geom_bodyid[con->geom2];
    orc_jac(m, d, j1, NULL, con->pos, b1);
    orc_jac(m, d, j2, NULL, con->pos, b2);
    for (int k = 0; k < 3 * nv; k++) j2[k] -= j1[k]; /* jacdif = jac2 - jac1 */
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < nv; k++)
        jc[r * nv + k] = con->frame[3 * r] * j2[k] + con->frame[3 * r + 1] * j2[nv + k] + con->frame[3 * r + 2] * j2[2 * nv + k];
    if (con->dim == 1) {
      int r = orc_add_row(m, d, EFC_CONTACT_FRICTIONLESS, c, con->dist, con->includemargin);
      if (r >= 0) memcpy(d->efc_J + (size_t)r * nv, jc, sizeof(double) * nv);
    } else {
      for (int k = 1; k < con->dim; k++) {
        int r0 = orc_add_row(m, d, EFC_CONTACT_PYRAMIDAL, c, con->dist, con->includemargin);
        int r1 = orc_add_row(m, d, EFC_CONTACT_PYRAMIDAL, c, con->dist, con->includemargin);
        if (r0 < 0 || r1 < 0) break;
        for (int q = 0; q < nv; q++) {
          d->efc_J[(size_t)r0 * nv + q] = jc[q] + con->friction[k - 1] * jc[k * nv + q];
          d->efc_J[(size_t)r1 * nv + q] = jc[q] - con->friction[k - 1] * jc[k * nv + q];
        }
      }
    }
  }
}

/* getimpedance (engine_core_constraint.c) */
static double orc_impedance(const double* solimp, double pos, double margin) {
  double dmin = orc_clip(solimp[0], ORC_MINIMP, ORC_MAXIMP), dmax = orc_clip(solimp[1], ORC_MINIMP, ORC_MAXIMP);
  double width = orc_max(ORC_MINVAL, solimp[2]), mid = orc_clip(solimp[3], ORC_MINIMP, ORC_MAXIMP), power = orc_max(1, solimp[4]);
  if (dmin == dmax || width <= ORC_MINVAL) return 0.5 * (dmin + dmax);
  double x = (pos - margin) / width;
  if (x < 0) x = -x;
  if (x >= 1) return dmax;
  if (x == 0) return dmin;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = (1 / pow(mid, power - 1)) * pow(x, power);
  else y = 1 - (1 / pow(1 - mid, power - 1)) * pow(1 - x, power);
  return dmin + y * (dmax - dmin);
}

/* mj_diagApprox + mj_makeImpedance + mj_referenceConstraint */
static void orc_make_impedance(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  for (int i = 0; i < d->nefc; i++) {
    const double *solref, *solimp;
    int id = d->efc_id[i];
    switch (d->efc_type[i]) {
      case EFC_LIMIT_JOINT:
        solref = v->jnt_solref + 2 * id; solimp = v->jnt_solimp + 5 * id;
        d->efc_diagApprox[i] = v->dof_invweight0[v->jnt_dofadr[id]];
        break;
      case EFC_LIMIT_TENDON:
        solref = v->tendon_solref + 2 * id; solimp = v->tendon_solimp + 5 * id;
        d->efc_diagApprox[i] = v->tendon_invweight0[id];
        break;
      default: {
        const orc_contact* con = d->contact + id;
        int b1 = v->geom_bodyid[con->geom1], b2 = v->geom_bodyid[con->geom2];
        double tran = v->body_invweight0[2 * b1] + v->body_invweight0[2 * b2];
        solref = con->solref; solimp = con->solimp;
        if (d->efc_type[i] == EFC_CONTACT_FRICTIONLESS) d->efc_diagApprox[i] = tran;
        else {
          /* rows of one contact are consecutive; j = row index within the contact */
          int j = 0;
          while (i - j - 1 >= 0 && d->efc_type[i - j - 1] == EFC_CONTACT_PYRAMIDAL && d->efc_id[i - j - 1] == id) j++;
          double fri = con->friction[j / 2];
          d->efc_diagApprox[i] = tran + fri * fri * tran;
        }
      }
    }
    double pos = d->efc_pos[i], margin = d->efc_margin[i];
    double imp = orc_impedance(solimp, pos, margin);
    double dmax = orc_clip(solimp[1], ORC_MINIMP, ORC_MAXIMP);
    double K, B;
    if (solref[0] > 0) {
      double timeconst = orc_max(solref[0], 2 * v->timestep), dampratio = solref[1];
      K = 1 / orc_max(ORC_MINVAL, dmax * dmax * timeconst * timeconst * dampratio * dampratio);
      B = 2 / orc_max(ORC_MINVAL, dmax * timeconst);
    } else {
      K = -solref[0] / orc_max(ORC_MINVAL, dmax * dmax);
      B = -solref[1] / orc_max(ORC_MINVAL, dmax);
    }
    d->efc_R[i] = orc_max(ORC_MINVAL, (1 - imp) * d->efc_diagApprox[i] / imp);
    /* mj_referenceConstraint */
    double vel = 0;
    for (int k = 0; k < nv; k++) vel += d->efc_J[(size_t)i * nv + k] * d->qvel[k];
    d->efc_vel[i] = vel;
    d->efc_aref[i] = -B * vel - K * imp * (pos - margin);
  }
  /* pyramidal contacts: all rows share R = 2 mu^2 R(first row), mu = friction[0] (impratio 1) */
  for (int i = 0; i < d->nefc; i++) {
    if (d->efc_type[i] != EFC_CONTACT_PYRAMIDAL) continue;
    const orc_contact* con = d->contact + d->efc_id[i];
    int rows = 2 * (con->dim - 1);
    double Rpy = 2 * con->friction[0] * con->friction[0] * d->efc_R[i];
    for (int j = 0; j < rows; j++) d->efc_R[i + j] = Rpy;
    i += rows - 1;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

/* ------------------------------------------------------------------ velocity stage */
/* mj_comVel */
static void orc_comvel(const b2m_view* v, orc_data* d) {
  memset(d->cvel, 0, sizeof(double) * 6);
  for (int i = 1; i < v->nbody; i++) {
    double cvel[6], tmp[6];
    int bda = v->body_dofadr[i];
    memcpy(cvel, d->cvel + 6 * v->body_parentid[i], sizeof(cvel));
    for (int j = 0; j < v->body_dofnum[i]; j++) {
      if (v->jnt_type[v->dof_jntid[bda + j]] == JNT_FREE) {
        memset(d->cdof_dot + 6 * bda, 0, sizeof(double) * 18);
        for (int k = 0; k < 6; k++) {
          tmp[k] = 0;
          for (int q = 0; q < 3; q++) tmp[k] += d->cdof[6 * (bda + q) + k] * d->qvel[bda + q];
          cvel[k] += tmp[k];
        }
        for (int k = 0; k < 3; k++) sp_cross_motion(d->cdof_dot + 6 * (bda + 3 + k), cvel, d->cdof + 6 * (bda + 3 + k));
        for (int k = 0; k < 6; k++) {
          tmp[k] = 0;
          for (int q = 3; q < 6; q++) tmp[k] += d->cdof[6 * (bda + q) + k] * d->qvel[bda + q];
          cvel[k] += tmp[k];
        }
        j += 5;
      } else {
        sp_cross_motion(d->cdof_dot + 6 * (bda + j), cvel, d->cdof + 6 * (bda + j));
        for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6 * (bda + j) + k] * d->qvel[bda + j];
      }
    }
    memcpy(d->cvel + 6 * i, cvel, sizeof(cvel));
  }
}

/* mj_inertiaBoxFluidModel (engine_passive.c) */
static void orc_fluid_body(const orc_model* m, orc_data* d, int i, double* jacp, double* jacr) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  const double* inertia = v->body_inertia + 3 * i;
  double mass = v->body_mass[i], box[3], lvel[6], wind[6] = {0, 0, 0, v->wind[0], v->wind[1], v->wind[2]}, lwind[6], lfrc[6] = {0}, bfrc[6];
  box[0] = sqrt(orc_max(ORC_MINVAL, inertia[1] + inertia[2] - inertia[0]) / mass * 6.0);
  box[1] = sqrt(orc_max(ORC_MINVAL, inertia[0] + inertia[2] - inertia[1]) / mass * 6.0);
  box[2] = sqrt(orc_max(ORC_MINVAL, inertia[0] + inertia[1] - inertia[2]) / mass * 6.0);
  const double* com = d->subtree_com + 3 * v->body_rootid[i];
  sp_transform(lvel, d->cvel + 6 * i, 0, d->xipos + 3 * i, com, d->ximat + 9 * i);
  sp_transform(lwind, wind, 0, d->xipos + 3 * i, com, d->ximat + 9 * i);
  lvel[3] -= lwind[3]; lvel[4] -= lwind[4]; lvel[5] -= lwind[5];
  if (v->viscosity > 0) {
    double diam = (box[0] + box[1] + box[2]) / 3.0;
    v3_scl(lfrc, lvel, -ORC_PI * diam * diam * diam * v->viscosity);
    v3_scl(lfrc + 3, lvel + 3, -3.0 * ORC_PI * diam * v->viscosity);
  }
  if (v->density > 0) {
    lfrc[3] -= 0.5 * v->density * box[1] * box[2] * fabs(lvel[3]) * lvel[3];
    lfrc[4] -= 0.5 * v->density * box[0] * box[2] * fabs(lvel[4]) * lvel[4];
    lfrc[5] -= 0.5 * v->density * box[0] * box[1] * fabs(lvel[5]) * lvel[5];
    lfrc[0] -= v->density * box[0] * (pow(box[1], 4) + pow(box[2], 4)) * fabs(lvel[0]) * lvel[0] / 64.0;
    lfrc[1] -= v->density * box[1] * (pow(box[0], 4) + pow(box[2], 4)) * fabs(lvel[1]) * lvel[1] / 64.0;
    lfrc[2] -= v->density * box[2] * (pow(box[0], 4) + pow(box[1], 4)) * fabs(lvel[2]) * lvel[2] / 64.0;
  }
  m3_mulvec(bfrc, d->ximat + 9 * i, lfrc);
  m3_mulvec(bfrc + 3, d->ximat + 9 * i, lfrc + 3);
  /* mj_applyFT at the body CoM */
  orc_jac(m, d, jacp, jacr, d->xipos + 3 * i, i);
  for (int k = 0; k < nv; k++)
    d->qfrc_passive[k] += jacp[k] * bfrc[3] + jacp[nv + k] * bfrc[4] + jacp[2 * nv + k] * bfrc[5] +
                          jacr[k] * bfrc[0] + jacr[nv + k] * bfrc[1] + jacr[2 * nv + k] * bfrc[2];
}

/* mj_passive */
static void orc_passive(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  for (int k = 0; k < nv; k++) d->qfrc_passive[k] = 0;
  for (int j = 0; j < v->njnt; j++) {
    double k = v->jnt_stiffness[j];
    if (k == 0) continue;
    int pa = v->jnt_qposadr[j], da = v->jnt_dofadr[j];
    if (v->jnt_type[j] == JNT_SLIDE || v->jnt_type[j] == JNT_HINGE)
      d->qfrc_passive[da] -= k * (d->qpos[pa] - v->qpos_spring[pa]);
    else if (v->jnt_type[j] == JNT_FREE) {
      double dif[3], q[4], qs[4];
      for (int t = 0; t < 3; t++) d->qfrc_passive[da + t] -= k * (d->qpos[pa + t] - v->qpos_spring[pa + t]);
      memcpy(q, d->qpos + pa + 3, sizeof(q)); q_normalize(q);
      q_neg(qs, v->qpos_spring + pa + 3);
      double qd[4]; q_mul(qd, qs, q); q_tovel(dif, qd, 1);
      for (int t = 0; t < 3; t++) d->qfrc_passive[da + 3 + t] -= k * dif[t];
    }
  }
  for (int k = 0; k < nv; k++) d->qfrc_passive[k] -= v->dof_damping[k] * d->qvel[k];
  for (int t = 0; t < v->ntendon; t++) {
    double stiff = v->tendon_stiffness[t], damp = v->tendon_damping[t], frc = 0;
    if (stiff == 0 && damp == 0) continue;
    double lo = v->tendon_lengthspring[2 * t], hi = v->tendon_lengthspring[2 * t + 1], L = d->ten_length[t];
    if (L > hi) frc = stiff * (hi - L); else if (L < lo) frc = stiff * (lo - L);
    frc -= damp * d->ten_velocity[t];
    for (int k = 0; k < nv; k++) d->qfrc_passive[k] += d->ten_J[t * nv + k] * frc;
  }
  if (v->density > 0 || v->viscosity > 0) {
    double* jacp = SCRATCH(3 * nv);
    double* jacr = SCRATCH(3 * nv);
    for (int i = 1; i < v->nbody; i++)
      if (v->body_mass[i] >= ORC_MINVAL) orc_fluid_body(m, d, i, jacp, jacr);
  }
}

/* mj_rne with flg_acc = 0: bias forces */
static void orc_rne(const b2m_view* v, orc_data* d) {
  int nb = v->nbody, nv = v->nv;
  double* cacc = SCRATCH(6 * nb);
  double* cfrc = SCRATCH(6 * nb);
  cacc[3] = -v->gravity[0]; cacc[4] = -v->gravity[1]; cacc[5] = -v->gravity[2];
  for (int i = 1; i < nb; i++) {
    int bda = v->body_dofadr[i];
    double tmp[6] = {0}, tmp1[6];
    for (int j = 0; j < v->body_dofnum[i]; j++)
      for (int k = 0; k < 6; k++) tmp[k] += d->cdof_dot[6 * (bda + j) + k] * d->qvel[bda + j];
    for (int k = 0; k < 6; k++) cacc[6 * i + k] = cacc[6 * v->body_parentid[i] + k] + tmp[k];
    sp_mul_inert(cfrc + 6 * i, d->cinert + 10 * i, cacc + 6 * i);
    sp_mul_inert(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    sp_cross_force(tmp1, d->cvel + 6 * i, tmp);
    for (int k = 0; k < 6; k++) cfrc[6 * i + k] += tmp1[k];
  }
  for (int i = nb - 1; i > 0; i--)
    if (v->body_parentid[i])
      for (int k = 0; k < 6; k++) cfrc[6 * v->body_parentid[i] + k] += cfrc[6 * i + k];
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += d->cdof[6 * i + k] * cfrc[6 * v->dof_bodyid[i] + k];
    d->qfrc_bias[i] = s;
  }
}

/* ------------------------------------------------------------------ acceleration stage */
/* mj_fwdActuation (dyntype none, gaintype fixed, biastype none/affine) */
static void orc_actuation(const b2m_view* v, orc_data* d) {
  int nv = v->nv;
  for (int k = 0; k < nv; k++) d->qfrc_actuator[k] = 0;
  for (int i = 0; i < v->nu; i++) {
    double ctrl = d->ctrl[i];
    if (v->actuator_ctrllimited[i]) ctrl = orc_clip(ctrl, v->actuator_ctrlrange[2 * i], v->actuator_ctrlrange[2 * i + 1]);
    const double* bp = v->actuator_biasprm + 3 * i;
    double force = v->actuator_gainprm[i] * ctrl + bp[0] + bp[1] * d->actuator_length[i] + bp[2] * d->actuator_velocity[i];
    if (v->actuator_forcelimited[i]) force = orc_clip(force, v->actuator_forcerange[2 * i], v->actuator_forcerange[2 * i + 1]);
    if (v->actuator_disabled[i]) force = 0;
    d->actuator_force[i] = force;
  }
  for (int i = 0; i < v->nu; i++)
    for (int k = 0; k < nv; k++) d->qfrc_actuator[k] += d->actuator_moment[i * nv + k] * d->actuator_force[i];
}

/* mj_constraintUpdate restricted to inequality rows: forces, cost, qfrc_constraint */
static double orc_constraint_update(const orc_model* m, orc_data* d, const double* jar, int write_qfrc) {
  int nv = m->v.nv;
  double cost = 0;
  for (int i = 0; i < d->nefc; i++) {
    if (jar[i] >= 0) d->efc_force[i] = 0;
    else { d->efc_force[i] = -d->efc_D[i] * jar[i]; cost += 0.5 * d->efc_D[i] * jar[i] * jar[i]; }
  }
  if (write_qfrc) {
    for (int k = 0; k < nv; k++) d->qfrc_constraint[k] = 0;
    for (int i = 0; i < d->nefc; i++) {
      if (d->efc_force[i] == 0) continue;
      for (int k = 0; k < nv; k++) d->qfrc_constraint[k] += d->efc_J[(size_t)i * nv + k] * d->efc_force[i];
    }
  }
  return cost;
}

typedef struct { double alpha, cost, deriv[2]; } ls_pnt;
typedef struct {
  const orc_model* m; orc_data* d;
  double *Jaref, *Jv, *Ma, *Mv, *grad, *Mgrad, *search, *quad, *H;
  double quadGauss[3], cost, gauss;
  int LSiter;
} newton_ctx;

/* PrimalEval */
static void ls_eval(newton_ctx* c, double alpha, ls_pnt* p) {
  c->LSiter++;
  double qt[3] = {c->quadGauss[0], c->quadGauss[1], c->quadGauss[2]};
  for (int i = 0; i < c->d->nefc; i++) {
    double x = c->Jaref[i] + alpha * c->Jv[i];
    if (x < 0) { qt[0] += c->quad[3 * i]; qt[1] += c->quad[3 * i + 1]; qt[2] += c->quad[3 * i + 2]; }
  }
  p->alpha = alpha;
  p->cost = alpha * alpha * qt[2] + alpha * qt[1] + qt[0];
  p->deriv[0] = 2 * alpha * qt[2] + qt[1];
  p->deriv[1] = 2 * qt[2];
  if (p->deriv[1] <= 0) p->deriv[1] = ORC_MINVAL;
}
static int ls_update_bracket(newton_ctx* c, ls_pnt* p, const ls_pnt cand[3], ls_pnt* pnext) {
  int flag = 0;
  for (int i = 0; i < 3; i++) {
    if (p->deriv[0] < 0 && cand[i].deriv[0] < 0 && p->deriv[0] < cand[i].deriv[0]) { *p = cand[i]; flag = 1; }
    else if (p->deriv[0] > 0 && cand[i].deriv[0] > 0 && p->deriv[0] > cand[i].deriv[0]) { *p = cand[i]; flag = 2; }
  }
  if (flag) ls_eval(c, p->alpha - p->deriv[0] / p->deriv[1], pnext);
  return flag;
}
/* PrimalSearch: exact line search on the piecewise-quadratic cost (engine_solver.c) */
static double ls_search(newton_ctx* c) {
  const b2m_view* v = &c->m->v;
  orc_data* d = c->d;
  int nv = v->nv, nefc = d->nefc;
  ls_pnt p0, p1, p2, pmid, p1next, p2next;
  c->LSiter = 0;
  double snorm = 0;
  for (int k = 0; k < nv; k++) snorm += c->search[k] * c->search[k];
  snorm = sqrt(snorm);
  if (snorm < ORC_MINVAL) return 0;
  double scale = 1 / (v->meaninertia * (nv > 1 ? nv : 1));
  double gtol = v->tolerance * v->ls_tolerance * snorm / scale;
  orc_mul_m(v, d, c->Mv, c->search);
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * c->search[k];
    c->Jv[i] = s;
  }
  /* PrimalPrepare */
  c->quadGauss[0] = c->gauss;
  double a = 0, b = 0, e = 0;
  for (int k = 0; k < nv; k++) { a += c->search[k] * c->Ma[k]; b += d->qfrc_smooth[k] * c->search[k]; e += c->search[k] * c->Mv[k]; }
  c->quadGauss[1] = a - b;
  c->quadGauss[2] = 0.5 * e;
  for (int i = 0; i < nefc; i++) {
    double DJ0 = d->efc_D[i] * c->Jaref[i];
    c->quad[3 * i] = 0.5 * c->Jaref[i] * DJ0;
    c->quad[3 * i + 1] = c->Jv[i] * DJ0;
    c->quad[3 * i + 2] = 0.5 * c->Jv[i] * d->efc_D[i] * c->Jv[i];
  }
  ls_eval(c, 0, &p0);
  ls_eval(c, p0.alpha - p0.deriv[0] / p0.deriv[1], &p1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  int dir = (p1.deriv[0] < 0 ? +1 : -1);
  int p2update = 0;
  int maxls = v->ls_iterations;
  while (p1.deriv[0] * dir <= -gtol && c->LSiter < maxls) {
    p2 = p1; p2update = 1;
    ls_eval(c, p1.alpha - p1.deriv[0] / p1.deriv[1], &p1);
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  }
  if (c->LSiter >= maxls) return p1.alpha;
  if (!p2update) return p1.alpha;
  p2next = p1;
  ls_eval(c, p1.alpha - p1.deriv[0] / p1.deriv[1], &p1next);
  while (c->LSiter < maxls) {
    ls_eval(c, 0.5 * (p1.alpha + p2.alpha), &pmid);
    ls_pnt cand[3] = {p1next, p2next, pmid};
    double bestcost = 0; int bestind = -1;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i].deriv[0]) < gtol && (bestind == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; bestind = i; }
    if (bestind >= 0) return cand[bestind].alpha;
    int b1 = ls_update_bracket(c, &p1, cand, &p1next);
    int b2 = ls_update_bracket(c, &p2, cand, &p2next);
    if (!b1 && !b2) return pmid.alpha;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}

/* mju_cholFactor / mju_cholSolve (dense, lower) */
static void chol_factor(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double tmp = A[j * n + j];
    for (int k = 0; k < j; k++) tmp -= A[j * n + k] * A[j * n + k];
    if (tmp < ORC_MINVAL) tmp = ORC_MINVAL;
    A[j * n + j] = sqrt(tmp);
    tmp = 1 / A[j * n + j];
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s * tmp;
    }
  }
}
static void chol_solve(double* x, const double* L, const double* b, int n) {
  for (int i = 0; i < n; i++) x[i] = b[i];
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < i; k++) x[i] -= L[i * n + k] * x[k];
    x[i] /= L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    for (int k = i + 1; k < n; k++) x[i] -= L[k * n + i] * x[k];
    x[i] /= L[i * n + i];
  }
}

static void newton_update_constraint(newton_ctx* c) {
  const b2m_view* v = &c->m->v;
  orc_data* d = c->d;
  c->cost = orc_constraint_update(c->m, d, c->Jaref, 1);
  double g = 0;
  for (int k = 0; k < v->nv; k++) g += (c->Ma[k] - d->qfrc_smooth[k]) * (d->qacc[k] - d->qacc_smooth[k]);
  c->gauss = 0.5 * g;
  c->cost += c->gauss;
}
static void newton_update_gradient(newton_ctx* c) {
  const b2m_view* v = &c->m->v;
  orc_data* d = c->d;
  int nv = v->nv;
  for (int k = 0; k < nv; k++) c->grad[k] = c->Ma[k] - d->qfrc_smooth[k] - d->qfrc_constraint[k];
  /* H = M + J' diag(D_active) J, Cholesky, Mgrad = H^-1 grad */
  memcpy(c->H, d->qM, sizeof(double) * nv * nv);
  for (int i = 0; i < d->nefc; i++) {
    if (c->Jaref[i] >= 0) continue;
    const double* J = d->efc_J + (size_t)i * nv;
    for (int r = 0; r < nv; r++) {
      if (J[r] == 0) continue;
      double s = d->efc_D[i] * J[r];
      for (int q = 0; q <= r; q++) c->H[r * nv + q] += s * J[q];
    }
  }
  chol_factor(c->H, nv);
  chol_solve(c->Mgrad, c->H, c->grad, nv);
}

/* mj_solNewton (mj_solPrimal with flg_Newton) */
static void orc_sol_newton(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv, nefc = d->nefc;
  newton_ctx c;
  memset(&c, 0, sizeof(c));
  c.m = m; c.d = d;
  c.Jaref = SCRATCH(nefc); c.Jv = SCRATCH(nefc); c.Ma = SCRATCH(nv); c.Mv = SCRATCH(nv); c.grad = SCRATCH(nv);
  c.Mgrad = SCRATCH(nv); c.search = SCRATCH(nv); c.quad = SCRATCH(3 * nefc); c.H = SCRATCH(nv * nv);
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qacc[k];
    c.Jaref[i] = s - d->efc_aref[i];
  }
  orc_mul_m(v, d, c.Ma, d->qacc);
  newton_update_constraint(&c);
  newton_update_gradient(&c);
  for (int k = 0; k < nv; k++) c.search[k] = -c.Mgrad[k];
  double scale = 1 / (v->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  while (iter < v->iterations) {
    double alpha = ls_search(&c);
    if (alpha == 0) break;
    for (int k = 0; k < nv; k++) { d->qacc[k] += alpha * c.search[k]; c.Ma[k] += alpha * c.Mv[k]; }
    for (int i = 0; i < nefc; i++) c.Jaref[i] += alpha * c.Jv[i];
    double oldcost = c.cost;
    newton_update_constraint(&c);
    newton_update_gradient(&c);
    double improvement = scale * (oldcost - c.cost), gn = 0;
    for (int k = 0; k < nv; k++) gn += c.grad[k] * c.grad[k];
    double gradient = scale * sqrt(gn);
    iter++;
    if (improvement < v->tolerance || gradient < v->tolerance) break;
    for (int k = 0; k < nv; k++) c.search[k] = -c.Mgrad[k];
  }
  d->solver_iter = iter;
}

/* mj_fwdConstraint */
static void orc_fwd_constraint(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv, nefc = d->nefc;
  d->solver_iter = 0;
  if (!nefc) {
    for (int k = 0; k < nv; k++) { d->qacc[k] = d->qacc_smooth[k]; d->qacc_warmstart[k] = d->qacc_smooth[k]; d->qfrc_constraint[k] = 0; }
    return;
  }
  double* jar = SCRATCH(nefc);
  double* Ma = SCRATCH(nv);
  /* warmstart(): pick the cheaper of qacc_warmstart and qacc_smooth */
  for (int k = 0; k < nv; k++) d->qacc[k] = d->qacc_warmstart[k];
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qacc[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost_warm = orc_constraint_update(m, d, jar, 0);
  orc_mul_m(v, d, Ma, d->qacc);
  for (int k = 0; k < nv; k++) cost_warm += 0.5 * (Ma[k] - d->qfrc_smooth[k]) * (d->qacc[k] - d->qacc_smooth[k]);
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qacc_smooth[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost_smooth = orc_constraint_update(m, d, jar, 0);
  if (cost_warm > cost_smooth) for (int k = 0; k < nv; k++) d->qacc[k] = d->qacc_smooth[k];
  orc_sol_newton(m, d);
  for (int k = 0; k < nv; k++) d->qacc_warmstart[k] = d->qacc[k];
}

/* ------------------------------------------------------------------ forward / step */
static void orc_fwd_position(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  orc_kinematics(v, d);
  orc_compos(v, d);
  orc_tendon(v, d);
  orc_crb(v, d);
  orc_factor(v, d->qM, d->qLD, d->qLDiagInv);
  orc_collision(m, d);
  orc_make_constraint(m, d);
  orc_transmission(m, d);
}
static void orc_fwd_velocity(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  for (int i = 0; i < v->nu; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->actuator_moment[i * nv + k] * d->qvel[k];
    d->actuator_velocity[i] = s;
  }
  for (int t = 0; t < v->ntendon; t++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->ten_J[t * nv + k] * d->qvel[k];
    d->ten_velocity[t] = s;
  }
  orc_comvel(v, d);
  orc_passive(m, d);
  orc_make_impedance(m, d); /* mj_makeImpedance (position) + mj_referenceConstraint (velocity) */
  orc_rne(v, d);
}
static void orc_fwd_acceleration(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  for (int k = 0; k < v->nv; k++) {
    d->qfrc_smooth[k] = d->qfrc_passive[k] - d->qfrc_bias[k];
    d->qfrc_smooth[k] += d->qfrc_actuator[k];
    d->qacc_smooth[k] = d->qfrc_smooth[k];
  }
  orc_solve_ld(v, d->qLD, d->qLDiagInv, d->qacc_smooth);
}
/* ------------------------------------------------------------------ sensors
 * mj_sensorPos / mj_sensorVel / mj_sensorAcc (engine_sensor.c) for the subset compiled by mjcf.py: jointpos, jointvel,
 * framepos, framequat (no reference frame), gyro, velocimeter, accelerometer.  Type codes: mjcf.SENSOR_TYPES.
 * Feeds data.sensordata behind ObservationSpec(include_sensordata=True) (reference mujoco_template/observations.py:117-127;
 * sensors declared in reference examples/drone/x2.xml:83-87). */
enum { SENS_JOINTPOS, SENS_JOINTVEL, SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_GYRO, SENS_VELOCIMETER, SENS_ACCELEROMETER };

/* the body-frame acceleration pass of mj_rnePostConstraint: cacc = cacc_parent + cdof_dot*qvel + cdof*qacc */
static void orc_body_acc(const b2m_view* v, orc_data* d) {
  memset(d->cacc, 0, sizeof(double) * 6);
  for (int k = 0; k < 3; k++) d->cacc[3 + k] = -v->gravity[k];
  for (int i = 1; i < v->nbody; i++) {
    int bda = v->body_dofadr[i];
    double* a = d->cacc + 6 * i;
    memcpy(a, d->cacc + 6 * v->body_parentid[i], sizeof(double) * 6);
    for (int j = 0; j < v->body_dofnum[i]; j++)
      for (int k = 0; k < 6; k++)
        a[k] += d->cdof_dot[6 * (bda + j) + k] * d->qvel[bda + j] + d->cdof[6 * (bda + j) + k] * d->qacc[bda + j];
  }
}
/* mj_objectVelocity / mj_objectAcceleration for a site, local frame */
static void site_velocity(const b2m_view* v, const orc_data* d, int site, double* res) {
  int b = v->site_bodyid[site];
  sp_transform(res, d->cvel + 6 * b, 0, d->site_xpos + 3 * site, d->subtree_com + 3 * v->body_rootid[b], d->site_xmat + 9 * site);
}
static void site_acceleration(const b2m_view* v, const orc_data* d, int site, double* res) {
  int b = v->site_bodyid[site];
  double vel[6], corr[3];
  site_velocity(v, d, site, vel);
  sp_transform(res, d->cacc + 6 * b, 0, d->site_xpos + 3 * site, d->subtree_com + 3 * v->body_rootid[b], d->site_xmat + 9 * site);
  v3_cross(corr, vel, vel + 3);
  v3_addto(res + 3, corr);
}
static void orc_sensors(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int need_acc = 0;
  for (int s = 0; s < v->nsensor; s++) need_acc |= v->sensor_type[s] == SENS_ACCELEROMETER;
  if (need_acc) orc_body_acc(v, d);
  for (int s = 0; s < v->nsensor; s++) {
    double* out = d->sensordata + v->sensor_adr[s];
    int id = v->sensor_objid[s], ot = v->sensor_objtype[s], dim = v->sensor_dim[s];
    double tmp[6];
    switch (v->sensor_type[s]) {
      case SENS_JOINTPOS: out[0] = d->qpos[v->jnt_qposadr[id]]; break;
      case SENS_JOINTVEL: out[0] = d->qvel[v->jnt_dofadr[id]]; break;
      case SENS_FRAMEPOS:
        v3_copy(out, ot == 1 ? d->xipos + 3 * id : ot == 2 ? d->xpos + 3 * id : ot == 5 ? d->geom_xpos + 3 * id : d->site_xpos + 3 * id);
        break;
      case SENS_FRAMEQUAT:
        if (ot == 2) memcpy(out, d->xquat + 4 * id, sizeof(double) * 4);
        else if (ot == 1) q_mul(out, d->xquat + 4 * id, v->body_iquat + 4 * id);
        else if (ot == 5) q_mul(out, d->xquat + 4 * v->geom_bodyid[id], v->geom_quat + 4 * id);
        else q_mul(out, d->xquat + 4 * v->site_bodyid[id], v->site_quat + 4 * id);
        q_normalize(out);
        break;
      case SENS_GYRO: site_velocity(v, d, id, tmp); v3_copy(out, tmp); break;
      case SENS_VELOCIMETER: site_velocity(v, d, id, tmp); v3_copy(out, tmp + 3); break;
      case SENS_ACCELEROMETER: site_acceleration(v, d, id, tmp); v3_copy(out, tmp + 3); break;
      default: break;
    }
    /* cutoff applies to real-valued outputs only (quaternions are left alone) */
    double cut = v->sensor_cutoff[s];
    if (cut > 0 && v->sensor_type[s] != SENS_FRAMEQUAT)
      for (int k = 0; k < dim; k++) out[k] = orc_clip(out[k], -cut, cut);
  }
}

/* mj_forwardSkip(skipstage = none, skipsensor) */
static void orc_forward_skip(const orc_model* m, orc_data* d, int skipsensor) {
  ARENA_MARK;
  orc_fwd_position(m, d);
  orc_fwd_velocity(m, d);
  orc_actuation(&m->v, d);
  orc_fwd_acceleration(m, d);
  orc_fwd_constraint(m, d);
  if (m->v.nsensor && !skipsensor) orc_sensors(m, d);
  ARENA_RELEASE;
}
/* mj_forward */
void orc_forward(const orc_model* m, orc_data* d) { orc_forward_skip(m, d, 0); }

/* mj_inverse (continuous-time): qfrc_inverse = M*qacc + qfrc_bias - qfrc_passive - qfrc_constraint, with the
 * constraint forces of mj_invConstraint (the given qacc fixes every row's residual: f = -D * min(0, J qacc - aref)).
 * Used by steady_ctrl0 (reference mujoco_template/setpoints.py:23-30).  Also refreshes actuator_moment. */
void orc_inverse(const orc_model* m, orc_data* d, const double* qacc, double* qfrc_inverse) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  ARENA_MARK;
  orc_fwd_position(m, d);
  orc_fwd_velocity(m, d);
  double* jar = SCRATCH(d->nefc);
  double* Ma = SCRATCH(nv);
  for (int i = 0; i < d->nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * qacc[k];
    jar[i] = s - d->efc_aref[i];
  }
  orc_constraint_update(m, d, jar, 1);
  orc_mul_m(v, d, Ma, qacc);
  for (int k = 0; k < nv; k++) qfrc_inverse[k] = Ma[k] + d->qfrc_bias[k] - d->qfrc_passive[k] - d->qfrc_constraint[k];
  ARENA_RELEASE;
}

/* mj_integratePos */
void orc_integrate_pos(const orc_model* m, double* qpos, const double* qvel, double dt) {
  const b2m_view* v = &m->v;
  for (int j = 0; j < v->njnt; j++) {
    int pa = v->jnt_qposadr[j], va = v->jnt_dofadr[j];
    if (v->jnt_type[j] == JNT_FREE) {
      for (int i = 0; i < 3; i++) qpos[pa + i] += dt * qvel[va + i];
      q_integrate(qpos + pa + 3, qvel + va + 3, dt);
    } else qpos[pa] += dt * qvel[va];
  }
}
/* mj_differentiatePos: qvel = (qpos2 - qpos1) / dt in the tangent space */
void orc_differentiate_pos(const orc_model* m, double* qvel, double dt, const double* qpos1, const double* qpos2) {
  const b2m_view* v = &m->v;
  for (int j = 0; j < v->njnt; j++) {
    int pa = v->jnt_qposadr[j], va = v->jnt_dofadr[j];
    if (v->jnt_type[j] == JNT_FREE) {
      double neg[4], dif[4];
      for (int i = 0; i < 3; i++) qvel[va + i] = (qpos2[pa + i] - qpos1[pa + i]) / dt;
      q_neg(neg, qpos1 + pa + 3);
      q_mul(dif, neg, qpos2 + pa + 3);
      q_tovel(qvel + va + 3, dif, dt);
    } else qvel[va] = (qpos2[pa] - qpos1[pa]) / dt;
  }
}

static void orc_check(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  for (int i = 0; i < v->nq; i++) if (!(fabs(d->qpos[i]) <= 1e10)) d->warn_bad_qpos = 1;
  for (int i = 0; i < v->nv; i++) if (!(fabs(d->qvel[i]) <= 1e10)) d->warn_bad_qvel = 1;
}

/* mj_Euler: semi-implicit with implicit joint damping */
static void orc_euler(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  double h = v->timestep;
  double* qacc = SCRATCH(nv);
  if (!v->has_dofdamping) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    double* MhB = SCRATCH(nv * nv);
    memcpy(MhB, d->qM, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) MhB[i * nv + i] += h * v->dof_damping[i];
    orc_factor(v, MhB, d->qH, d->qHDiagInv);
    for (int k = 0; k < nv; k++) qacc[k] = d->qfrc_smooth[k] + d->qfrc_constraint[k];
    orc_solve_ld(v, d->qH, d->qHDiagInv, qacc);
  }
  for (int k = 0; k < nv; k++) d->qvel[k] += qacc[k] * h;
  orc_integrate_pos(m, d->qpos, d->qvel, h);
  d->time += h;
}

/* mj_RungeKutta(N=4) */
static void orc_rk4(const orc_model* m, orc_data* d) {
  const b2m_view* v = &m->v;
  int nq = v->nq, nv = v->nv;
  double h = v->timestep, time = d->time;
  static const double A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1}, B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  double C[3], T[3];
  for (int i = 1; i < 4; i++) { C[i - 1] = 0; for (int j = 0; j < i; j++) C[i - 1] += A[(i - 1) * 3 + j]; T[i - 1] = time + C[i - 1] * h; }
  double *X[4], *F[4], *dX = SCRATCH(2 * nv);
  for (int i = 0; i < 4; i++) { X[i] = SCRATCH(nq + nv); F[i] = SCRATCH(nv); }
  memcpy(X[0], d->qpos, sizeof(double) * nq); memcpy(X[0] + nq, d->qvel, sizeof(double) * nv);
  memcpy(F[0], d->qacc, sizeof(double) * nv);
  for (int i = 1; i < 4; i++) {
    memset(dX, 0, sizeof(double) * 2 * nv);
    for (int j = 0; j < i; j++) {
      double a = A[(i - 1) * 3 + j];
      for (int k = 0; k < nv; k++) { dX[k] += a * X[j][nq + k]; dX[nv + k] += a * F[j][k]; }
    }
    memcpy(X[i], X[0], sizeof(double) * (nq + nv));
    orc_integrate_pos(m, X[i], dX, h);
    for (int k = 0; k < nv; k++) X[i][nq + k] += h * dX[nv + k];
    memcpy(d->qpos, X[i], sizeof(double) * nq); memcpy(d->qvel, X[i] + nq, sizeof(double) * nv);
    d->time = T[i - 1];
    orc_forward_skip(m, d, 1); /* RK4 sub-stages do not refresh sensordata */
    memcpy(F[i], d->qacc, sizeof(double) * nv);
  }
  memset(dX, 0, sizeof(double) * 2 * nv);
  for (int j = 0; j < 4; j++)
    for (int k = 0; k < nv; k++) { dX[k] += B[j] * X[j][nq + k]; dX[nv + k] += B[j] * F[j][k]; }
  d->time = time;
  memcpy(d->qpos, X[0], sizeof(double) * nq); memcpy(d->qvel, X[0] + nq, sizeof(double) * nv);
  for (int k = 0; k < nv; k++) d->qvel[k] += dX[nv + k] * h;
  orc_integrate_pos(m, d->qpos, dX, h);
  d->time += h;
}

/* mj_step */
void orc_step(const orc_model* m, orc_data* d) {
  ARENA_MARK;
  orc_check(m, d);
  orc_forward(m, d);
  for (int i = 0; i < m->v.nv; i++) if (!(fabs(d->qacc[i]) <= 1e10)) d->warn_bad_qacc = 1;
  if (m->v.integrator == 1) orc_rk4(m, d); else orc_euler(m, d);
  ARENA_RELEASE;
}

/* ------------------------------------------------------------------ mjd_transitionFD */
typedef struct { double time, *qpos, *qvel, *ctrl, *warm; } orc_state;
static void state_save(const b2m_view* v, orc_data* d, orc_state* s) {
  s->time = d->time;
  s->qpos = SCRATCH(v->nq); s->qvel = SCRATCH(v->nv); s->ctrl = SCRATCH(v->nu); s->warm = SCRATCH(v->nv);
  memcpy(s->qpos, d->qpos, sizeof(double) * v->nq); memcpy(s->qvel, d->qvel, sizeof(double) * v->nv);
  memcpy(s->ctrl, d->ctrl, sizeof(double) * v->nu); memcpy(s->warm, d->qacc_warmstart, sizeof(double) * v->nv);
}
static void state_restore(const b2m_view* v, orc_data* d, const orc_state* s) {
  d->time = s->time;
  memcpy(d->qpos, s->qpos, sizeof(double) * v->nq); memcpy(d->qvel, s->qvel, sizeof(double) * v->nv);
  memcpy(d->ctrl, s->ctrl, sizeof(double) * v->nu); memcpy(d->qacc_warmstart, s->warm, sizeof(double) * v->nv);
}
static void get_next(const b2m_view* v, const orc_data* d, double* y) {
  memcpy(y, d->qpos, sizeof(double) * v->nq); memcpy(y + v->nq, d->qvel, sizeof(double) * v->nv);
}
/* stateDiff: ds = (s2 - s1) / h with the position block in the tangent space */
static void state_diff(const orc_model* m, double* ds, const double* s1, const double* s2, double h) {
  const b2m_view* v = &m->v;
  orc_differentiate_pos(m, ds, h, s1, s2);
  for (int k = 0; k < v->nv; k++) ds[v->nv + k] = (s2[v->nq + k] - s1[v->nq + k]) * (1.0 / h);
}
static int in_range(double x1, double x2, const double* r) { return x1 >= r[0] && x1 <= r[1] && x2 >= r[0] && x2 <= r[1]; }

/* A: (2nv x 2nv) row-major, B: (2nv x nu) row-major; state restored on return */
void orc_transition_fd(const orc_model* m, orc_data* d, double eps, int centered, double* A, double* B) {
  const b2m_view* v = &m->v;
  int nq = v->nq, nv = v->nv, nu = v->nu, ndx = 2 * nv;
  ARENA_MARK;
  orc_state s;
  state_save(v, d, &s);
  double *next = SCRATCH(nq + nv), *plus = SCRATCH(nq + nv), *minus = SCRATCH(nq + nv), *col = SCRATCH(ndx), *dpos = SCRATCH(nv);
  double* sens = SCRATCH(v->nsensordata); /* FD rollouts run with skipsensor: sensordata is left as it was */
  memcpy(sens, d->sensordata, sizeof(double) * v->nsensordata);
  orc_step(m, d);
  get_next(v, d, next);
  state_restore(v, d, &s);
  for (int i = 0; i < nu; i++) {
    int limited = v->actuator_ctrllimited[i];
    const double* r = v->actuator_ctrlrange + 2 * i;
    int fwd = !limited || in_range(d->ctrl[i], d->ctrl[i] + eps, r);
    if (fwd) { d->ctrl[i] += eps; orc_step(m, d); get_next(v, d, plus); state_restore(v, d, &s); }
    int back = (centered || !fwd) && (!limited || in_range(d->ctrl[i] - eps, d->ctrl[i], r));
    if (back) { d->ctrl[i] -= eps; orc_step(m, d); get_next(v, d, minus); state_restore(v, d, &s); }
    if (fwd && !back) state_diff(m, col, next, plus, eps);
    else if (!fwd && back) state_diff(m, col, minus, next, eps);
    else if (fwd && back) state_diff(m, col, minus, plus, 2 * eps);
    else memset(col, 0, sizeof(double) * ndx);
    if (B) for (int r2 = 0; r2 < ndx; r2++) B[r2 * nu + i] = col[r2];
  }
  for (int i = 0; i < nv; i++) {
    d->qvel[i] += eps; orc_step(m, d); get_next(v, d, plus); state_restore(v, d, &s);
    if (centered) { d->qvel[i] -= eps; orc_step(m, d); get_next(v, d, minus); state_restore(v, d, &s); }
    state_diff(m, col, centered ? minus : next, plus, centered ? 2 * eps : eps);
    if (A) for (int r2 = 0; r2 < ndx; r2++) A[r2 * ndx + nv + i] = col[r2];
  }
  for (int i = 0; i < nv; i++) {
    memset(dpos, 0, sizeof(double) * nv); dpos[i] = 1;
    orc_integrate_pos(m, d->qpos, dpos, eps); orc_step(m, d); get_next(v, d, plus); state_restore(v, d, &s);
    if (centered) { orc_integrate_pos(m, d->qpos, dpos, -eps); orc_step(m, d); get_next(v, d, minus); state_restore(v, d, &s); }
    state_diff(m, col, centered ? minus : next, plus, centered ? 2 * eps : eps);
    if (A) for (int r2 = 0; r2 < ndx; r2++) A[r2 * ndx + i] = col[r2];
  }
  memcpy(d->sensordata, sens, sizeof(double) * v->nsensordata);
  ARENA_RELEASE;
}

/* ------------------------------------------------------------------ ctypes accessors */
double* orc_ptr(orc_data* d, const char* name) {
#define F(n) if (!strcmp(name, #n)) return d->n;
  F(qpos) F(qvel) F(ctrl) F(qacc) F(qacc_warmstart) F(xpos) F(xquat) F(xmat) F(xipos) F(ximat) F(geom_xpos) F(geom_xmat)
  F(site_xpos) F(site_xmat) F(subtree_com) F(cdof) F(qM) F(qfrc_bias) F(qfrc_passive) F(qfrc_actuator) F(qfrc_smooth)
  F(qacc_smooth) F(qfrc_constraint) F(efc_J) F(efc_pos) F(efc_D) F(efc_R) F(efc_aref) F(efc_force) F(cvel) F(cinert)
  F(actuator_moment) F(actuator_force) F(ten_length) F(sensordata) F(cacc)
#undef F
  return NULL;
}
double orc_get_time(const orc_data* d) { return d->time; }
void orc_set_time(orc_data* d, double t) { d->time = t; }
int orc_ncon(const orc_data* d) { return d->ncon; }
int orc_nefc(const orc_data* d) { return d->nefc; }
int orc_solver_iter(const orc_data* d) { return d->solver_iter; }
int orc_warnings(const orc_data* d) { return d->warn_bad_qpos | (d->warn_bad_qvel << 1) | (d->warn_bad_qacc << 2) | (d->warn_overflow << 3); }
/* contact k -> out[0]=dist, out[1..3]=pos, out[4..12]=frame, out[13]=dim, out[14]=geom1, out[15]=geom2 */
void orc_get_contact(const orc_data* d, int k, double* out) {
  const orc_contact* c = d->contact + k;
  out[0] = c->dist; memcpy(out + 1, c->pos, 3 * sizeof(double)); memcpy(out + 4, c->frame, 9 * sizeof(double));
  out[13] = c->dim; out[14] = c->geom1; out[15] = c->geom2;
}

/* mj_setConst cross-check: M(qpos0)^-1 based weights recomputed from the oracle's own CRB
 * out_dof[nv], out_body[2*nbody], out_tendon[ntendon], returns meaninertia */
double orc_setconst_check(const orc_model* m, double* out_dof, double* out_body, double* out_tendon) {
  const b2m_view* v = &m->v;
  int nv = v->nv;
  orc_data* d = orc_data_create(m);
  orc_reset(m, d, -1);
  orc_kinematics(v, d); orc_compos(v, d); orc_tendon(v, d); orc_crb(v, d);
  orc_factor(v, d->qM, d->qLD, d->qLDiagInv);
  double mean = 0;
  for (int i = 0; i < nv; i++) mean += d->qM[i * nv + i];
  mean /= (nv > 0 ? nv : 1);
  double* Minv = ALLOC(nv * nv);
  double* col = ALLOC(nv);
  for (int c = 0; c < nv; c++) {
    for (int k = 0; k < nv; k++) col[k] = (k == c);
    orc_solve_ld(v, d->qLD, d->qLDiagInv, col);
    for (int k = 0; k < nv; k++) Minv[k * nv + c] = col[k];
  }
  for (int j = 0; j < v->njnt; j++) {
    int da = v->jnt_dofadr[j];
    if (v->jnt_type[j] == JNT_FREE) {
      double t = 0, r = 0;
      for (int k = 0; k < 3; k++) { t += Minv[(da + k) * nv + da + k]; r += Minv[(da + 3 + k) * nv + da + 3 + k]; }
      for (int k = 0; k < 3; k++) { out_dof[da + k] = t / 3; out_dof[da + 3 + k] = r / 3; }
    } else out_dof[da] = Minv[da * nv + da];
  }
  double* J = ALLOC(6 * nv);
  for (int b = 0; b < v->nbody; b++) {
    out_body[2 * b] = out_body[2 * b + 1] = 0;
    if (!v->body_weldid[b]) continue;
    orc_jac_bodycom(m, d, J, J + 3 * nv, b);
    double diag[6];
    for (int r = 0; r < 6; r++) {
      double s = 0;
      for (int a = 0; a < nv; a++) for (int c = 0; c < nv; c++) s += J[r * nv + a] * Minv[a * nv + c] * J[r * nv + c];
      diag[r] = s;
    }
    out_body[2 * b] = orc_max(ORC_MINVAL, (diag[0] + diag[1] + diag[2]) / 3);
    out_body[2 * b + 1] = orc_max(ORC_MINVAL, (diag[3] + diag[4] + diag[5]) / 3);
  }
  for (int t = 0; t < v->ntendon; t++) {
    double s = 0;
    for (int a = 0; a < nv; a++) for (int c = 0; c < nv; c++) s += d->ten_J[t * nv + a] * Minv[a * nv + c] * d->ten_J[t * nv + c];
    out_tendon[t] = s;
  }
  free(Minv); free(col); free(J);
  orc_data_free(d);
  return mean;
}

/* ------------------------------------------------------------------ batched CPU baseline
 * N independent envs, AoS per env: qpos[N][nq], qvel[N][nv], ctrl[N][nu], warm[N][nv].
 * Each of `nsteps` steps optionally computes the FD linearization first (lin != 0), as the
 * reference Env.step does for needs_linearization controllers (env.py:178-190).
 * A/B (may be NULL) receive the last linearization: A[N][2nv*2nv], B[N][2nv*nu]. */
typedef struct {
  const orc_model* m; int e0, e1, nsteps, lin; double eps;
  double *qpos, *qvel, *ctrl, *warm, *A, *B;
} batch_job;
static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const b2m_view* v = &j->m->v;
  int nq = v->nq, nv = v->nv, nu = v->nu, nx = 2 * nv;
  orc_data* d = orc_data_create(j->m);
  double* A = ALLOC(nx * nx);
  double* B = ALLOC(nx * nu);
  for (int e = j->e0; e < j->e1; e++) {
    orc_reset(j->m, d, -1);
    memcpy(d->qpos, j->qpos + (size_t)e * nq, sizeof(double) * nq);
    memcpy(d->qvel, j->qvel + (size_t)e * nv, sizeof(double) * nv);
    memcpy(d->ctrl, j->ctrl + (size_t)e * nu, sizeof(double) * nu);
    memcpy(d->qacc_warmstart, j->warm + (size_t)e * nv, sizeof(double) * nv);
    for (int s = 0; s < j->nsteps; s++) {
      if (j->lin) orc_transition_fd(j->m, d, j->eps, 1, A, B);
      orc_step(j->m, d);
    }
    memcpy(j->qpos + (size_t)e * nq, d->qpos, sizeof(double) * nq);
    memcpy(j->qvel + (size_t)e * nv, d->qvel, sizeof(double) * nv);
    memcpy(j->warm + (size_t)e * nv, d->qacc_warmstart, sizeof(double) * nv);
    if (j->lin && j->A) memcpy(j->A + (size_t)e * nx * nx, A, sizeof(double) * nx * nx);
    if (j->lin && j->B) memcpy(j->B + (size_t)e * nx * nu, B, sizeof(double) * nx * nu);
  }
  free(A); free(B);
  orc_data_free(d);
  return NULL;
}
void orc_batch_rollout(const orc_model* m, int N, double* qpos, double* qvel, double* ctrl, double* warm, int nsteps,
                       int lin, double eps, double* A, double* B, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > N) nthreads = N > 0 ? N : 1;
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  batch_job* jobs = (batch_job*)calloc((size_t)nthreads, sizeof(batch_job));
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (batch_job){m, (int)((long long)N * t / nthreads), (int)((long long)N * (t + 1) / nthreads), nsteps, lin, eps,
                          qpos, qvel, ctrl, warm, A, B};
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

#ifdef ORC_COUNT_OPS
void orc_count_reset(void) { memset(&g_orc_ops, 0, sizeof(g_orc_ops)); }
void orc_count_read(unsigned long long* out) {
  out[0] = g_orc_ops.add; out[1] = g_orc_ops.mul; out[2] = g_orc_ops.div; out[3] = g_orc_ops.sqrt_; out[4] = g_orc_ops.trans;
}
#endif
#ifdef __cplusplus
}
#endif
