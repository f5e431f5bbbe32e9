// Device-side model constants.
//
// The engine never touches a model struct directly; it reads the model through a *provider*
// policy class with one static accessor per field (`M::nbody()`, `M::body_parentid(i)`, ...):
//
//  * RuntimeModel<T, D>  (b2_kernels_impl.cuh) returns fields of a POD image in __constant__
//    memory -- uniform constant-bank operands, any model that fits size class D;
//  * generated static providers (csrc/generated/spec_<model>.cuh, emitted by
//    mujoco_template/_specialize.py) return compile-time constants from function-local
//    constexpr tables, so that after unrolling every index, branch and model constant folds
//    and the whole per-env state lives in registers.
//
// The image is filled on the host from the compiled-model blob (include/b2_model_layout.h),
// which replaces the reference's mj.MjModel (reference mujoco_template/model.py:14-25).
#pragma once
#include "../../include/b2_model_layout.h"

namespace b2 {

// Compile-time capacities of the generic size classes.  A model is mapped to the smallest
// class that holds it.
struct DimsTiny {   // pendulum, cartpole, the reference test fixture
  static constexpr int NB = 4, NJ = 4, NQ = 4, NV = 4, NU = 2, NG = 4, NS = 2, NT = 1, NW = 2, NPAIR = 4, NCON = 8, NEFC = 36, NSEN = 4, NSD = 16;
};
struct DimsSmall {  // drone: one free body with many geoms
  static constexpr int NB = 4, NJ = 4, NQ = 10, NV = 8, NU = 8, NG = 12, NS = 8, NT = 1, NW = 2, NPAIR = 12, NCON = 24, NEFC = 100, NSEN = 8, NSD = 32;
};
struct DimsLarge {  // humanoid
  static constexpr int NB = 20, NJ = 24, NQ = 32, NV = 32, NU = 24, NG = 24, NS = 4, NT = 4, NW = 8, NPAIR = 176, NCON = 48, NEFC = 160, NSEN = 16, NSD = 48;
};

// FD linearisation work split (k_linearize): threads per env
#define B2_FD_GROUP 4
__host__ __device__ inline int fd_group_count(int nv, int nu) { return (nv + nu + B2_FD_GROUP - 1) / B2_FD_GROUP; }
__host__ __device__ inline int fd_task_count(int integrator, int nv, int nu) {
  return integrator == 0 ? fd_group_count(nv, nu) + nv : 2 * nv + nu;
}

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4, GEOM_BOX = 6 };
enum { TRN_JOINT = 0, TRN_SITE = 4 };
enum { SENS_JOINTPOS = 0, SENS_JOINTVEL, SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_GYRO, SENS_VELOCIMETER, SENS_ACCELEROMETER };
enum { ROW_LIMIT_JOINT = 0, ROW_LIMIT_TENDON = 1, ROW_CONTACT_1 = 2, ROW_CONTACT_PYR = 3 };

// Field lists (X-macros): X(name) for scalars, X(name, capacity) for arrays.
#define B2_MODEL_INT_SCALARS(X) \
  X(nq) X(nv) X(nu) X(nbody) X(njnt) X(ngeom) X(nsite) X(ntendon) X(npair) X(integrator) X(iterations) X(ls_iterations) \
  X(has_fluid) X(has_dofdamping) X(maxdepth) X(nsensor) X(nsensordata)
#define B2_MODEL_REAL_SCALARS(X) X(timestep) X(density) X(viscosity) X(tolerance) X(ls_tolerance) X(meaninertia)
#define B2_MODEL_INT_ARRAYS(X)                                                                                        \
  X(body_parentid, D::NB) X(body_rootid, D::NB) X(body_jntnum, D::NB) X(body_jntadr, D::NB) X(body_dofnum, D::NB)    \
  X(body_dofadr, D::NB) X(body_anc, D::NB) X(body_depth, D::NB) X(jnt_type, D::NJ) X(jnt_qposadr, D::NJ) X(jnt_dofadr, D::NJ) X(jnt_bodyid, D::NJ)           \
  X(jnt_limited, D::NJ) X(dof_bodyid, D::NV) X(dof_jntid, D::NV) X(dof_parentid, D::NV) X(dof_anc, D::NV) X(dof_nanc, D::NV) X(dof_anclist, D::NV * D::NV)            \
  X(geom_type, D::NG) X(geom_bodyid, D::NG) X(site_bodyid, D::NS) X(tendon_adr, D::NT) X(tendon_num, D::NT)          \
  X(tendon_limited, D::NT) X(wrap_jntid, D::NW) X(actuator_trntype, D::NU) X(actuator_trnid, D::NU)                  \
  X(actuator_ctrllimited, D::NU) X(actuator_forcelimited, D::NU) X(actuator_disabled, D::NU) X(pair_geom1, D::NPAIR) \
  X(pair_geom2, D::NPAIR) X(pair_dim, D::NPAIR) X(sensor_type, D::NSEN) X(sensor_objtype, D::NSEN)                   \
  X(sensor_objid, D::NSEN) X(sensor_adr, D::NSEN) X(sensor_dim, D::NSEN)
#define B2_MODEL_REAL_ARRAYS(X)                                                                                       \
  X(gravity, 3) X(wind, 3) X(body_pos, 3 * D::NB) X(body_quat, 4 * D::NB) X(body_ipos, 3 * D::NB)                     \
  X(body_iquat, 4 * D::NB) X(body_mass, D::NB) X(body_subtreemass, D::NB) X(body_inertia, 3 * D::NB)                  \
  X(body_invweight0, 2 * D::NB) X(jnt_pos, 3 * D::NJ) X(jnt_axis, 3 * D::NJ) X(jnt_stiffness, D::NJ)                  \
  X(jnt_range, 2 * D::NJ) X(jnt_margin, D::NJ) X(jnt_solref, 2 * D::NJ) X(jnt_solimp, 5 * D::NJ) X(qpos0, D::NQ)      \
  X(qpos_spring, D::NQ) X(dof_armature, D::NV) X(dof_damping, D::NV) X(dof_invweight0, D::NV) X(geom_size, 3 * D::NG) \
  X(geom_rbound, D::NG) X(geom_pos, 3 * D::NG) X(geom_quat, 4 * D::NG) X(site_pos, 3 * D::NS) X(site_quat, 4 * D::NS) \
  X(tendon_range, 2 * D::NT) X(tendon_margin, D::NT) X(tendon_solref, 2 * D::NT) X(tendon_solimp, 5 * D::NT)          \
  X(tendon_invweight0, D::NT) X(tendon_stiffness, D::NT) X(tendon_damping, D::NT) X(tendon_lengthspring, 2 * D::NT)   \
  X(wrap_coef, D::NW) X(actuator_gear, 6 * D::NU) X(actuator_ctrlrange, 2 * D::NU) X(actuator_forcerange, 2 * D::NU)  \
  X(actuator_gainprm, D::NU) X(actuator_biasprm, 3 * D::NU) X(pair_margin, D::NPAIR) X(pair_gap, D::NPAIR)            \
  X(pair_friction, 2 * D::NPAIR) X(pair_solref, 2 * D::NPAIR) X(pair_solimp, 5 * D::NPAIR) X(sensor_cutoff, D::NSEN)

template <typename T, class D>
struct DevModel {
#define X(name) int name;
  B2_MODEL_INT_SCALARS(X)
#undef X
#define X(name) T name;
  B2_MODEL_REAL_SCALARS(X)
#undef X
#define X(name, cap) int name[cap];
  B2_MODEL_INT_ARRAYS(X)
#undef X
#define X(name, cap) T name[cap];
  B2_MODEL_REAL_ARRAYS(X)
#undef X
};

template <class D>
inline bool model_fits(const b2m_view& v) {
  return v.nbody <= D::NB && v.njnt <= D::NJ && v.nq <= D::NQ && v.nv <= D::NV && v.nu <= D::NU && v.ngeom <= D::NG &&
         v.nsite <= D::NS && v.ntendon <= D::NT && v.nwrap <= D::NW && v.npair <= D::NPAIR && v.nv <= 32 && v.nsensor <= D::NSEN &&
         v.nsensordata <= D::NSD;
}

template <typename T, typename S>
inline void fill(T* dst, const S* src, int n) {
  for (int i = 0; i < n; i++) dst[i] = (T)src[i];
}

// bit j of dof_anc[i] is set iff dof j is i itself or an ancestor of i in the kinematic tree
inline void dof_ancestor_masks(const b2m_view& v, int* out) {
  for (int i = 0; i < v.nv; i++) {
    unsigned mask = 0;
    for (int j = i; j >= 0; j = v.dof_parentid[j]) mask |= 1u << j;
    out[i] = (int)mask;
  }
}

template <typename T, class D>
inline void fill_dev_model(DevModel<T, D>& m, const b2m_view& v, const int* actuator_disabled) {
  memset(&m, 0, sizeof(m));
  m.nq = v.nq; m.nv = v.nv; m.nu = v.nu; m.nbody = v.nbody; m.njnt = v.njnt; m.ngeom = v.ngeom; m.nsite = v.nsite;
  m.ntendon = v.ntendon; m.npair = v.npair; m.integrator = v.integrator; m.iterations = v.iterations;
  m.ls_iterations = v.ls_iterations; m.has_fluid = v.has_fluid; m.has_dofdamping = v.has_dofdamping;
  m.nsensor = v.nsensor; m.nsensordata = v.nsensordata;
  fill(m.sensor_type, v.sensor_type, v.nsensor); fill(m.sensor_objtype, v.sensor_objtype, v.nsensor);
  fill(m.sensor_objid, v.sensor_objid, v.nsensor); fill(m.sensor_adr, v.sensor_adr, v.nsensor);
  fill(m.sensor_dim, v.sensor_dim, v.nsensor); fill(m.sensor_cutoff, v.sensor_cutoff, v.nsensor);
  m.timestep = (T)v.timestep; fill(m.gravity, v.gravity, 3); fill(m.wind, v.wind, 3);
  m.density = (T)v.density; m.viscosity = (T)v.viscosity; m.tolerance = (T)v.tolerance; m.ls_tolerance = (T)v.ls_tolerance;
  m.meaninertia = (T)v.meaninertia;
  int nb = v.nbody, nj = v.njnt, nv = v.nv, ng = v.ngeom, ns = v.nsite, nt = v.ntendon, nu = v.nu, np = v.npair;
  fill(m.body_parentid, v.body_parentid, nb); fill(m.body_rootid, v.body_rootid, nb); fill(m.body_jntnum, v.body_jntnum, nb);
  fill(m.body_jntadr, v.body_jntadr, nb); fill(m.body_dofnum, v.body_dofnum, nb); fill(m.body_dofadr, v.body_dofadr, nb);
  fill(m.body_pos, v.body_pos, 3 * nb); fill(m.body_quat, v.body_quat, 4 * nb); fill(m.body_ipos, v.body_ipos, 3 * nb);
  fill(m.body_iquat, v.body_iquat, 4 * nb); fill(m.body_mass, v.body_mass, nb); fill(m.body_subtreemass, v.body_subtreemass, nb);
  fill(m.body_inertia, v.body_inertia, 3 * nb); fill(m.body_invweight0, v.body_invweight0, 2 * nb);
  fill(m.jnt_type, v.jnt_type, nj); fill(m.jnt_qposadr, v.jnt_qposadr, nj); fill(m.jnt_dofadr, v.jnt_dofadr, nj);
  fill(m.jnt_bodyid, v.jnt_bodyid, nj); fill(m.jnt_limited, v.jnt_limited, nj);
  fill(m.jnt_pos, v.jnt_pos, 3 * nj); fill(m.jnt_axis, v.jnt_axis, 3 * nj); fill(m.jnt_stiffness, v.jnt_stiffness, nj);
  fill(m.jnt_range, v.jnt_range, 2 * nj); fill(m.jnt_margin, v.jnt_margin, nj); fill(m.jnt_solref, v.jnt_solref, 2 * nj);
  fill(m.jnt_solimp, v.jnt_solimp, 5 * nj); fill(m.qpos0, v.qpos0, v.nq); fill(m.qpos_spring, v.qpos_spring, v.nq);
  fill(m.dof_bodyid, v.dof_bodyid, nv); fill(m.dof_jntid, v.dof_jntid, nv); fill(m.dof_parentid, v.dof_parentid, nv);
  dof_ancestor_masks(v, m.dof_anc);
  // proper-ancestor lists (nearest first) with stride D::NV, for the warp engine's pivot updates
  for (int i = 0; i < nv; i++) {
    int n = 0;
    for (int j = v.dof_parentid[i]; j >= 0; j = v.dof_parentid[j]) m.dof_anclist[i * D::NV + n++] = j;
    m.dof_nanc[i] = n;
  }
  // body_anc[i]: bit j set iff body j is i or an ancestor of i (needs nbody <= 32); depth of world = 0
  m.maxdepth = 0;
  for (int i = 0; i < nb && i < 32; i++) {
    unsigned mask = 1u << i;
    int depth = 0;
    for (int j = i; j > 0; j = v.body_parentid[j]) { mask |= 1u << v.body_parentid[j]; depth++; }
    m.body_anc[i] = (int)mask; m.body_depth[i] = depth;
    if (depth > m.maxdepth) m.maxdepth = depth;
  }
  fill(m.dof_armature, v.dof_armature, nv); fill(m.dof_damping, v.dof_damping, nv); fill(m.dof_invweight0, v.dof_invweight0, nv);
  fill(m.geom_type, v.geom_type, ng); fill(m.geom_bodyid, v.geom_bodyid, ng); fill(m.geom_size, v.geom_size, 3 * ng);
  fill(m.geom_rbound, v.geom_rbound, ng); fill(m.geom_pos, v.geom_pos, 3 * ng); fill(m.geom_quat, v.geom_quat, 4 * ng);
  fill(m.site_bodyid, v.site_bodyid, ns); fill(m.site_pos, v.site_pos, 3 * ns); fill(m.site_quat, v.site_quat, 4 * ns);
  fill(m.tendon_adr, v.tendon_adr, nt); fill(m.tendon_num, v.tendon_num, nt); fill(m.tendon_limited, v.tendon_limited, nt);
  fill(m.wrap_jntid, v.wrap_jntid, v.nwrap); fill(m.wrap_coef, v.wrap_coef, v.nwrap);
  fill(m.tendon_range, v.tendon_range, 2 * nt); fill(m.tendon_margin, v.tendon_margin, nt); fill(m.tendon_solref, v.tendon_solref, 2 * nt);
  fill(m.tendon_solimp, v.tendon_solimp, 5 * nt); fill(m.tendon_invweight0, v.tendon_invweight0, nt);
  fill(m.tendon_stiffness, v.tendon_stiffness, nt); fill(m.tendon_damping, v.tendon_damping, nt);
  fill(m.tendon_lengthspring, v.tendon_lengthspring, 2 * nt);
  fill(m.actuator_trntype, v.actuator_trntype, nu); fill(m.actuator_trnid, v.actuator_trnid, nu);
  fill(m.actuator_ctrllimited, v.actuator_ctrllimited, nu); fill(m.actuator_forcelimited, v.actuator_forcelimited, nu);
  fill(m.actuator_disabled, actuator_disabled ? actuator_disabled : v.actuator_disabled, nu);
  fill(m.actuator_gear, v.actuator_gear, 6 * nu); fill(m.actuator_ctrlrange, v.actuator_ctrlrange, 2 * nu);
  fill(m.actuator_forcerange, v.actuator_forcerange, 2 * nu); fill(m.actuator_gainprm, v.actuator_gainprm, nu);
  fill(m.actuator_biasprm, v.actuator_biasprm, 3 * nu);
  fill(m.pair_geom1, v.pair_geom1, np); fill(m.pair_geom2, v.pair_geom2, np); fill(m.pair_dim, v.pair_dim, np);
  fill(m.pair_margin, v.pair_margin, np); fill(m.pair_gap, v.pair_gap, np); fill(m.pair_solref, v.pair_solref, 2 * np);
  fill(m.pair_solimp, v.pair_solimp, 5 * np);
  // condim <= 3 uses only the tangential friction coefficient; keep friction[0..1]
  for (int p = 0; p < np; p++) { m.pair_friction[2 * p] = (T)v.pair_friction[5 * p]; m.pair_friction[2 * p + 1] = (T)v.pair_friction[5 * p + 1]; }
}

}  // namespace b2
