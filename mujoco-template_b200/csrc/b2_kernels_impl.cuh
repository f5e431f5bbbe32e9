// Kernels of the lane engine for ONE arithmetic type.  Included by b2_kernels_f64.cu and
// b2_kernels_f32.cu with B2_REAL / B2_SUFFIX defined; each translation unit owns its own
// __constant__ model images (one per size class), so the two precisions do not share the
// 64 KB constant bank.
//
// SoA layout: element (k, e) of a (dim, nenv) array is base[k * nenv + e]; a warp reads 32
// consecutive envs of one component -> fully coalesced 256 B (FP64) transactions.
#include <cuda_runtime.h>

#include "../../include/b2mj.h"
#include "b2_engine.cuh"
#include "b2_kernels.h"

namespace b2 {

typedef B2_REAL real;

__constant__ DevModel<real, DimsTiny> c_model_tiny;
__constant__ DevModel<real, DimsSmall> c_model_small;
__constant__ DevModel<real, DimsLarge> c_model_large;

template <class D> struct ConstModel;
template <> struct ConstModel<DimsTiny> { static B2_DEV const DevModel<real, DimsTiny>& get() { return c_model_tiny; } };
template <> struct ConstModel<DimsSmall> { static B2_DEV const DevModel<real, DimsSmall>& get() { return c_model_small; } };
template <> struct ConstModel<DimsLarge> { static B2_DEV const DevModel<real, DimsLarge>& get() { return c_model_large; } };

struct StateDev { real *qpos, *qvel, *ctrl, *warm; int* flags; };
struct DerivedDev { real *xpos, *xquat, *xipos, *geom_xpos, *site_xpos, *subtree_com, *qacc, *qfrc_bias; int *ncon, *nefc, *solver_iter; };

template <class D>
__device__ __forceinline__ void load_state(LaneEnv<real, D>& env, const StateDev& st, int N, int e) {
  const auto& m = env.m;
  for (int k = 0; k < m.nq; k++) env.qpos[k] = st.qpos[(size_t)k * N + e];
  for (int k = 0; k < m.nv; k++) env.qvel[k] = st.qvel[(size_t)k * N + e];
  for (int k = 0; k < m.nu; k++) env.ctrl[k] = st.ctrl[(size_t)k * N + e];
  for (int k = 0; k < m.nv; k++) env.warm[k] = st.warm ? st.warm[(size_t)k * N + e] : real(0);
}
template <class D>
__device__ __forceinline__ void store_derived(const LaneEnv<real, D>& env, const DerivedDev& o, int N, int e) {
  const auto& m = env.m;
  if (o.xpos) for (int k = 0; k < 3 * m.nbody; k++) o.xpos[(size_t)k * N + e] = env.xpos[k];
  if (o.xquat) for (int k = 0; k < 4 * m.nbody; k++) o.xquat[(size_t)k * N + e] = env.xquat[k];
  if (o.xipos) for (int k = 0; k < 3 * m.nbody; k++) o.xipos[(size_t)k * N + e] = env.xipos[k];
  if (o.geom_xpos) for (int k = 0; k < 3 * m.ngeom; k++) o.geom_xpos[(size_t)k * N + e] = env.geom_xpos[k];
  if (o.site_xpos) for (int k = 0; k < 3 * m.nsite; k++) o.site_xpos[(size_t)k * N + e] = env.site_xpos[k];
  if (o.subtree_com) for (int k = 0; k < 3 * m.nbody; k++) o.subtree_com[(size_t)k * N + e] = env.com[k];
  if (o.qacc) for (int k = 0; k < m.nv; k++) o.qacc[(size_t)k * N + e] = env.qacc[k];
  if (o.qfrc_bias) for (int k = 0; k < m.nv; k++) o.qfrc_bias[(size_t)k * N + e] = env.f_bias[k];
  if (o.ncon) o.ncon[e] = env.ncon;
  if (o.nefc) o.nefc[e] = env.nefc;
  if (o.solver_iter) o.solver_iter[e] = env.niter;
}

// nsteps x mj_step with ctrl held; nsteps == 0 means mj_forward (no integration)
template <class D>
__global__ void __launch_bounds__(128) k_step(StateDev st, DerivedDev out, int want_derived, int N, int nsteps) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  LaneEnv<real, D> env(ConstModel<D>::get());
  load_state(env, st, N, e);
  const auto& m = env.m;
  if (nsteps == 0) {
    env.check_state();
    env.forward();
    if (want_derived) store_derived(env, out, N, e);
  } else {
    for (int s = 0; s < nsteps; s++) {
      if (s == nsteps - 1 && want_derived) env.step([&] { store_derived(env, out, N, e); });
      else env.step();
    }
    for (int k = 0; k < m.nq; k++) st.qpos[(size_t)k * N + e] = env.qpos[k];
    for (int k = 0; k < m.nv; k++) st.qvel[(size_t)k * N + e] = env.qvel[k];
  }
  if (st.warm) for (int k = 0; k < m.nv; k++) st.warm[(size_t)k * N + e] = env.warm[k];
  if (st.flags && env.flags) st.flags[e] |= env.flags;
}

// Centred / one-sided finite differences of one step: one thread per (env, input column).
// Columns 0..nv-1 perturb tangent-space position, nv..2nv-1 velocity, 2nv..2nv+nu-1 control.
// Replaces mjd_transitionFD (reference mujoco_template/linearization.py:16-35).
template <class D>
__global__ void __launch_bounds__(128) k_linearize(StateDev st, int N, real eps, int centered, real* A, real* B) {
  const auto& m = ConstModel<D>::get();
  const int nq = m.nq, nv = m.nv, nu = m.nu, ndx = 2 * nv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * (ndx + nu)) return;
  const int e = (int)(idx % N), c = (int)(idx / N);
  LaneEnv<real, D> env(m);
  real q0[D::NQ], v0[D::NV], u0[D::NU], w0[D::NV], yp[D::NQ + D::NV], ym[D::NQ + D::NV], col[2 * D::NV];
  for (int k = 0; k < nq; k++) q0[k] = st.qpos[(size_t)k * N + e];
  for (int k = 0; k < nv; k++) { v0[k] = st.qvel[(size_t)k * N + e]; w0[k] = st.warm ? st.warm[(size_t)k * N + e] : real(0); }
  for (int k = 0; k < nu; k++) u0[k] = st.ctrl[(size_t)k * N + e];
  // one perturbed rollout: kind 0 nominal, 1 position, 2 velocity, 3 control
  auto rollout = [&](int kind, int i, real delta, real* y) {
    for (int k = 0; k < nq; k++) env.qpos[k] = q0[k];
    for (int k = 0; k < nv; k++) { env.qvel[k] = v0[k]; env.warm[k] = w0[k]; }
    for (int k = 0; k < nu; k++) env.ctrl[k] = u0[k];
    if (kind == 1) {
      real dp[D::NV];
      for (int k = 0; k < nv; k++) dp[k] = 0;
      dp[i] = 1;
      env.integrate_pos(env.qpos, dp, delta);
    } else if (kind == 2) env.qvel[i] += delta;
    else if (kind == 3) env.ctrl[i] += delta;
    env.step();
    for (int k = 0; k < nq; k++) y[k] = env.qpos[k];
    for (int k = 0; k < nv; k++) y[nq + k] = env.qvel[k];
  };
  int kind, i, fwd = 1, back = centered;
  if (c < nv) { kind = 1; i = c; }
  else if (c < ndx) { kind = 2; i = c - nv; }
  else {
    kind = 3; i = c - ndx;
    const int lim = m.actuator_ctrllimited[i];
    const real lo = m.actuator_ctrlrange[2 * i], hi = m.actuator_ctrlrange[2 * i + 1], u = u0[i];
    fwd = !lim || (u >= lo && u <= hi && u + eps >= lo && u + eps <= hi);
    back = (centered || !fwd) && (!lim || (u - eps >= lo && u - eps <= hi && u >= lo && u <= hi));
  }
  real h;
  const real *s1, *s2;
  if (fwd) rollout(kind, i, eps, yp);
  if (back) rollout(kind, i, -eps, ym);
  if (fwd && back) { s1 = ym; s2 = yp; h = 2 * eps; }
  else if (fwd) { rollout(0, 0, 0, ym); s1 = ym; s2 = yp; h = eps; }
  else if (back) { rollout(0, 0, 0, yp); s1 = ym; s2 = yp; h = eps; }
  else { s1 = s2 = nullptr; h = 1; }
  if (s1) {
    env.differentiate_pos(col, h, s1, s2);
    const real ih = real(1) / h;
    for (int k = 0; k < nv; k++) col[nv + k] = (s2[nq + k] - s1[nq + k]) * ih;
  } else {
    for (int k = 0; k < ndx; k++) col[k] = 0;
  }
  if (c < ndx) { if (A) for (int r = 0; r < ndx; r++) A[((size_t)r * ndx + c) * N + e] = col[r]; }
  else if (B) for (int r = 0; r < ndx; r++) B[((size_t)r * nu + (c - ndx)) * N + e] = col[r];
  if (st.flags && env.flags) atomicOr(st.flags + e, env.flags);
}

// point Jacobians from the current qpos (mj_jacSite/Body/BodyCom/SubtreeCom)
template <class D>
__global__ void __launch_bounds__(128) k_jacobian(StateDev st, int N, int kind, int objid, real* jacp, real* jacr) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  LaneEnv<real, D> env(ConstModel<D>::get());
  const auto& m = env.m;
  const int nv = m.nv;
  for (int k = 0; k < m.nq; k++) env.qpos[k] = st.qpos[(size_t)k * N + e];
  env.kinematics();
  env.com_frame();
  real jp[3 * D::NV], jr[3 * D::NV];
  for (int k = 0; k < 3 * nv; k++) { jp[k] = 0; jr[k] = 0; }
  auto put = [&](int d, const real* p, const real* r) {
    for (int a = 0; a < 3; a++) { jp[a * nv + d] = p[a]; jr[a * nv + d] = r[a]; }
  };
  if (kind == B2_JAC_SITE) env.for_jac(m.site_bodyid[objid], env.site_xpos + 3 * objid, put);
  else if (kind == B2_JAC_BODY) env.for_jac(objid, env.xpos + 3 * objid, put);
  else if (kind == B2_JAC_BODYCOM) env.for_jac(objid, env.xipos + 3 * objid, put);
  else {
    for (int b = objid; b < m.nbody; b++) {
      if (b > objid && m.body_parentid[b] < objid) break;
      const real mass = m.body_mass[b];
      env.for_jac(b, env.xipos + 3 * b, [&](int d, const real* p, const real*) {
        for (int a = 0; a < 3; a++) jp[a * nv + d] += p[a] * mass;
      });
    }
    const real inv = real(1) / m.body_subtreemass[objid];
    for (int k = 0; k < 3 * nv; k++) jp[k] *= inv;
  }
  if (jacp) for (int k = 0; k < 3 * nv; k++) jacp[(size_t)k * N + e] = jp[k];
  if (jacr) for (int k = 0; k < 3 * nv; k++) jacr[(size_t)k * N + e] = jr[k];
}

template <class D>
__global__ void __launch_bounds__(128) k_integrate_pos(real* qpos, const real* qvel, real dt, int N) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  LaneEnv<real, D> env(ConstModel<D>::get());
  const auto& m = env.m;
  for (int k = 0; k < m.nq; k++) env.qpos[k] = qpos[(size_t)k * N + e];
  for (int k = 0; k < m.nv; k++) env.qvel[k] = qvel[(size_t)k * N + e];
  env.integrate_pos(env.qpos, env.qvel, dt);
  for (int k = 0; k < m.nq; k++) qpos[(size_t)k * N + e] = env.qpos[k];
}
template <class D>
__global__ void __launch_bounds__(128) k_differentiate_pos(real* out, real dt, const real* q1, const real* q2, int N) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  LaneEnv<real, D> env(ConstModel<D>::get());
  const auto& m = env.m;
  real a[D::NQ], b[D::NQ];
  for (int k = 0; k < m.nq; k++) { a[k] = q1[(size_t)k * N + e]; b[k] = q2[(size_t)k * N + e]; }
  env.differentiate_pos(env.qvel, dt, a, b);
  for (int k = 0; k < m.nv; k++) out[(size_t)k * N + e] = env.qvel[k];
}

// ------------------------------------------------------------------ host launchers
static inline StateDev to_dev(const b2_state* s) {
  StateDev d;
  d.qpos = (real*)s->qpos; d.qvel = (real*)s->qvel; d.ctrl = (real*)s->ctrl; d.warm = (real*)s->qacc_warmstart; d.flags = s->flags;
  return d;
}
static inline DerivedDev to_dev(const b2_derived* o) {
  DerivedDev d;
  memset(&d, 0, sizeof(d));
  if (o) {
    d.xpos = (real*)o->xpos; d.xquat = (real*)o->xquat; d.xipos = (real*)o->xipos; d.geom_xpos = (real*)o->geom_xpos;
    d.site_xpos = (real*)o->site_xpos; d.subtree_com = (real*)o->subtree_com; d.qacc = (real*)o->qacc;
    d.qfrc_bias = (real*)o->qfrc_bias; d.ncon = o->ncon; d.nefc = o->nefc; d.solver_iter = o->solver_iter;
  }
  return d;
}

#define B2_DISPATCH(cls, CALL)                         \
  switch (cls) {                                       \
    case 0: { typedef DimsTiny DD; CALL; break; }      \
    case 1: { typedef DimsSmall DD; CALL; break; }     \
    default: { typedef DimsLarge DD; CALL; break; }    \
  }

template <class D>
static cudaError_t upload_t(const b2m_view& v, const int* disabled, cudaStream_t s);
template <> cudaError_t upload_t<DimsTiny>(const b2m_view& v, const int* disabled, cudaStream_t s) {
  static DevModel<real, DimsTiny> h; fill_dev_model(h, v, disabled);
  return cudaMemcpyToSymbolAsync(c_model_tiny, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}
template <> cudaError_t upload_t<DimsSmall>(const b2m_view& v, const int* disabled, cudaStream_t s) {
  static DevModel<real, DimsSmall> h; fill_dev_model(h, v, disabled);
  return cudaMemcpyToSymbolAsync(c_model_small, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}
template <> cudaError_t upload_t<DimsLarge>(const b2m_view& v, const int* disabled, cudaStream_t s) {
  static DevModel<real, DimsLarge> h; fill_dev_model(h, v, disabled);
  return cudaMemcpyToSymbolAsync(c_model_large, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}

#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)
#define B2_FN(name) B2_CAT(name, B2_SUFFIX)

int B2_FN(b2k_upload)(int cls, const b2m_view* v, const int* disabled, void* stream) {
  cudaError_t err = cudaSuccess;
  B2_DISPATCH(cls, err = upload_t<DD>(*v, disabled, (cudaStream_t)stream));
  if (err == cudaSuccess) err = cudaStreamSynchronize((cudaStream_t)stream);  // host image is static: finish before reuse
  return (int)err;
}
int B2_FN(b2k_step)(int cls, const b2_state* st, const b2_derived* out, int N, int nsteps, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_step<DD><<<blocks, threads, 0, (cudaStream_t)stream>>>(to_dev(st), to_dev(out), out != nullptr, N, nsteps)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_linearize)(int cls, const b2_state* st, int N, int ncol, double eps, int centered, void* A, void* B, void* stream) {
  const int threads = 128;
  const long long total = (long long)N * ncol;
  const int blocks = (int)((total + threads - 1) / threads);
  B2_DISPATCH(cls, (k_linearize<DD><<<blocks, threads, 0, (cudaStream_t)stream>>>(to_dev(st), N, (real)eps, centered, (real*)A, (real*)B)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_jacobian)(int cls, const b2_state* st, int N, int kind, int objid, void* jacp, void* jacr, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_jacobian<DD><<<blocks, threads, 0, (cudaStream_t)stream>>>(to_dev(st), N, kind, objid, (real*)jacp, (real*)jacr)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_integrate_pos)(int cls, void* qpos, const void* qvel, double dt, int N, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_integrate_pos<DD><<<blocks, threads, 0, (cudaStream_t)stream>>>((real*)qpos, (const real*)qvel, (real)dt, N)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_differentiate_pos)(int cls, void* out, double dt, const void* q1, const void* q2, int N, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_differentiate_pos<DD><<<blocks, threads, 0, (cudaStream_t)stream>>>((real*)out, (real)dt, (const real*)q1, (const real*)q2, N)));
  return (int)cudaGetLastError();
}

}  // namespace b2
