// Kernels of the warp engine (one env per warp, workspace in shared memory).
#pragma once
#include "b2_kernel_templates.cuh"
#include "b2_warp_engine.cuh"

// The lock-step kernels are compiled for at most 200 registers (__maxnreg__): with 21.8 KB of shared memory per env that is
// five two-warp blocks = 10 warps per SM (254 registers and 26.5 KB: 8 warps; +3 % humanoid throughput).  A 168-register
// build (__launch_bounds__(64, 5) picks that) spills and is 5 % slower than the 8-warp one.

namespace b2 {

constexpr int kWarpIntsAsReals = (WarpCaps::NCON + WarpCaps::NEFC + 1) / 2;  // int metadata, counted in 8-byte units

template <typename T, class M>
B2_DEV void warp_store_derived(const WarpEnv<T, M>& env, const DerivedDev<T>& o, int N, int e) {
  const int lane = env.lane;
  if (o.xpos) WFOR(k, 3 * M::nbody()) o.xpos[(size_t)k * N + e] = env.xpos[k];
  if (o.xquat) WFOR(k, 4 * M::nbody()) o.xquat[(size_t)k * N + e] = env.xquat[k];
  if (o.xipos) WFOR(k, 3 * M::nbody()) o.xipos[(size_t)k * N + e] = env.xipos[k];
  if (o.geom_xpos) WFOR(k, 3 * M::ngeom()) o.geom_xpos[(size_t)k * N + e] = env.geom_xpos[k];
  if (o.subtree_com) WFOR(k, 3 * M::nbody()) o.subtree_com[(size_t)k * N + e] = env.com[k];
  if (o.qacc) WFOR(k, M::nv()) o.qacc[(size_t)k * N + e] = env.qacc[k];
  if (o.qfrc_bias) WFOR(k, M::nv()) o.qfrc_bias[(size_t)k * N + e] = env.f_bias[k];
  if (lane == 0) {
    if (o.ncon) o.ncon[e] = env.ncon;
    if (o.nefc) o.nefc[e] = env.nefc;
    if (o.solver_iter) o.solver_iter[e] = env.niter;
  }
}

// nsteps x mj_step (nsteps == 0: mj_forward) for envs gw, gw + nwarps, ...; one env per warp
template <typename T, class M>
__global__ void __launch_bounds__(64) k_warp_step(StateDev<T> st, DerivedDev<T> out, int want_derived, int N, int nsteps,
                                                  T* jscratch, int* queue, int ws_reals) {
  extern __shared__ double b2_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  T* base = reinterpret_cast<T*>(b2_smem) + (size_t)wib * (ws_reals + kWarpIntsAsReals * (int)(sizeof(double) / sizeof(T)));
  WarpEnv<T, M> env;
  env.bind(base, reinterpret_cast<int*>(base + ws_reals), jscratch + (size_t)gw * warp_slot_reals(M::nv()));
  const int total = nsteps > 0 ? nsteps : 1;
  // dynamic scheduling: env costs differ (contacts, Newton iterations), a static stride leaves SMs idle at the tail
  for (int e = gw; e < N; e = nw + __shfl_sync(0xffffffffu, lane == 0 ? atomicAdd(queue, 1) : 0, 0)) {
    WFOR(k, M::nq()) env.qpos[k] = st.qpos[(size_t)k * N + e];
    WFOR(k, M::nv()) { env.qvel[k] = st.qvel[(size_t)k * N + e]; env.warm[k] = st.warm ? st.warm[(size_t)k * N + e] : T(0); }
    WFOR(k, M::nu()) env.ctrl[k] = st.ctrl[(size_t)k * N + e];
    env.flags = 0;
    __syncwarp();
    bool frozen = false;  // diverged env: flagged and left as it is (see k_step)
    for (int s = 0; s < total && !frozen; s++) {
      env.check_state();
      if (env.flags & 3) { frozen = true; break; }
      env.forward();
      env.check_acc();
      if (want_derived && s == total - 1) warp_store_derived(env, out, N, e);
      if (nsteps > 0) env.euler();
    }
    if (frozen) {
      if (st.flags && lane == 0) st.flags[e] |= env.flags;
      __syncwarp();
      continue;
    }
    if (nsteps > 0) {
      WFOR(k, M::nq()) st.qpos[(size_t)k * N + e] = env.qpos[k];
      WFOR(k, M::nv()) st.qvel[(size_t)k * N + e] = env.qvel[k];
    }
    if (st.warm) WFOR(k, M::nv()) st.warm[(size_t)k * N + e] = env.warm[k];
    if (st.flags && env.flags && lane == 0) st.flags[e] |= env.flags;
    __syncwarp();
  }
}

// Lock-step variant: the warps of a block form groups of `gsize` warps; a group takes gsize envs at a time from the work
// queue and runs them stage by stage behind a named barrier (WarpEnv::forward<LS>).  Warps without an env of their own
// at the tail of the batch recompute the last env and discard the result, so that every warp reaches every barrier.
// map == 0: group = warp % ngroups (with 8 warps and pairs: warps w and w + 4, which share an SM sub-partition);
// map == 1: group = warp / gsize (adjacent warps).
template <typename T, class M, int LS>
__global__ void __maxnreg__(200) k_warp_step_ls(StateDev<T> st, DerivedDev<T> out, int want_derived, int N, int nsteps,
                                                        T* jscratch, int* queue, int ws_reals, int gsize, int map) {
  extern __shared__ double b2_smem[];
  __shared__ int s_next[8];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  const int ngroups = wpb / gsize;
  const int grp = map ? wib / gsize : wib % ngroups, rank = map ? wib % gsize : wib / ngroups;
  T* base = reinterpret_cast<T*>(b2_smem) + (size_t)wib * (ws_reals + kWarpIntsAsReals * (int)(sizeof(double) / sizeof(T)));
  WarpEnv<T, M> env;
  env.bind(base, reinterpret_cast<int*>(base + ws_reals), jscratch + (size_t)gw * warp_slot_reals(M::nv()));
  env.bar_id = 1 + grp; env.bar_cnt = 32 * gsize;
  const int total = nsteps > 0 ? nsteps : 1;
  int first = (blockIdx.x * ngroups + grp) * gsize;
  while (first < N) {
    const bool mine = first + rank < N;
    const int e = mine ? first + rank : N - 1;
    WFOR(k, M::nq()) env.qpos[k] = st.qpos[(size_t)k * N + e];
    WFOR(k, M::nv()) { env.qvel[k] = st.qvel[(size_t)k * N + e]; env.warm[k] = st.warm ? st.warm[(size_t)k * N + e] : T(0); }
    WFOR(k, M::nu()) env.ctrl[k] = st.ctrl[(size_t)k * N + e];
    env.flags = 0;
    __syncwarp();
    bool frozen = false;  // diverged env: flagged and left as it is (see k_step); its warp keeps pace on the rest pose
    for (int s = 0; s < total; s++) {
      env.check_state();
      if (!frozen && (env.flags & 3)) {
        frozen = true;
        if (mine && st.flags && lane == 0) st.flags[e] |= env.flags;
        WFOR(k, M::nq()) env.qpos[k] = M::qpos0(k);
        WFOR(k, M::nv()) { env.qvel[k] = 0; env.warm[k] = 0; }
        WFOR(k, M::nu()) env.ctrl[k] = 0;
        __syncwarp();
      }
      env.template forward<LS>();
      env.check_acc();
      if (want_derived && s == total - 1 && mine && !frozen) warp_store_derived(env, out, N, e);
      env.template stage_sync<LS>();
      if (nsteps > 0) env.euler();
    }
    if (mine && !frozen) {
      if (nsteps > 0) {
        WFOR(k, M::nq()) st.qpos[(size_t)k * N + e] = env.qpos[k];
        WFOR(k, M::nv()) st.qvel[(size_t)k * N + e] = env.qvel[k];
      }
      if (st.warm) WFOR(k, M::nv()) st.warm[(size_t)k * N + e] = env.warm[k];
      if (st.flags && env.flags && lane == 0) st.flags[e] |= env.flags;
    }
    if (rank == 0 && lane == 0) s_next[grp] = nw + atomicAdd(queue, gsize);
    env.template stage_sync<1>();
    first = s_next[grp];
    env.template stage_sync<1>();
  }
}

// FD linearisation on the warp engine (large models): the work items are (column, env) pairs -- column-major, so that the
// two warps of a lock-step pair run the same kind of column -- and an item is two full rollouts of one step from the
// nominal state: plus / minus for a centred column, plus (or minus) and the nominal step where a control sits at its
// range bound (mjd_transitionFD's one-sidedness, as k_linearize).  Every rollout restores qacc_warmstart.  The columns go
// to A (2nv x 2nv x N) and B (2nv x nu x N), env fastest.  extra = 2 (nq + nv) reals per warp hold the two end states.
template <typename T, class M, int LS>
__global__ void __maxnreg__(200) k_warp_linearize(StateDev<T> st, int N, T eps, int centered, T* A, T* B, T* jscratch, int* queue,
                                                          int ws_reals, int extra, int gsize) {
  extern __shared__ double b2_smem[];
  __shared__ int s_next[8];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  const int ngroups = wpb / gsize, grp = wib % ngroups, rank = wib / ngroups;
  const int ints = kWarpIntsAsReals * (int)(sizeof(double) / sizeof(T));
  T* base = reinterpret_cast<T*>(b2_smem) + (size_t)wib * (ws_reals + ints + extra);
  WarpEnv<T, M> env;
  env.bind(base, reinterpret_cast<int*>(base + ws_reals), jscratch + (size_t)gw * warp_slot_reals(M::nv()));
  env.bar_id = 1 + grp; env.bar_cnt = 32 * gsize;
  const int nq = M::nq(), nv = M::nv(), nu = M::nu(), ndx = 2 * nv, ncol = ndx + nu;
  T* ends[2] = {base + ws_reals + ints, base + ws_reals + ints + nq + nv};  // [0]: minus side (s1), [1]: plus side (s2)
  const int total = N * ncol;
  int flags_all = 0;
  int first = (blockIdx.x * ngroups + grp) * gsize;
  while (first < total) {
    const bool mine = first + rank < total;
    const int item = mine ? first + rank : total - 1;
    const int c = item / N, e = item - c * N;
    int kind, i;
    bool fwd = true, back = centered != 0;
    if (c < nv) { kind = 1; i = c; }
    else if (c < ndx) { kind = 2; i = c - nv; }
    else {
      kind = 3; i = c - ndx;
      const T u = st.ctrl[(size_t)i * N + e], lo = M::actuator_ctrlrange(2 * i), hi = M::actuator_ctrlrange(2 * i + 1);
      const bool lim = M::actuator_ctrllimited(i) != 0;
      fwd = !lim || (u >= lo && u <= hi && u + eps >= lo && u + eps <= hi);
      back = (centered || !fwd) && (!lim || (u - eps >= lo && u - eps <= hi && u >= lo && u <= hi));
    }
    // two rollouts per item whatever the column needs, so that both warps of a pair meet at every barrier
    for (int side = 1; side >= 0; side--) {
      const T delta = side ? (fwd ? eps : T(0)) : (back ? -eps : T(0));
      WFOR(k, nq) env.qpos[k] = st.qpos[(size_t)k * N + e];
      WFOR(k, nv) { env.qvel[k] = st.qvel[(size_t)k * N + e] + ((kind == 2 && k == i) ? delta : T(0)); env.warm[k] = st.warm ? st.warm[(size_t)k * N + e] : T(0); }
      WFOR(k, nu) env.ctrl[k] = st.ctrl[(size_t)k * N + e] + ((kind == 3 && k == i) ? delta : T(0));
      env.flags = 0;
      __syncwarp();
      if (kind == 1 && lane == 0) {  // tangent-space step along dof i (mj_integratePos of a unit vector)
        const int j = M::dof_jntid(i), pa = M::jnt_qposadr(j), d = i - M::jnt_dofadr(j);
        if (M::jnt_type(j) == JNT_FREE && d >= 3) {
          T q[4] = {env.qpos[pa + 3], env.qpos[pa + 4], env.qpos[pa + 5], env.qpos[pa + 6]}, w[3] = {0, 0, 0};
          w[d - 3] = 1;
          quat_integrate(q, w, delta);
          for (int k = 0; k < 4; k++) env.qpos[pa + 3 + k] = q[k];
        } else env.qpos[pa + (M::jnt_type(j) == JNT_FREE ? d : 0)] += delta;
      }
      __syncwarp();
      env.check_state();
      env.template forward<LS>();
      env.check_acc();
      env.template stage_sync<LS>();
      env.euler();
      flags_all |= env.flags;
      T* dst = ends[side];
      WFOR(k, nq) dst[k] = env.qpos[k];
      WFOR(k, nv) dst[nq + k] = env.qvel[k];
      __syncwarp();
    }
    if (mine) {
      const T* s1 = ends[0];
      const T* s2 = ends[1];
      const T ih = (fwd && back) ? T(0.5) / eps : ((fwd || back) ? T(1) / eps : T(0));
      T* out = c < ndx ? A + ((size_t)c) * N + e : B + ((size_t)(c - ndx)) * N + e;
      const size_t rs = (size_t)(c < ndx ? ndx : nu) * N;  // row stride
      if (c < ndx ? A != nullptr : B != nullptr) {
        WFOR(j, M::njnt()) {
          const int pa = M::jnt_qposadr(j), va = M::jnt_dofadr(j);
          if (M::jnt_type(j) == JNT_FREE) {
            T neg[4] = {s1[pa + 3], -s1[pa + 4], -s1[pa + 5], -s1[pa + 6]}, q2[4] = {s2[pa + 3], s2[pa + 4], s2[pa + 5], s2[pa + 6]}, dq[4], rv[3];
            quat_mul(dq, neg, q2);
            quat_to_vel(rv, dq, T(1));
            for (int k = 0; k < 3; k++) { out[(va + k) * rs] = (s2[pa + k] - s1[pa + k]) * ih; out[(va + 3 + k) * rs] = rv[k] * ih; }
          } else out[va * rs] = (s2[pa] - s1[pa]) * ih;
        }
        WFOR(k, nv) out[(nv + k) * rs] = (s2[nq + k] - s1[nq + k]) * ih;
      }
      if (st.flags && flags_all && lane == 0) atomicOr(st.flags + e, flags_all);
    }
    flags_all = 0;
    __syncwarp();
    if (rank == 0 && lane == 0) s_next[grp] = nw + atomicAdd(queue, gsize);
    env.template stage_sync<1>();
    first = s_next[grp];
    env.template stage_sync<1>();
  }
}

}  // namespace b2
