// Kernels of the warp engine (one env per warp, workspace in shared memory).
#pragma once
#include "b2_kernel_templates.cuh"
#include "b2_warp_engine.cuh"

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
  env.bind(base, reinterpret_cast<int*>(base + ws_reals), jscratch + (size_t)gw * WarpCaps::NEFC * (M::nv() + 6));
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
__global__ void __launch_bounds__(256, 1) k_warp_step_ls(StateDev<T> st, DerivedDev<T> out, int want_derived, int N, int nsteps,
                                                        T* jscratch, int* queue, int ws_reals, int gsize, int map) {
  extern __shared__ double b2_smem[];
  __shared__ int s_next[8];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  const int ngroups = wpb / gsize;
  const int grp = map ? wib / gsize : wib % ngroups, rank = map ? wib % gsize : wib / ngroups;
  T* base = reinterpret_cast<T*>(b2_smem) + (size_t)wib * (ws_reals + kWarpIntsAsReals * (int)(sizeof(double) / sizeof(T)));
  WarpEnv<T, M> env;
  env.bind(base, reinterpret_cast<int*>(base + ws_reals), jscratch + (size_t)gw * WarpCaps::NEFC * (M::nv() + 6));
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

}  // namespace b2
