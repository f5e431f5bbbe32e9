// Batched discrete-time LQR synthesis on the device: the gain the reference's example
// controllers compute once per system on the host with scipy.linalg.solve_discrete_are
// (reference examples/drone/controllers/lqr.py:350-378, examples/humanoid/controllers/lqr.py:114-115), for every env of a
// batch from its own FD linearisation (A, B) -- the arrays b2_linearize / b2_control_tick write, in their SoA layout.
//
// One warp per env, all matrices of the env in shared memory.  The DARE  P = A'PA - A'PB (R + B'PB)^-1 B'PA + Q  is solved
// by the structure-preserving doubling iteration (k doublings = 2^k steps of the Riccati recursion, quadratic convergence):
//     W = I + G H;   [V1 V2] = W^-1 [A G];   H <- H + A' H V1;   G <- G + A V2 A';   A <- A V1        (G0 = B R^-1 B', H0 = Q)
// W^-1 [A G] by Gauss-Jordan elimination with partial pivoting on the augmented nx x 3nx block; afterwards
//     K = (R + B'PB)^-1 B'PA      (u = -K x).
// Work per doubling: five nx^3 products and one nx x 3nx elimination; nx = 4 (cartpole), 12 (drone), 54 (humanoid).
#pragma once
#include "b2_math.cuh"

namespace b2 {

// shared-memory reals one env needs: A, G, H, the augmented [W | V1 V2] block (3 nx^2) and two scratch matrices
__host__ __device__ inline size_t dare_ws_reals(int nx, int nu) { return (size_t)8 * nx * nx + (size_t)2 * nx * nu + (size_t)2 * nu * nu + 32; }

template <typename T>
B2_DEV void dare_matmul(T* C, const T* X, bool tx, const T* Y, bool ty, int n, int lane) {
  // C (n x n) = op(X) op(Y), one entry per lane and sweep (C is write-only: it may hold anything, NaN included)
  for (int e = lane; e < n * n; e += 32) {
    const int i = e / n, j = e - i * n;
    T s = 0;
    for (int k = 0; k < n; k++) s += (tx ? X[k * n + i] : X[i * n + k]) * (ty ? Y[j * n + k] : Y[k * n + j]);
    C[e] = s;
  }
  __syncwarp();
}

// In-place Gauss-Jordan with partial pivoting on M (n rows, w columns, the first n of them the matrix to invert): on return the
// trailing w - n columns hold inverse * (their old content).  Returns false when a pivot underflows.
template <typename T>
B2_DEV bool dare_gauss_jordan(T* M, int n, int w, int lane) {
  for (int c = 0; c < n; c++) {
    // pivot search over rows c..n-1 of column c (warp arg-max)
    T best = -1; int row = c;
    for (int r = c + lane; r < n; r += 32) { const T a = fabs(M[r * w + c]); if (a > best) { best = a; row = r; } }
    for (int o = 16; o > 0; o >>= 1) {
      const T ob = __shfl_xor_sync(0xffffffffu, best, o); const int orow = __shfl_xor_sync(0xffffffffu, row, o);
      if (ob > best || (ob == best && orow < row)) { best = ob; row = orow; }
    }
    if (!(best > Num<T>::minval())) return false;
    if (row != c) for (int k = lane; k < w; k += 32) { const T t = M[c * w + k]; M[c * w + k] = M[row * w + k]; M[row * w + k] = t; }
    __syncwarp();
    const T inv = T(1) / M[c * w + c];
    __syncwarp();
    for (int k = lane; k < w; k += 32) M[c * w + k] *= inv;
    __syncwarp();
    // eliminate column c from every other row: entries (r, k), k > c (column c itself is not needed again)
    for (int e = lane; e < n * (w - c - 1); e += 32) {
      const int r = e / (w - c - 1), k = c + 1 + e - r * (w - c - 1);
      if (r != c) M[r * w + k] -= M[r * w + c] * M[c * w + k];
    }
    __syncwarp();
  }
  return true;
}

// A (nx, nx, N), B (nx, nu, N), K (nu, nx, N), P (nx, nx, N): env fastest (element (i, j, e) at [(i * cols + j) * N + e]).
// Q (nx x nx), R (nu x nu), Rinv (nu x nu): row-major, shared by all envs.  status[e]: doublings used, negated on failure.
template <typename T>
__global__ void __launch_bounds__(128) k_dare(const T* __restrict__ Ag, const T* __restrict__ Bg, const T* __restrict__ Q,
                                              const T* __restrict__ R, const T* __restrict__ Rinv, int nx, int nu, int N,
                                              int max_doublings, T tol, T* Kg, T* Pg, int* status) {
  extern __shared__ double b2_dare_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, e = blockIdx.x * (blockDim.x >> 5) + wib;
  if (e >= N) return;
  const int nn = nx * nx;
  T* ws = reinterpret_cast<T*>(b2_dare_smem) + (size_t)wib * dare_ws_reals(nx, nu);
  T *A = ws, *G = A + nn, *H = G + nn, *M = H + nn /* nx x 3nx */, *T1 = M + 3 * nn, *T2 = T1 + nn, *Bm = T2 + nn /* nx x nu */,
    *BR = Bm + nx * nu /* nx x nu: B Rinv, later P B */, *S = BR + nx * nu /* nu x 2nu */;
#ifdef B2_DARE_POISON  // debugging aid: every read of workspace the kernel has not written shows up as a failed env
  for (int k = lane; k < (int)dare_ws_reals(nx, nu); k += 32) ws[k] = T(__int_as_float(0x7fc00000));
  __syncwarp();
#endif
  for (int k = lane; k < nn; k += 32) { A[k] = Ag[(size_t)k * N + e]; H[k] = Q[k]; }
  for (int k = lane; k < nx * nu; k += 32) Bm[k] = Bg[(size_t)k * N + e];
  __syncwarp();
  // G0 = B Rinv B'
  for (int k = lane; k < nx * nu; k += 32) {
    const int i = k / nu, a = k - i * nu;
    T s = 0;
    for (int b = 0; b < nu; b++) s += Bm[i * nu + b] * Rinv[b * nu + a];
    BR[k] = s;
  }
  __syncwarp();
  for (int k = lane; k < nn; k += 32) {
    const int i = k / nx, j = k - i * nx;
    T s = 0;
    for (int a = 0; a < nu; a++) s += BR[i * nu + a] * Bm[j * nu + a];
    G[k] = s;
  }
  __syncwarp();
  int used = 0;
  bool ok = true;
  for (int it = 0; it < max_doublings; it++) {
    // M = [I + G H | A | G]
    dare_matmul(T1, G, false, H, false, nx, lane);
    for (int k = lane; k < nn; k += 32) {
      const int i = k / nx, j = k - i * nx;
      M[i * 3 * nx + j] = T1[k] + (i == j ? T(1) : T(0));
      M[i * 3 * nx + nx + j] = A[k];
      M[i * 3 * nx + 2 * nx + j] = G[k];
    }
    __syncwarp();
    ok = dare_gauss_jordan(M, nx, 3 * nx, lane);
    if (!ok) break;
    // T1 = V1, T2 = V2 (contiguous copies)
    for (int k = lane; k < nn; k += 32) { const int i = k / nx, j = k - i * nx; T1[k] = M[i * 3 * nx + nx + j]; T2[k] = M[i * 3 * nx + 2 * nx + j]; }
    __syncwarp();
    // H <- H + A' (H V1): M[0..nn) = H V1, then accumulate; track the change of H
    dare_matmul(M, H, false, T1, false, nx, lane);
    T dmax = 0, hmax = 0;
    for (int k = lane; k < nn; k += 32) {
      const int i = k / nx, j = k - i * nx;
      T s = 0;
      for (int q = 0; q < nx; q++) s += A[q * nx + i] * M[q * nx + j];
      M[nn + k] = s;  // increment, symmetrised below
    }
    __syncwarp();
    for (int k = lane; k < nn; k += 32) {
      const int i = k / nx, j = k - i * nx;
      const T inc = T(0.5) * (M[nn + k] + M[nn + j * nx + i]);
      const T h = H[k] + inc;
      dmax = fmax(dmax, fabs(inc)); hmax = fmax(hmax, fabs(h));
      M[2 * nn + k] = h;
    }
    __syncwarp();
    for (int k = lane; k < nn; k += 32) H[k] = M[2 * nn + k];
    // G <- G + A V2 A': M[0..nn) = A V2, then (A V2) A' symmetrised
    dare_matmul(M, A, false, T2, false, nx, lane);
    dare_matmul(M + nn, M, false, A, true, nx, lane);
    for (int k = lane; k < nn; k += 32) { const int i = k / nx, j = k - i * nx; G[k] += T(0.5) * (M[nn + k] + M[nn + j * nx + i]); }
    // A <- A V1
    dare_matmul(M, A, false, T1, false, nx, lane);
    for (int k = lane; k < nn; k += 32) A[k] = M[k];
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) { dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o)); hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o)); }
    used = it + 1;
    if (it >= 3 && dmax <= tol * fmax(hmax, T(1))) break;
  }
  // P = H; K = (R + B'PB)^-1 B'PA with the ORIGINAL A (re-read)
  for (int k = lane; k < nn; k += 32) { Pg[(size_t)k * N + e] = H[k]; A[k] = Ag[(size_t)k * N + e]; }
  __syncwarp();
  for (int k = lane; k < nx * nu; k += 32) {  // BR = P B (nx x nu)
    const int i = k / nu, a = k - i * nu;
    T s = 0;
    for (int q = 0; q < nx; q++) s += H[i * nx + q] * Bm[q * nu + a];
    BR[k] = s;
  }
  __syncwarp();
  // augmented [R + B'PB | B'PA] : nu x (nu + nx), in M
  const int w = nu + nx;
  for (int k = lane; k < nu * nu; k += 32) {
    const int a = k / nu, b = k - a * nu;
    T s = R[k];
    for (int q = 0; q < nx; q++) s += Bm[q * nu + a] * BR[q * nu + b];
    M[a * w + b] = s;
  }
  for (int k = lane; k < nu * nx; k += 32) {
    const int a = k / nx, j = k - a * nx;
    T s = 0;
    for (int q = 0; q < nx; q++) s += BR[q * nu + a] * A[q * nx + j];  // (P B)' A = B' P A (P symmetric)
    M[a * w + nu + j] = s;
  }
  __syncwarp();
  ok = dare_gauss_jordan(M, nu, w, lane) && ok;
  for (int k = lane; k < nu * nx; k += 32) { const int a = k / nx, j = k - a * nx; Kg[(size_t)k * N + e] = M[a * w + nu + j]; }
  if (lane == 0 && status) status[e] = ok ? used : -used - 1;
  (void)S;
}

}  // namespace b2

namespace b2 {

// Thread-per-env variant for small systems (pendulum 2 x 1, cartpole 4 x 1, ...): compile-time sizes, every matrix in
// registers, same iteration and stopping rule as k_dare.  A warp per 4 x 4 system leaves 30 lanes idle; this form
// synthesises the gains of 65,536 cartpoles in a fraction of a millisecond, which makes re-synthesis at every control tick
// (time-varying LQR from the tick's own (A, B)) affordable.
template <typename T, int N>
B2_DEV void small_mul(T* C, const T* X, const T* Y) {  // C = X Y
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {
      T s = 0;
#pragma unroll
      for (int k = 0; k < N; k++) s += X[i * N + k] * Y[k * N + j];
      C[i * N + j] = s;
    }
}
template <typename T, int N>
B2_DEV void small_mul_tn(T* C, const T* X, const T* Y) {  // C = X' Y
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {
      T s = 0;
#pragma unroll
      for (int k = 0; k < N; k++) s += X[k * N + i] * Y[k * N + j];
      C[i * N + j] = s;
    }
}
template <typename T, int N>
B2_DEV void small_mul_nt(T* C, const T* X, const T* Y) {  // C = X Y'
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {
      T s = 0;
#pragma unroll
      for (int k = 0; k < N; k++) s += X[i * N + k] * Y[j * N + k];
      C[i * N + j] = s;
    }
}
// Gauss-Jordan with partial pivoting on [W | R1 R2] (N x N and two N x N right-hand sides), fully unrolled: the row swap is a
// chain of selects, so that no index is a run-time value and everything stays in registers.
template <typename T, int N, int RHS>
B2_DEV bool small_solve(T* W, T* R) {  // R: N x (RHS) row-major, overwritten by W^-1 R
  bool ok = true;
#pragma unroll
  for (int c = 0; c < N; c++) {
    // bring the largest |W[r][c]|, r >= c, to row c
#pragma unroll
    for (int r = c + 1; r < N; r++) {
      const bool sw = fabs(W[r * N + c]) > fabs(W[c * N + c]);
#pragma unroll
      for (int k = 0; k < N; k++) { const T a = W[c * N + k], b = W[r * N + k]; W[c * N + k] = sw ? b : a; W[r * N + k] = sw ? a : b; }
#pragma unroll
      for (int k = 0; k < RHS; k++) { const T a = R[c * RHS + k], b = R[r * RHS + k]; R[c * RHS + k] = sw ? b : a; R[r * RHS + k] = sw ? a : b; }
    }
    if (!(fabs(W[c * N + c]) > Num<T>::minval())) ok = false;
    const T inv = T(1) / W[c * N + c];
#pragma unroll
    for (int k = 0; k < N; k++) W[c * N + k] *= inv;
#pragma unroll
    for (int k = 0; k < RHS; k++) R[c * RHS + k] *= inv;
#pragma unroll
    for (int r = 0; r < N; r++) {
      if (r == c) continue;
      const T f = W[r * N + c];
#pragma unroll
      for (int k = 0; k < N; k++) W[r * N + k] -= f * W[c * N + k];
#pragma unroll
      for (int k = 0; k < RHS; k++) R[r * RHS + k] -= f * R[c * RHS + k];
    }
  }
  return ok;
}

template <typename T, int NX, int NU>
__global__ void __launch_bounds__(64) k_dare_small(const T* __restrict__ Ag, const T* __restrict__ Bg, const T* __restrict__ Q,
                                                   const T* __restrict__ R, const T* __restrict__ Rinv, int N, int max_doublings, T tol,
                                                   T* Kg, T* Pg, int* status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  constexpr int NN = NX * NX;
  T A0[NN], A[NN], G[NN], H[NN], B[NX * NU];
#pragma unroll
  for (int k = 0; k < NN; k++) { A0[k] = Ag[(size_t)k * N + e]; A[k] = A0[k]; H[k] = Q[k]; }
#pragma unroll
  for (int k = 0; k < NX * NU; k++) B[k] = Bg[(size_t)k * N + e];
  {  // G0 = B Rinv B'
    T BR[NX * NU];
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int a = 0; a < NU; a++) { T s = 0;
#pragma unroll
        for (int b = 0; b < NU; b++) s += B[i * NU + b] * Rinv[b * NU + a];
        BR[i * NU + a] = s; }
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int j = 0; j < NX; j++) { T s = 0;
#pragma unroll
        for (int a = 0; a < NU; a++) s += BR[i * NU + a] * B[j * NU + a];
        G[i * NX + j] = s; }
  }
  int used = 0;
  bool ok = true;
#pragma unroll 1
  for (int it = 0; it < max_doublings; it++) {
    T W[NN], V[NX * 2 * NX];  // V = [A | G] -> W^-1 [A | G]
    small_mul<T, NX>(W, G, H);
#pragma unroll
    for (int i = 0; i < NX; i++) {
      W[i * NX + i] += T(1);
#pragma unroll
      for (int j = 0; j < NX; j++) { V[i * 2 * NX + j] = A[i * NX + j]; V[i * 2 * NX + NX + j] = G[i * NX + j]; }
    }
    ok = small_solve<T, NX, 2 * NX>(W, V);
    if (!ok) break;
    T V1[NN], V2[NN], t1[NN], t2[NN];
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int j = 0; j < NX; j++) { V1[i * NX + j] = V[i * 2 * NX + j]; V2[i * NX + j] = V[i * 2 * NX + NX + j]; }
    small_mul<T, NX>(t1, H, V1);
    small_mul_tn<T, NX>(t2, A, t1);  // A' H V1
    T dmax = 0, hmax = 0;
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int j = 0; j < NX; j++) {
        const T inc = T(0.5) * (t2[i * NX + j] + t2[j * NX + i]);
        t1[i * NX + j] = H[i * NX + j] + inc;
        dmax = fmax(dmax, fabs(inc)); hmax = fmax(hmax, fabs(t1[i * NX + j]));
      }
#pragma unroll
    for (int k = 0; k < NN; k++) H[k] = t1[k];
    small_mul<T, NX>(t1, A, V2);
    small_mul_nt<T, NX>(t2, t1, A);  // A V2 A'
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int j = 0; j < NX; j++) t1[i * NX + j] = G[i * NX + j] + T(0.5) * (t2[i * NX + j] + t2[j * NX + i]);
#pragma unroll
    for (int k = 0; k < NN; k++) G[k] = t1[k];
    small_mul<T, NX>(t1, A, V1);
#pragma unroll
    for (int k = 0; k < NN; k++) A[k] = t1[k];
    used = it + 1;
    if (it >= 3 && dmax <= tol * fmax(hmax, T(1))) break;
  }
#pragma unroll
  for (int k = 0; k < NN; k++) Pg[(size_t)k * N + e] = H[k];
  // K = (R + B'PB)^-1 B'P A0
  T PB[NX * NU], S[NU * NU], RH[NU * NX];
#pragma unroll
  for (int i = 0; i < NX; i++)
#pragma unroll
    for (int a = 0; a < NU; a++) { T s = 0;
#pragma unroll
      for (int q = 0; q < NX; q++) s += H[i * NX + q] * B[q * NU + a];
      PB[i * NU + a] = s; }
#pragma unroll
  for (int a = 0; a < NU; a++) {
#pragma unroll
    for (int b = 0; b < NU; b++) { T s = R[a * NU + b];
#pragma unroll
      for (int q = 0; q < NX; q++) s += B[q * NU + a] * PB[q * NU + b];
      S[a * NU + b] = s; }
#pragma unroll
    for (int j = 0; j < NX; j++) { T s = 0;
#pragma unroll
      for (int q = 0; q < NX; q++) s += PB[q * NU + a] * A0[q * NX + j];
      RH[a * NX + j] = s; }
  }
  ok = small_solve<T, NU, NX>(S, RH) && ok;
#pragma unroll
  for (int k = 0; k < NU * NX; k++) Kg[(size_t)k * N + e] = RH[k];
  if (status) status[e] = ok ? used : -used - 1;
}

}  // namespace b2
