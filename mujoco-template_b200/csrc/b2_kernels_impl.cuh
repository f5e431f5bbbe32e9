// Generic (runtime-model) instantiation of the lane-engine kernels for ONE arithmetic type.
// Included by b2_kernels_f64.cu and b2_kernels_f32.cu with B2_REAL / B2_SUFFIX defined; each
// translation unit owns its own __constant__ model images (one per size class), so the two
// precisions do not share the 64 KB constant bank.
#include <cuda_runtime.h>

#include "b2_kernel_templates.cuh"
#include "b2_kernels.h"
#include "b2_warp_kernels.cuh"

namespace b2 {

typedef B2_REAL real;

__constant__ DevModel<real, DimsTiny> c_model_tiny;
__constant__ DevModel<real, DimsSmall> c_model_small;
__constant__ DevModel<real, DimsLarge> c_model_large;

template <class D> struct ConstImage;
template <> struct ConstImage<DimsTiny> { static B2_DEV const DevModel<real, DimsTiny>& get() { return c_model_tiny; } };
template <> struct ConstImage<DimsSmall> { static B2_DEV const DevModel<real, DimsSmall>& get() { return c_model_small; } };
template <> struct ConstImage<DimsLarge> { static B2_DEV const DevModel<real, DimsLarge>& get() { return c_model_large; } };

// provider that reads the constant-bank image: uniform operands, any model of the size class
template <class DD>
struct RuntimeModel {
  typedef DD D;
#define X(name) static B2_DEV int name() { return ConstImage<D>::get().name; }
  B2_MODEL_INT_SCALARS(X)
#undef X
#define X(name) static B2_DEV real name() { return ConstImage<D>::get().name; }
  B2_MODEL_REAL_SCALARS(X)
#undef X
#define X(name, cap) static B2_DEV int name(int i) { return ConstImage<D>::get().name[i]; }
  B2_MODEL_INT_ARRAYS(X)
#undef X
#define X(name, cap) static B2_DEV real name(int i) { return ConstImage<D>::get().name[i]; }
  B2_MODEL_REAL_ARRAYS(X)
#undef X
};

// Global-memory copy of the large image for the warp engine: its lanes read the model with
// lane-dependent indices, which the constant cache would serialise (one address per cycle);
// read-only global loads are gathered by L1 instead.
__device__ DevModel<real, DimsLarge> g_model_large;
struct GlobalModelLarge {
  typedef DimsLarge D;
#define X(name) static B2_DEV int name() { return __ldg(&g_model_large.name); }
  B2_MODEL_INT_SCALARS(X)
#undef X
#define X(name) static B2_DEV real name() { return __ldg(&g_model_large.name); }
  B2_MODEL_REAL_SCALARS(X)
#undef X
#define X(name, cap) static B2_DEV int name(int i) { return __ldg(&g_model_large.name[i]); }
  B2_MODEL_INT_ARRAYS(X)
#undef X
#define X(name, cap) static B2_DEV real name(int i) { return __ldg(&g_model_large.name[i]); }
  B2_MODEL_REAL_ARRAYS(X)
#undef X
};

// register-resident FMA chain: measures the CUDA-core FP pipe peak the roofline is quoted against
__global__ void __launch_bounds__(256) k_fma_peak(real* out, int iters, real a, real b) {
  real x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
    x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
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
  cudaError_t e = cudaMemcpyToSymbolAsync(g_model_large, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  if ((e = upload_tri_tables(s)) != cudaSuccess) return e;
  if ((e = upload_ld_plan(v, s)) != cudaSuccess) return e;
  if ((e = upload_chol_plan(v.nv, s)) != cudaSuccess) return e;
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
// returns measured TFLOP/s (2 flops per FMA) or a negative cudaError
double B2_FN(b2k_fma_peak)(void* stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, blocks = sms * 8, iters = 1 << 15;
  real* out = nullptr;
  if (cudaMalloc(&out, sizeof(real) * threads * blocks) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaStream_t s = (cudaStream_t)stream;
  k_fma_peak<<<blocks, threads, 0, s>>>(out, iters, (real)0.999999, (real)1e-7);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, s);
    k_fma_peak<<<blocks, threads, 0, s>>>(out, iters, (real)0.999999, (real)1e-7);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (err != cudaSuccess) return -(double)err;
  const double flops = 2.0 * 8.0 * iters * (double)threads * blocks;
  return flops / (best * 1e-3) / 1e12;
}
int B2_FN(b2k_step)(int cls, const b2_state* st, const b2_derived* out, int count, int N, int nsteps, const void* gain,
                    const b2_state* park, void* stream) {
  const int threads = 128, blocks = (count + threads - 1) / threads;
  B2_DISPATCH(cls, (k_step<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       to_dev<real>(st), to_dev<real>(out), out != nullptr, count, N, nsteps, (const real*)gain, to_dev<real>(park))));
  return (int)cudaGetLastError();
}
// ---- warp engine (large models): launch geometry and per-warp scratch are sized by the host
static int warp_ws_reals_of(const b2m_view* v) {
  return warp_ws_reals<void>(v->nq, v->nv, v->nu, v->nbody, v->njnt, v->ngeom, v->ntendon);
}
static size_t warp_block_smem(const b2m_view* v, int wpb) {
  size_t extra = 0;
  if (const char* x = getenv("B2_WARP_EXTRA_SMEM")) extra = (size_t)atoi(x);  // tuning: lowers the resident blocks per SM
  return extra + (size_t)wpb * ((size_t)warp_ws_reals_of(v) * sizeof(real) + (size_t)kWarpIntsAsReals * sizeof(double));
}
// lock-step launch shape (k_warp_step_ls): B2_WARP_LOCKSTEP=0 falls back to independent two-warp blocks
static int warp_lockstep() {  // 0: independent warps, 1: lock step everywhere, 2: lock step outside the Newton loop
  const char* x = getenv("B2_WARP_LOCKSTEP");
  return x ? atoi(x) : 1;
}
// chooses warps-per-block / grid so that every SM is filled; returns the number of warp slots
int B2_FN(b2k_warp_plan)(const b2m_view* v, int N, int* out_wpb, int* out_blocks) {
  int dev = 0, sms = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (warp_lockstep()) {
    // one block per SM with as many warps (envs) as the shared-memory workspace allows, at most 8
    int wpb = 2;
    if (const char* x = getenv("B2_WARP_LS_WPB")) wpb = atoi(x) == 1 ? 1 : 2;  // the kernels are compiled for 64-thread blocks
    while (wpb > 1 && warp_block_smem(v, wpb) + 1024 > (size_t)smem_max) wpb--;
    const size_t smem = warp_block_smem(v, wpb);
    auto kern = warp_lockstep() == 2 ? k_warp_step_ls<real, GlobalModelLarge, 2> : k_warp_step_ls<real, GlobalModelLarge, 1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem) != cudaSuccess || per_sm < 1) return -1;
    int blocks = sms * per_sm;
    const int need = (N + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    *out_wpb = wpb; *out_blocks = blocks;
    return blocks * wpb;
  }
  const int wpb = 2;
  const size_t smem = warp_block_smem(v, wpb);
  auto kern = k_warp_step<real, GlobalModelLarge>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem) != cudaSuccess || per_sm < 1) return -1;
  int blocks = sms * per_sm;
  const int need = (N + wpb - 1) / wpb;
  if (blocks > need) blocks = need;
  *out_wpb = wpb; *out_blocks = blocks;
  return blocks * wpb;
}
size_t B2_FN(b2k_warp_scratch_bytes)(const b2m_view* v, int slots) {
  return (size_t)slots * warp_slot_reals(v->nv) * sizeof(real);
}
int B2_FN(b2k_warp_step)(const b2m_view* v, const b2_state* st, const b2_derived* out, int N, int nsteps, void* jscratch,
                         void* counter, int wpb, int blocks, void* stream) {
  const size_t smem = warp_block_smem(v, wpb);
  // envs are handed out through a work queue: the first gridDim * wpb statically, the rest by atomic counter
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  // lock-step group: B2_WARP_LS_GROUP warps (default: the whole block), B2_WARP_LS_MAP picks which warps form a group
  int gsize = wpb, map = 0;
  if (const char* x = getenv("B2_WARP_LS_GROUP")) { const int g = atoi(x); if (g > 0 && wpb % g == 0) gsize = g; }
  if (const char* x = getenv("B2_WARP_LS_MAP")) map = atoi(x) != 0;
  if (warp_lockstep() == 2)
    k_warp_step_ls<real, GlobalModelLarge, 2><<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        to_dev<real>(st), to_dev<real>(out), out != nullptr, N, nsteps, (real*)jscratch, (int*)counter, warp_ws_reals_of(v), gsize, map);
  else if (warp_lockstep())
    k_warp_step_ls<real, GlobalModelLarge, 1><<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        to_dev<real>(st), to_dev<real>(out), out != nullptr, N, nsteps, (real*)jscratch, (int*)counter, warp_ws_reals_of(v), gsize, map);
  else
    k_warp_step<real, GlobalModelLarge><<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        to_dev<real>(st), to_dev<real>(out), out != nullptr, N, nsteps, (real*)jscratch, (int*)counter, warp_ws_reals_of(v));
  return (int)cudaGetLastError();
}
// FD linearisation on the warp engine: same grid and scratch slots as the step plan (wpb warps per block)
int B2_FN(b2k_warp_linearize)(const b2m_view* v, const b2_state* st, int N, double eps, int centered, void* A, void* B, void* jscratch,
                              void* counter, int wpb, int blocks, void* stream) {
  const int extra = 2 * (v->nq + v->nv);
  const size_t smem = warp_block_smem(v, wpb) + (size_t)wpb * extra * sizeof(real);
  auto kern = k_warp_linearize<real, GlobalModelLarge, 1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem) != cudaSuccess || per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  if (blocks > sms * per_sm) blocks = sms * per_sm;  // persistent: never more blocks than fit at once (scratch slots are per block)
  e = cudaMemsetAsync(counter, 0, sizeof(int), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  kern<<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(to_dev<real>(st), N, (real)eps, centered, (real*)A, (real*)B, (real*)jscratch,
                                                         (int*)counter, warp_ws_reals_of(v), extra, wpb);
  return (int)cudaGetLastError();
}
int B2_FN(b2k_linearize)(int cls, const b2_state* st, int count, int N, int ncol, double eps, int centered, void* A, void* B,
                         const void* gain, const b2_state* shadow, void* stream) {
  const int threads = 128;
  const long long total = (long long)count * ncol;  // ncol: FD tasks per env (b2_capi: nv + 1 under Euler, else 2nv + nu)
  const int blocks = (int)((total + threads - 1) / threads);
  B2_DISPATCH(cls, (k_linearize<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       to_dev<real>(st), count, N, (real)eps, centered, (real*)A, (real*)B, (const real*)gain, to_dev<real>(shadow))));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_commit_state)(const b2_state* st, const b2_state* shadow, int count, int N, int nq, int nv, int nu, void* stream) {
  int bx = (count + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  k_commit_state<real><<<dim3(bx, nq + 2 * nv + nu), 256, 0, (cudaStream_t)stream>>>(to_dev<real>(st), to_dev<real>(shadow), count, N, nq, nv, nu);
  return (int)cudaGetLastError();
}
int B2_FN(b2k_jacobian)(int cls, const b2_state* st, int N, int kind, int objid, void* jacp, void* jacr, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_jacobian<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       to_dev<real>(st), N, kind, objid, (real*)jacp, (real*)jacr)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_inverse)(int cls, const b2_state* st, int N, const void* qacc, void* qfrc, void* moment, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_inverse<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       to_dev<real>(st), N, (const real*)qacc, (real*)qfrc, (real*)moment)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_lqr_control)(int cls, const b2_state* st, int count, int N, const void* gain, void* stream) {
  const int threads = 128, blocks = (count + threads - 1) / threads;
  B2_DISPATCH(cls, (k_lqr_control<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       to_dev<real>(st), count, N, (const real*)gain)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_integrate_pos)(int cls, void* qpos, const void* qvel, double dt, int N, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_integrate_pos<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       (real*)qpos, (const real*)qvel, (real)dt, N)));
  return (int)cudaGetLastError();
}
int B2_FN(b2k_differentiate_pos)(int cls, void* out, double dt, const void* q1, const void* q2, int N, void* stream) {
  const int threads = 128, blocks = (N + threads - 1) / threads;
  B2_DISPATCH(cls, (k_differentiate_pos<real, DD, RuntimeModel<DD>><<<blocks, threads, 0, (cudaStream_t)stream>>>(
                       (real*)out, (real)dt, (const real*)q1, (const real*)q2, N)));
  return (int)cudaGetLastError();
}

}  // namespace b2
