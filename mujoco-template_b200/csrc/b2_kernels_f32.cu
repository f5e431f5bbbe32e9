// FP32 instantiation of the lane-engine kernels (optional reduced-precision mode).
#define B2_REAL float
#define B2_SUFFIX _f32
#include "b2_kernels_impl.cuh"
