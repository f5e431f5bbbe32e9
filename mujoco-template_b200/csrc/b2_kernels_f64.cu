// FP64 instantiation of the lane-engine kernels (the parity path).
#define B2_REAL double
#define B2_SUFFIX _f64
#include "b2_kernels_impl.cuh"
