#!/bin/bash
# Builds one library variant for tools/ab_run.sh:  tools/ab_build.sh <name> "<extra nvcc flags>"
# Only the FP64 generic/warp translation unit is recompiled with the extra flags; the other objects are reused from csrc/.
set -e
name=$1; extra=$2
cd "$(dirname "$0")/../mujoco-template_b200/csrc"
mkdir -p ../../ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $extra \
  -c b2_kernels_f64.cu -o ../../ab/f64_$name.o 2> ../../ab/f64_$name.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/lib_$name.so ../../ab/f64_$name.o b2_kernels_f32.o b2_capi.o generated/spec_*.o -lcudart
grep -A2 "k_warp_step_ls.*DimsStatic" ../../ab/f64_$name.ptxas.log | grep -E "registers|spill" | head -4
