// TEST INFRASTRUCTURE ONLY.  Lets the reference's CUDA kernels (the anonymous-namespace blocks of
// detectron/ops/*.cu, extracted UNMODIFIED at build time into oracle/_ref/gen/*.inc by
// oracle/build_ref_kernels.py) compile and run as ordinary host C++: one "thread" walks the whole
// CUDA_1D_KERNEL_LOOP range, atomics become plain read-modify-writes, and the unqualified math calls
// resolve to the overloads CUDA's device headers would have chosen (float arguments -> float functions).
// Every operation in those kernels is an IEEE basic operation, an integer operation or logf, so the host
// run reproduces the device arithmetic except for logf's last ulp and the (sequential) atomic order.
#pragma once
#include <cfloat>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define CUDA_1D_KERNEL_LOOP(i, n) for (int i = 0; i < (n); ++i)

namespace cuda_host {
template <typename T>
inline T atomicAdd(T* address, T val) {
  const T old = *address;
  *address = old + val;
  return old;
}
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline float log(float x) { return logf(x); }
inline float floor(float x) { return floorf(x); }
inline float ceil(float x) { return ceilf(x); }
}  // namespace cuda_host
