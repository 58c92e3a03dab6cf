// caffe2::math CPU primitives used by the reference ops on the path (plain loops; the
// real ones are Eigen/BLAS elementwise kernels with identical per-element arithmetic).
#pragma once
#include "caffe2/core/operator.h"
#include "caffe2/core/logging.h"
namespace caffe2 { namespace math {
template <typename T, class Context>
void Set(const int64_t n, const T alpha, T* y, Context*) { for (int64_t i = 0; i < n; ++i) y[i] = alpha; }
template <typename TA, typename T, class Context>
void Scale(const int64_t n, const TA alpha, const T* x, T* y, Context*) { for (int64_t i = 0; i < n; ++i) y[i] = x[i] * alpha; }
template <typename T, class Context>
void Add(const int64_t n, const T* a, const T* b, T* y, Context*) { for (int64_t i = 0; i < n; ++i) y[i] = a[i] + b[i]; }
template <typename TA, typename T, class Context>
void Axpy(const int64_t n, const TA alpha, const T* x, T* y, Context*) { for (int64_t i = 0; i < n; ++i) y[i] += alpha * x[i]; }
}}  // namespace caffe2::math
