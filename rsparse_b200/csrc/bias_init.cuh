// bias_init.cuh -- initialize_biases<T> on the device.
// Reference: inst/include/wrmf_utils.hpp:32-82 (explicit), :84-167 (implicit), :170-183 (dispatch); entry points
// src/wrmf_init.cpp:6-34; caller R/model_WRMF.R:260-289.
// The reference runs five alternating sweeps on one host thread.  Within a sweep every column depends only on the
// *other* side's biases, so a sweep is one kernel with a thread per column that walks the column's entries in the
// reference's order with the reference's arithmetic (accumulator type T, `wsum` recurrences in double).  The two scalar
// pre-passes -- the mean of the ratings (:40-43, a running mean) and the means of the bias vectors (:131-134,:147-149,
// also running means) -- are computed as double-precision tree sums / n: same value up to the last bits of a double.
#pragma once
#include "common.cuh"

namespace b200als {

template <typename T>
__global__ void __launch_bounds__(256) sum_to_partials_kernel(const T* __restrict__ v, long long n, double* __restrict__ partials) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    acc += (double)v[e];
  const double tot = block_sum_double(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// values -= *shift  (wrmf_utils.hpp:48-51)
__global__ void shift_values_kernel(double* __restrict__ a, double* __restrict__ b, long long n, const double* __restrict__ sum,
                                    double inv_n) {
  const double shift = *sum * inv_n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    a[e] -= shift;
    b[e] -= shift;
  }
}

// one explicit sweep (wrmf_utils.hpp:55-66 / :68-79): out[col] = sum(val - other[idx]) / (lambda_use + cnt)
template <typename T>
__global__ void bias_sweep_explicit_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                           const double* __restrict__ val, int n_cols, const T* __restrict__ other,
                                           T* __restrict__ out, T lambda, int dynamic_lambda, int non_negative) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  const int p1 = ptr[col], p2 = ptr[col + 1];
  const T lambda_use = (T)((double)lambda * (dynamic_lambda ? (double)static_cast<T>(p2 - p1) : 1.));
  T acc = T(0);
  for (int ix = p1; ix < p2; ix++) acc = (T)((double)acc + (val[ix] - (double)other[idx[ix]]));
  acc /= lambda_use + static_cast<T>(p2 - p1);
  if (non_negative) acc = (T)fmax((double)T(0), (double)acc);
  out[col] = acc;
}

// implicit pre-pass (wrmf_utils.hpp:101-124): per-row mean and shrinkage factor
__global__ void bias_means_implicit_kernel(const int32_t* __restrict__ ptr, const double* __restrict__ val, int n_rows,
                                           int n_other, double lambda, double* __restrict__ means, double* __restrict__ adjust) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int p1 = ptr[row], p2 = ptr[row + 1];
  if (p2 > p1) {
    double a = 0.0;
    for (int ix = p1; ix < p2; ix++) a += val[ix];
    means[row] = a / (a + (double)(n_other - (p2 - p1)));
    a += (double)(n_other - (p2 - p1));
    adjust[row] = a / (a + lambda);
  } else {
    means[row] = 0.0;
    adjust[row] = (double)n_other / ((double)n_other + lambda);
  }
}

// one implicit sweep (wrmf_utils.hpp:135-141 / :150-156)
template <typename T>
__global__ void bias_sweep_implicit_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                           const double* __restrict__ val, int n_cols, int n_other, const T* __restrict__ other,
                                           const double* __restrict__ other_sum /* nullptr: bias_mean = 0 */,
                                           const double* __restrict__ means, const double* __restrict__ adjust,
                                           double global_bias, int non_negative, T* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  double wsum = (double)n_other;
  double bias_this = other_sum ? (*other_sum / (double)n_other) : 0.0;
  for (int ix = ptr[col]; ix < ptr[col + 1]; ix++) {
    const double c1 = val[ix] - 1;
    wsum += c1;
    bias_this += (c1 * ((double)other[idx[ix]] - bias_this)) / wsum;
  }
  T b = (T)((means[col] - bias_this - global_bias) * adjust[col]);
  if (non_negative) b = (T)fmax((double)T(0), (double)b);
  out[col] = b;
}

}  // namespace b200als
