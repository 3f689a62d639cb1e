// engine_helpers.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, n ? n : 1);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  int32_t* i32() const { return reinterpret_cast<int32_t*>(p); }
  float* f32() const { return reinterpret_cast<float*>(p); }
  double* f64() const { return reinterpret_cast<double*>(p); }
  unsigned long long* u64() const { return reinterpret_cast<unsigned long long*>(p); }
};

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (TO)in[i];
}
__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out, int accumulate) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += partials[i];
    out[0] = accumulate ? out[0] + s : s;
  }
}
// regulariser: sum_j w_j ||x_j||^2 (w_j = cnt_X[j] or 1), per-block partials in double
template <typename T>
__global__ void __launch_bounds__(256) sqnorm_kernel(const T* __restrict__ X, int k, long long n, const T* __restrict__ cnt,
                                                     double* __restrict__ partials) {
  __shared__ double s_red[32];
  double acc = 0.0;
  const long long total = n * (long long)k;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const double v = (double)X[e];
    const double w = cnt ? (double)cnt[e / k] : 1.0;
    acc += v * v * w;
  }
  const double tot = block_sum_double(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}
// rows by length class: 0 -> empty (Y row zeroed here), 1..max_short -> short list, else long list
__global__ void classify_rows_kernel(const int32_t* __restrict__ ptr, int n_rows, int max_short, int32_t* __restrict__ short_list,
                                     int32_t* __restrict__ long_list, int* __restrict__ counts /* [3]: short, long, empty */) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int n = ptr[r + 1] - ptr[r];
  if (n <= 0) atomicAdd(&counts[2], 1);
  else if (n <= max_short) short_list[atomicAdd(&counts[0], 1)] = r;
  else long_list[atomicAdd(&counts[1], 1)] = r;
}
// flag[0] |= 1 if any value is below 1 (or NaN)
__global__ void any_below_one_kernel(const float* __restrict__ v, long long n, int* __restrict__ flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= !(v[i] >= 1.0f);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}
template <typename T>
__global__ void zero_empty_rows_kernel(const int32_t* __restrict__ ptr, int n_rows, int k, T* __restrict__ Y) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n_rows * k) return;
  const int r = (int)(e / k);
  if (ptr[r + 1] - ptr[r] <= 0) Y[e] = T(0);
}

// ---- bias layouts (with_user_item_bias): X / Y carry rank+2 rows, the solve sees rank+1 of them -----------------------
// dst[r][0..k) = src[r][off .. off+k)   (drop_row, wrmf_utils.hpp:3-10, on the device)
template <typename T>
__global__ void pack_cols_kernel(const T* __restrict__ src, int ld, int off, int k, long long n, T* __restrict__ dst) {
  const long long total = n * (long long)k;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / k;
    const int f = (int)(e - r * k);
    dst[e] = src[r * ld + off + f];
  }
}
template <typename T>
__global__ void unpack_cols_kernel(const T* __restrict__ src, int k, long long n, T* __restrict__ dst, int ld, int off) {
  const long long total = n * (long long)k;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / k;
    const int f = (int)(e - r * k);
    dst[r * ld + off + f] = src[e];
  }
}
// partials[b][f] = sum over the block's rows of X[r][f] * ((w ? w[r] : 0) + wadd): the building block of
// rhs_init = -X (x_biases + global_bias) and global_bias_base = -global_bias * sum(X, 1)  (wrmf_implicit.hpp:111-154)
template <typename T>
__global__ void __launch_bounds__(256) weighted_colsum_kernel(const T* __restrict__ X, int k, long long n, const T* __restrict__ w,
                                                              T wadd, double* __restrict__ partials) {
  __shared__ double sh[256];
  const int rpi = max(1, 256 / k);             // rows per iteration of the block
  const int f = threadIdx.x % k, rl = threadIdx.x / k;
  double acc = 0.0;
  if (rl < rpi && threadIdx.x < rpi * k) {
    for (long long r = (long long)blockIdx.x * rpi + rl; r < n; r += (long long)gridDim.x * rpi) {
      const T wr = (w ? w[r] : T(0)) + wadd;
      acc += (double)(X[r * k + f] * wr);
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < k) {
    double t = 0.0;
    for (int q = 0; q < rpi; q++) t += sh[q * k + threadIdx.x];
    partials[(size_t)blockIdx.x * k + threadIdx.x] = t;
  }
}
template <typename T>
__global__ void finish_colsum_kernel(const double* __restrict__ partials, int grid, int k, double scale, T* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= k) return;
  double t = 0.0;
  for (int b = 0; b < grid; b++) t += partials[(size_t)b * k + f];
  out[f] = (T)(scale * t);
}
// sum of squares over columns [lo, hi) of an n x ld matrix, optionally weighted per row (loss regulariser over the
// learned rows only: wrmf_implicit.hpp:286-302, wrmf_explicit.hpp:148-172)
template <typename T>
__global__ void __launch_bounds__(256) sqnorm_cols_kernel(const T* __restrict__ X, int ld, int lo, int hi, long long n,
                                                          const T* __restrict__ cnt, double* __restrict__ partials) {
  __shared__ double s_red[32];
  double acc = 0.0;
  const int w = hi - lo;
  const long long total = n * (long long)w;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / w;
    const double v = (double)X[r * ld + lo + (int)(e - r * w)];
    acc += v * v * (cnt ? (double)cnt[r] : 1.0);
  }
  const double tot = block_sum_double(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// synthetic CSR: row r draws exactly nnz_per_row distinct ascending ids -- one per
// equal-width stratum of [0, n_cols) -- from a counter-based hash; values 1 + floor(10 u^2) (implicit
// confidences) or 1..5 (explicit ratings).
__host__ __device__ inline uint64_t synth_hash(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ inline void synth_entry(int64_t row, int j, int32_t n_cols, int32_t nnz_per_row, uint64_t seed,
                                            int explicit_values, int32_t* col, float* val) {
  const uint64_t h = synth_hash(seed * 0x100000001B3ull + (uint64_t)row * (uint64_t)nnz_per_row + (uint64_t)j);
  const int64_t lo = ((int64_t)j * n_cols) / nnz_per_row, hi = ((int64_t)(j + 1) * n_cols) / nnz_per_row;
  *col = (int32_t)(lo + (int64_t)(h % (uint64_t)(hi - lo)));
  const float u = (float)((h >> 40) & 0xFFFFFF) / 16777216.0f;
  *val = explicit_values ? (1.0f + floorf(u * 5.0f)) : (1.0f + floorf(10.0f * u * u));
}
__global__ void synth_csr_kernel(int32_t n_rows, int32_t n_cols, int32_t nnz_per_row, uint64_t seed, int explicit_values,
                                 int64_t row_offset, int32_t* __restrict__ ptr, int32_t* __restrict__ idx,
                                 float* __restrict__ val) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_rows * nnz_per_row;
  if (e <= n_rows) ptr[e] = (int32_t)(e * nnz_per_row);
  if (e >= total) return;
  const int64_t r = e / nnz_per_row;
  const int j = (int)(e - r * nnz_per_row);
  synth_entry(r + row_offset, j, n_cols, nnz_per_row, seed, explicit_values, &idx[e], &val[e]);
}
// factor init: N(0,1)/100 from a counter-based Box-Muller (R/model_WRMF.R:203-215, src/utils.cpp:131-143)
// `decay` > 0 gives feature f the extra scale (1+f)^-decay: a trained-like, ill-conditioned Gram
__global__ void init_normal_kernel(float* __restrict__ out, long long n, uint64_t seed, float scale, int k = 1,
                                   float decay = 0.f) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (decay != 0.f) scale *= powf(1.0f + (float)(i % k), -decay);
  const uint64_t h1 = synth_hash(seed ^ (uint64_t)(2 * i)), h2 = synth_hash(seed ^ (uint64_t)(2 * i + 1));
  const float u1 = ((float)((h1 >> 40) & 0xFFFFFF) + 1.0f) / 16777217.0f;
  const float u2 = (float)((h2 >> 40) & 0xFFFFFF) / 16777216.0f;
  out[i] = scale * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
