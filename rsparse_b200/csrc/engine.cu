// engine.cu -- host side of libb200als.so: the C ABI of include/b200als.h, device memory, kernel
// selection and the ALS outer loop.  Reference counterparts (relative to /root/reference):
//   stateless calls   src/wrmf_implicit.cpp:5-31, src/wrmf_explicit.cpp:5-27 (+ src/RcppExports.cpp:329-416)
//   XtX               R/model_WRMF.R:474-486
//   session / fit     R/model_WRMF.R:173-360 (outer loop :318-338, final transform_ :412-452)
// No CPU fallback anywhere: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is resolved lazily with dlopen (see NcclApi)

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <limits>
#include <type_traits>
#include <vector>

#include "../../include/b200als.h"
#include "als_chol_tile.cuh"
#include "als_chol_rows.cuh"
#include "als_chol_rows_split.cuh"
#include "als_generic.cuh"
#include "als_resident.cuh"
#include "eig.cuh"
#include "gram.cuh"
#include "gram_tc.cuh"
#include "rotate_tc.cuh"
#include "topk.cuh"
#include "bias_init.cuh"

using namespace b200als;

// ------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(B200ALS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__) + " @" +      \
                                     __FILE__ + ":" + std::to_string(__LINE__));                   \
  } while (0)
#define NC(expr)                                                                                   \
  do {                                                                                             \
    ncclResult_t e__ = (expr);                                                                     \
    if (e__ != ncclSuccess)                                                                        \
      return fail(B200ALS_ENCCL, std::string(#expr) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e__) : "nccl error") + " @" +      \
                                     __FILE__ + ":" + std::to_string(__LINE__));                   \
  } while (0)
#define TRY(expr)                \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != B200ALS_OK) return rc__; \
  } while (0)

static unsigned long long g_launches = 0;  // kernels launched by this library (bench.py reports the delta)
#define LAUNCHED() (++g_launches)

extern "C" const char* b200als_last_error(void) { return g_err.c_str(); }
extern "C" unsigned long long b200als_launch_count(void) { return g_launches; }
extern "C" int b200als_version(void) { return B200ALS_VERSION; }
extern "C" int b200als_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (count) *count = (e == cudaSuccess) ? n : 0;
  if (e != cudaSuccess || n == 0) return fail(B200ALS_ECUDA, "no CUDA device visible (this engine has no CPU fallback)");
  return B200ALS_OK;
}
extern "C" int b200als_set_device(int device) {
  CU(cudaSetDevice(device));
  return B200ALS_OK;
}

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

// ------------------------------------------------------------------------------------------------------
// device context shared by the stateless calls and sessions
// ------------------------------------------------------------------------------------------------------
struct Ctx {
  int device = -1;
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  DevBuf ticket, loss_partials, loss_acc, status, gram_partials, reg_partials, rot_rt;
  bool attrs_set = false;
  int init() {
    if (stream) return B200ALS_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
      return fail(B200ALS_ECUDA, "no CUDA device visible (this engine has no CPU fallback)");
    CU(cudaGetDevice(&device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CU(ticket.ensure(sizeof(unsigned long long)));
    CU(loss_acc.ensure(4 * sizeof(double)));
    CU(status.ensure(sizeof(int)));
    return B200ALS_OK;
  }
};
static Ctx& ctx() {
  static thread_local Ctx c;
  return c;
}

template <typename T>
struct CscDev {
  int32_t n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  DevBuf ptr, idx, val;
  // row classes for the resident kernel (built lazily)
  DevBuf short_list, long_list;
  int n_short = -1, n_long = 0, n_empty = 0;
  bool all_short = false;
};

template <typename T>
static int upload_csc(const b200als_csc* A, CscDev<T>& D, cudaStream_t st) {
  if (!A || !A->ptr || (A->nnz > 0 && (!A->idx || (!A->val_f64 && !A->val_f32))))
    return fail(B200ALS_EINVAL, "b200als_csc: null ptr/idx/val");
  if (A->n_cols < 0 || A->n_rows < 0 || A->nnz < 0) return fail(B200ALS_EINVAL, "b200als_csc: negative size");
  D.n_rows = A->n_rows;
  D.n_cols = A->n_cols;
  D.nnz = A->nnz;
  CU(D.ptr.ensure(sizeof(int32_t) * ((size_t)A->n_cols + 1)));
  CU(D.idx.ensure(sizeof(int32_t) * (size_t)A->nnz));
  CU(D.val.ensure(sizeof(T) * (size_t)A->nnz));
  CU(cudaMemcpyAsync(D.ptr.p, A->ptr, sizeof(int32_t) * ((size_t)A->n_cols + 1), cudaMemcpyHostToDevice, st));
  if (A->nnz) {
    CU(cudaMemcpyAsync(D.idx.p, A->idx, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyHostToDevice, st));
    const bool same_f64 = A->val_f64 && sizeof(T) == 8, same_f32 = !A->val_f64 && sizeof(T) == 4;
    if (same_f64 || same_f32) {
      CU(cudaMemcpyAsync(D.val.p, A->val_f64 ? (const void*)A->val_f64 : (const void*)A->val_f32,
                         sizeof(T) * (size_t)A->nnz, cudaMemcpyHostToDevice, st));
    } else {
      // double -> float (or float -> double) once at upload; the reference converts per visit
      // (wrmf_implicit.hpp:182-183), same rounding
      DevBuf tmp;
      const size_t eb = A->val_f64 ? 8 : 4;
      CU(tmp.ensure(eb * (size_t)A->nnz));
      CU(cudaMemcpyAsync(tmp.p, A->val_f64 ? (const void*)A->val_f64 : (const void*)A->val_f32, eb * (size_t)A->nnz,
                         cudaMemcpyHostToDevice, st));
      const int bs = 256;
      const unsigned gs = (unsigned)((A->nnz + bs - 1) / bs);
      if (A->val_f64) convert_kernel<double, T><<<gs, bs, 0, st>>>(tmp.f64(), D.val.template as<T>(), A->nnz);
      else convert_kernel<float, T><<<gs, bs, 0, st>>>(tmp.f32(), D.val.template as<T>(), A->nnz);
      LAUNCHED(); CU(cudaGetLastError());
      CU(cudaStreamSynchronize(st));
    }
  }
  D.n_short = -1;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Gram
// ------------------------------------------------------------------------------------------------------
// B200ALS_GRAM=ffma forces the fp32 FMA kernel; default at rank 128 / fp32 is the tcgen05 3xTF32 kernel
static bool gram_use_tensor_cores() {
  const char* e = getenv("B200ALS_GRAM");
  return !(e && (e[0] == 'f' || e[0] == 'F'));
}
template <typename T>
static int run_gram(Ctx& c, const T* X, int k, long long n, double lambda, T* G, double* G64) {
  if constexpr (sizeof(T) == 4) {
    if (k == kTcK && n >= 8192 && gram_use_tensor_cores()) {   // small inputs: the exact fp32 FMA kernel
      long long rows_per = std::max<long long>(1024, (n + 887) / 888);
      rows_per = ((rows_per + 255) / 256) * 256;   // whole drain windows
      const long long n_cta = (n + rows_per - 1) / rows_per;
      CU(c.gram_partials.ensure(sizeof(double) * (size_t)n_cta * kTcK * kTcK));
      const size_t smem = sizeof(GramTcSmem);
      CU(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gram_tc_kernel<<<(unsigned)n_cta, 128, smem, c.stream>>>((const float*)X, n, rows_per, c.gram_partials.f64());
      LAUNCHED(); CU(cudaGetLastError());
      gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, c.stream>>>(c.gram_partials.f64(), (int)n_cta, 1, k, lambda, G, G64);
      LAUNCHED(); CU(cudaGetLastError());
      return B200ALS_OK;
    }
  }
  const int nt1 = (k + kGramTile - 1) / kGramTile, n_tiles = nt1 * (nt1 + 1) / 2;
  long long n_cta = std::min<long long>((long long)c.sm_count * 2 / std::max(1, n_tiles) + 1, (n + 255) / 256);
  n_cta = std::max<long long>(1, n_cta);
  long long rows_per = (n + n_cta - 1) / n_cta;
  rows_per = ((rows_per + kGramRows - 1) / kGramRows) * kGramRows;
  n_cta = std::max<long long>(1, (n + rows_per - 1) / rows_per);
  CU(c.gram_partials.ensure(sizeof(double) * (size_t)n_cta * n_tiles * kGramTile * kGramTile));
  gram_partial_kernel<T><<<dim3((unsigned)n_cta, (unsigned)n_tiles), 256, 0, c.stream>>>(X, k, n, rows_per,
                                                                                         c.gram_partials.f64(), nt1);
  LAUNCHED(); CU(cudaGetLastError());
  gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, c.stream>>>(c.gram_partials.f64(), (int)n_cta, n_tiles, k,
                                                                   lambda, G, G64);
  LAUNCHED(); CU(cudaGetLastError());
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// half-iteration dispatch
// ------------------------------------------------------------------------------------------------------
struct HalfOpts {
  int feedback, solver, cg_steps, dynamic_lambda, kernel;
  double lambda;
  int stage = 0;  // tile staging of the resident kernel: 0 default, 1 cp.async.bulk (UBLKCP), 2 cp.async (LDGSTS)
  int ctas = 0;   // resident CTAs per SM the kernel is compiled for: 0 default, 3 or 4
  int row_begin = 0, row_count = -1;  // solve only rows [row_begin, row_begin + row_count) of the block (-1: all)
  bool reset_loss = true;             // zero the loss accumulator first (false: add to it)
  // bias terms, all on compact matrices (see stateless_half): device pointers of the element type being solved
  int with_biases = 0;
  double gbias = 0.0;                 // global_bias after the sqrt(eps) cut-off (wrmf_implicit.hpp:108-109)
  const void* xbias = nullptr;        // [n_src]
  const void* rhs_init = nullptr;     // [k]
  int reg_ld = 0, reg_lo = 0, reg_hi = 0;  // loss regulariser over columns [lo, hi) of the n_src x ld matrix (0: whole matrix)
};
constexpr int kDefaultCtas = 3;
constexpr int kDefaultStage = 1;  // LDGSTS: measured 6 % faster than the UBLKCP variant on C3 (profiles/)

template <typename T>
static int classify_rows(Ctx& c, CscDev<T>& A) {
  if (A.n_short >= 0) return B200ALS_OK;
  CU(A.short_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  CU(A.long_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  DevBuf counts;
  CU(counts.ensure(3 * sizeof(int)));
  CU(cudaMemsetAsync(counts.p, 0, 3 * sizeof(int), c.stream));
  if (A.n_cols > 0) {
    classify_rows_kernel<<<(A.n_cols + 255) / 256, 256, 0, c.stream>>>(A.ptr.i32(), A.n_cols, kResMaxN,
                                                                      A.short_list.i32(), A.long_list.i32(),
                                                                      counts.i32());
    LAUNCHED(); CU(cudaGetLastError());
  }
  int h[3];
  CU(cudaMemcpyAsync(h, counts.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  A.n_short = h[0];
  A.n_long = h[1];
  A.n_empty = h[2];
  A.all_short = (h[0] == A.n_cols);
  return B200ALS_OK;
}

template <typename T, int KPL>
static int launch_cg_generic(Ctx& c, const SolveParams<T>& P, int n_work, int* grid_out) {
  const int grid = (int)std::min<long long>((long long)c.sm_count * 4, std::max(1, (n_work + 7) / 8));
  als_cg_generic_kernel<T, KPL><<<grid, 256, 0, c.stream>>>(P);
  LAUNCHED(); CU(cudaGetLastError());
  *grid_out = grid;
  return B200ALS_OK;
}

// Runs one half-iteration on device data.  `diag`/`rotated`: the caller has put X and Y in the eigenbasis of
// G (implicit CG, rank 128, resident kernel) and passes the eigenvalues.  Accumulates the loss numerator
// (sum over solved rows) into c.loss_acc[0].
template <typename T>
static int solve_rows(Ctx& c, CscDev<T>& A, const T* X, T* Y, const T* G, const float* diag, int k, const HalfOpts& o) {
  if (o.solver != B200ALS_CHOLESKY && o.solver != B200ALS_CONJUGATE_GRADIENT && o.solver != B200ALS_NNLS)
    return fail(B200ALS_EINVAL, "unknown solver code");
  if (o.feedback == B200ALS_IMPLICIT && !G && !diag) return fail(B200ALS_EINVAL, "implicit feedback needs XtX");
  if (o.reset_loss) {
    CU(cudaMemsetAsync(c.loss_acc.p, 0, sizeof(double), c.stream));
    CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
  }
  if (A.n_cols == 0) return B200ALS_OK;
  const bool sub_range = (o.row_count >= 0);
  const int n_rows_here = sub_range ? o.row_count : A.n_cols;
  if (n_rows_here == 0) return B200ALS_OK;
  SolveParams<T> P{};
  const bool biased = o.with_biases || o.gbias != 0.0;
  P.xbias = static_cast<const T*>(o.xbias);
  P.rhs_init = (o.feedback == B200ALS_IMPLICIT) ? static_cast<const T*>(o.rhs_init) : nullptr;
  P.gbias = (T)o.gbias;
  P.one_minus_g = (T)(1 - o.gbias);
  P.solve_empty = (o.feedback == B200ALS_IMPLICIT && biased) ? 1 : 0;
  P.ptr = A.ptr.i32();
  P.idx = A.idx.i32();
  P.val = A.val.template as<T>();
  P.X = X;
  P.Y = Y;
  P.G = (o.feedback == B200ALS_IMPLICIT) ? G : nullptr;
  P.k = k;
  P.n_targets = n_rows_here;
  P.row_begin = sub_range ? o.row_begin : 0;
  P.feedback = o.feedback;
  P.cg_steps = o.cg_steps;
  P.dynamic_lambda = o.dynamic_lambda;
  P.solver = o.solver;
  P.lambda = o.lambda;
  P.row_list = nullptr;
  P.n_list = 0;
  P.n_list_dev = nullptr;
  P.ptr_base = 0;
  P.ticket = c.ticket.u64();
  P.status = c.status.i32();
  const int max_grid = c.sm_count * 8;
  CU(c.loss_partials.ensure(sizeof(double) * (size_t)max_grid));
  P.loss_partials = c.loss_partials.f64();

  auto run_generic_cg = [&](const int32_t* list, int n_list) -> int {
    P.row_list = list;
    P.n_list = n_list;
    const int n_work = list ? n_list : n_rows_here;
    if (n_work == 0) return B200ALS_OK;
    CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
    int grid = 0;
    if (k <= 32) TRY((launch_cg_generic<T, 1>(c, P, n_work, &grid)));
    else if (k <= 64) TRY((launch_cg_generic<T, 2>(c, P, n_work, &grid)));
    else if (k <= 128) TRY((launch_cg_generic<T, 4>(c, P, n_work, &grid)));
    else if (k <= 256) TRY((launch_cg_generic<T, 8>(c, P, n_work, &grid)));
    else return fail(B200ALS_EUNSUPPORTED, "rank > 256 is not supported");
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  };

  if (o.solver == B200ALS_CHOLESKY || o.solver == B200ALS_NNLS) {
    auto run_generic_chol = [&](const int32_t* list, int n_list) -> int {
      P.row_list = list;
      P.n_list = n_list;
      const int n_work = list ? n_list : n_rows_here;
      if (n_work == 0) return B200ALS_OK;
      const size_t smem = chol_generic_smem_bytes<T>(k, o.solver);
      if (smem > c.smem_optin)
        return fail(B200ALS_EUNSUPPORTED, "cholesky / nnls: rank too large for the shared-memory factorisation (needs " +
                                             std::to_string(smem) + " B)");
      CU(cudaFuncSetAttribute(als_chol_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024)));
      const int grid = std::min(c.sm_count * per_sm, std::max(1, n_work));
      CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
      als_chol_generic_kernel<T><<<grid, 256, smem, c.stream>>>(P);
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
      return B200ALS_OK;
    };
    bool tiled = false;
    if constexpr (sizeof(T) == 4)
      tiled = (o.solver == B200ALS_CHOLESKY) && (k == 64 || k == 128) && o.kernel != 1 && !sub_range && !biased;
    if (!tiled) return run_generic_chol(nullptr, 0);
    if constexpr (sizeof(T) == 4) {
      // rows with 1..80 non-zeros: row-per-thread (or tile) kernel; longer rows: generic kernel; empty rows: zero
      TRY(classify_rows(c, A));
      if (A.n_empty > 0) {
        zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_short > 0) {
        P.row_list = A.all_short ? nullptr : A.short_list.i32();
        P.n_list = A.n_short;
        // default (and kernel = 4): row-per-thread panel kernel (als_chol_rows.cuh), measured 2.0x (rank 64) / 1.6x
        // (rank 128) faster than its predecessor, the 16 x 16 register-block kernel, which stays selectable as kernel = 5
        const bool rows_kernel = (o.kernel != 5);
        // persistent CTAs: exactly as many as are co-resident (registers AND shared memory), else a second wave
        int per_sm = 1, grid = 1;
        auto launch = [&](auto kern, int threads, size_t smem) -> cudaError_t {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
          if (e != cudaSuccess) return e;
          grid = std::min(c.sm_count * std::max(1, per_sm), A.n_short);
          kern<<<grid, threads, smem, c.stream>>>(P);
          return cudaSuccess;
        };
        if (rows_kernel) {
          if (k == 64) CU(launch(als_chol_rows_kernel<64, 8>, 64, sizeof(CholRowsSmem<64>)));
          else if (o.kernel == 6) CU(launch(als_chol_rows_kernel<128, 3, 1>, 128, sizeof(CholRowsSmem<128>)));   // tcgen05 Gram, single-buffered (experimental)
          else if (o.kernel == 8) CU(launch(als_chol_rows_split_kernel, kSplitThreads, sizeof(CholRowsSmem<128>)));   // split rows (experimental, not yet run on a GPU)
          else if (o.kernel == 7) CU(launch(als_chol_rows_kernel<128, 3, 2>, 128, sizeof(CholRowsSmem<128>)));   // tcgen05 Gram, pipelined (experimental, not yet run on a GPU)
          else if (o.ctas == 2) CU(launch(als_chol_rows_kernel<128, 2>, 128, sizeof(CholRowsSmem<128>)));
          else CU(launch(als_chol_rows_kernel<128, 3>, 128, sizeof(CholRowsSmem<128>)));   // measured: 134.5 vs 171.2 ms / 1 M rows
        } else {
          if (k == 64) CU(launch(als_chol_tile_kernel<64>, kCholThreads, sizeof(CholTileSmem<64>)));
          else CU(launch(als_chol_tile_kernel<128>, kCholThreads, sizeof(CholTileSmem<128>)));
        }
        LAUNCHED(); CU(cudaGetLastError());
        sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_long > 0) TRY(run_generic_chol(A.long_list.i32(), A.n_long));
    }
    return B200ALS_OK;
  }

  // ---- conjugate gradient ----
  bool resident = false;
  if constexpr (sizeof(T) == 4) {
    resident = (k == kResK) && (o.kernel != 1) && (o.feedback == B200ALS_EXPLICIT || G || diag) && !biased;
    if (o.kernel == 2 && !resident) return fail(B200ALS_EUNSUPPORTED, "resident kernel requires rank 128 fp32");
  }
  if (!resident) {
    if (diag && !G) return fail(B200ALS_EINVAL, "generic CG needs the full XtX");
    return run_generic_cg(nullptr, 0);
  }
  if constexpr (sizeof(T) == 4) {
    TRY(classify_rows(c, A));
    if (sub_range && !A.all_short) return fail(B200ALS_EINVAL, "row sub-ranges need a block without empty or long rows");
    if (A.n_empty > 0) {
      zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (A.n_short > 0) {
      ResidentParams R;
      R.ptr = P.ptr;
      R.idx = P.idx;
      R.val = (const float*)P.val;
      R.X = (const float*)X;
      R.Y = (float*)Y;
      R.diag = diag;
      R.G = (const float*)G;
      R.feedback = o.feedback;
      R.cg_steps = o.cg_steps;
      R.dynamic_lambda = o.dynamic_lambda;
      R.lambda = (float)o.lambda;
      R.row_list = A.all_short ? nullptr : A.short_list.i32();
      R.n_list = sub_range ? n_rows_here : A.n_short;
      R.n_list_dev = nullptr;
      R.ptr_base = 0;
      R.row_begin = sub_range ? o.row_begin : 0;
      R.loss_partials = P.loss_partials;
      const int ctas = (o.ctas == 3 || o.ctas == 4) ? o.ctas : kDefaultCtas;
      const int grid = std::min(c.sm_count * ctas, R.n_list);
      const size_t smem = sizeof(ResidentSmem);
      const bool full_g = (o.feedback == B200ALS_IMPLICIT) && !diag;
      auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kResThreads, smem, c.stream>>>(R);
        return cudaSuccess;
      };
      const int stage = (o.stage == 1) ? 0 : (o.stage == 2 ? 1 : kDefaultStage);
      if (full_g) {
        if (ctas == 4) CU(stage == 0 ? launch(als_cg_resident_kernel<true, 0, 4>) : launch(als_cg_resident_kernel<true, 1, 4>));
        else CU(stage == 0 ? launch(als_cg_resident_kernel<true, 0, 3>) : launch(als_cg_resident_kernel<true, 1, 3>));
      } else {
        if (ctas == 4) CU(stage == 0 ? launch(als_cg_resident_kernel<false, 0, 4>) : launch(als_cg_resident_kernel<false, 1, 4>));
        else CU(stage == 0 ? launch(als_cg_resident_kernel<false, 0, 3>) : launch(als_cg_resident_kernel<false, 1, 3>));
      }
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (A.n_long > 0) {
      if (diag && !G) return fail(B200ALS_EINVAL, "rows longer than 80 need the full XtX for the streaming kernel");
      TRY(run_generic_cg(A.long_list.i32(), A.n_long));
    }
  }
  return B200ALS_OK;
}

// loss = (sum_rows + lambda * regulariser) / nnz, rounded through T like the reference's return type
// (wrmf_implicit.hpp:286-304, wrmf_explicit.hpp:147-173)
template <typename T>
static int finish_loss(Ctx& c, const T* X, int k, long long n_src, const T* cnt_X, const HalfOpts& o, int64_t nnz,
                       double rows_sum, bool rows_sum_given, double* loss_out) {
  double reg = 0.0;
  if (o.lambda > 0) {
    const bool weighted = (o.feedback == B200ALS_EXPLICIT) && o.dynamic_lambda;
    const int grid = c.sm_count * 2;
    CU(c.reg_partials.ensure(sizeof(double) * (size_t)grid));
    if (o.reg_ld > 0)
      sqnorm_cols_kernel<T><<<grid, 256, 0, c.stream>>>(X, o.reg_ld, o.reg_lo, o.reg_hi, n_src, weighted ? cnt_X : nullptr,
                                                       c.reg_partials.f64());
    else
      sqnorm_kernel<T><<<grid, 256, 0, c.stream>>>(X, k, n_src, weighted ? cnt_X : nullptr, c.reg_partials.f64());
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.reg_partials.f64(), grid, c.loss_acc.f64() + 1, 0);
    LAUNCHED(); CU(cudaGetLastError());
  }
  double h[2] = {0, 0};
  int st = 0;
  CU(cudaMemcpyAsync(h, c.loss_acc.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(&st, c.status.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  if (st != 0) return fail(B200ALS_ENOTSPD, "a per-row system was not positive definite (Cholesky pivot <= 0)");
  if (o.lambda > 0) reg = h[1];
  const double rows = rows_sum_given ? rows_sum : h[0];
  if (loss_out) *loss_out = (double)(T)((rows + o.lambda * reg) / (double)nnz);
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// 1. stateless calls
// ------------------------------------------------------------------------------------------------------
// Bias arguments of the reference entry points (src/wrmf_implicit.cpp:5-31, src/wrmf_explicit.cpp:5-27).
template <typename T>
struct BiasArgs {
  int with_biases = 0, is_x_bias_last_row = 0;
  double global_bias = 0.0;
  T* global_bias_base = nullptr;   // host, [rank - with_biases], in/out
  int initialize_bias_base = 0;
};

// `rank` = rows of X and Y as the caller holds them (R's private$rank: rank + 2 with biases, model_WRMF.R:162-166).
// With biases the reference solves a (rank-1)-sized system on row-dropped views (drop_row, wrmf_utils.hpp:3-10):
//   is_x_bias_last_row:  X = [1, ..., x_bias]   Y = [y_bias, ..., 1]     X_nnz = X rows 0..rank-2, x_biases = last row
//   otherwise:           X = [x_bias, ..., 1]   Y = [1, ..., y_bias]     X_nnz = X rows 1..rank-1, x_biases = first row
// Here the views are materialised once on the device as compact matrices Xc (n_src x k), xb (n_src), Yc (n_tgt x k),
// k = rank - 1, the generic kernels run on those, and the solved rows are scattered back into Y.
template <typename T>
static int stateless_half(const b200als_csc* A, int rank, const T* X, T* Y, const T* XtX, const T* cnt_X, HalfOpts o,
                          double* loss, const BiasArgs<T>& ba = BiasArgs<T>()) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!A || !X || !Y) return fail(B200ALS_EINVAL, "null argument");
  if (rank <= 0) return fail(B200ALS_EINVAL, "rank must be positive");
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  const bool wb = ba.with_biases != 0, is_last = ba.is_x_bias_last_row != 0;
  double gbias = implicit ? ba.global_bias : 0.0;
  if (gbias < std::sqrt((double)std::numeric_limits<T>::epsilon())) gbias = 0.0;          // wrmf_implicit.hpp:108-109
  if (wb && rank < 2) return fail(B200ALS_EINVAL, "with_biases needs at least 2 rows in X / Y");
  if (!wb && gbias != 0.0 && !ba.global_bias_base) return fail(B200ALS_EINVAL, "global_bias needs global_bias_base");
  const int ks = wb ? rank - 1 : rank;           // size of the solved system
  const int xo = (wb && !is_last) ? 1 : 0;       // X_nnz = drop_row(X_nnz, is_x_bias_last_row)          (:190 / :88)
  const int xbcol = is_last ? rank - 1 : 0;      // x_biases                                             (:115-119)
  const int io = (wb && is_last) ? 1 : 0;        // init = drop_row(init, !is_x_bias_last_row), sic      (:191 / :90)
  const int oo = (wb && !is_last) ? 1 : 0;       // Y.head(rank-1) / Y.tail(rank-1)                      (:240-252)
  CscDev<T> D;
  TRY(upload_csc<T>(A, D, c.stream));
  const size_t k = (size_t)rank;
  const size_t n_src = (size_t)A->n_rows, n_tgt = (size_t)A->n_cols;
  DevBuf dX, dY, dG, dCnt, dXc, dYc, dXb, dRhs;
  CU(dX.ensure(sizeof(T) * k * n_src));
  CU(dY.ensure(sizeof(T) * k * n_tgt));
  CU(cudaMemcpyAsync(dX.p, X, sizeof(T) * k * n_src, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(dY.p, Y, sizeof(T) * k * n_tgt, cudaMemcpyHostToDevice, c.stream));
  const T* Xs = dX.template as<T>();   // what the kernels gather from
  T* Ys = dY.template as<T>();         // what they solve in place
  const int cp_grid = c.sm_count * 8;
  if (wb) {
    CU(dXc.ensure(sizeof(T) * (size_t)ks * std::max<size_t>(1, n_src)));
    CU(dYc.ensure(sizeof(T) * (size_t)ks * std::max<size_t>(1, n_tgt)));
    CU(dXb.ensure(sizeof(T) * std::max<size_t>(1, n_src)));
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dX.template as<T>(), rank, xo, ks, (long long)n_src, dXc.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dX.template as<T>(), rank, xbcol, 1, (long long)n_src, dXb.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dY.template as<T>(), rank, io, ks, (long long)n_tgt, dYc.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    Xs = dXc.template as<T>();
    Ys = dYc.template as<T>();
    o.with_biases = 1;
    o.xbias = dXb.p;
    o.reg_ld = rank;                     // every learned row of X: all but the row of ones (:286-302 / :148-172)
    o.reg_lo = is_last ? 1 : 0;
    o.reg_hi = is_last ? rank : rank - 1;
  }
  o.gbias = gbias;
  const T* G = nullptr;
  if (implicit) {
    CU(dG.ensure(sizeof(T) * (size_t)ks * ks));
    if (XtX) CU(cudaMemcpyAsync(dG.p, XtX, sizeof(T) * (size_t)ks * ks, cudaMemcpyHostToDevice, c.stream));
    else TRY(run_gram<T>(c, Xs, ks, A->n_rows, o.lambda, dG.template as<T>(), nullptr));   // R/model_WRMF.R:474-486
    G = dG.template as<T>();
    if (wb || gbias != 0.0) {
      // rhs_init = -X_nnz-view * (x_biases + global_bias) (:143-154) ; global_bias_base = sum(X, 1) * (-global_bias) (:111-112)
      CU(dRhs.ensure(sizeof(T) * (size_t)ks));
      const bool compute = wb || ba.initialize_bias_base;
      if (compute) {
        if (ks > 256) return fail(B200ALS_EUNSUPPORTED, "bias terms: rank > 256 is not supported");
        const int cs_grid = c.sm_count * 4;
        CU(c.reg_partials.ensure(sizeof(double) * (size_t)cs_grid * ks));
        weighted_colsum_kernel<T><<<cs_grid, 256, 0, c.stream>>>(Xs, ks, (long long)n_src, wb ? dXb.template as<T>() : nullptr,
                                                               wb ? (T)gbias : T(1), c.reg_partials.f64());
        LAUNCHED(); CU(cudaGetLastError());
        finish_colsum_kernel<T><<<(ks + 127) / 128, 128, 0, c.stream>>>(c.reg_partials.f64(), cs_grid, ks, wb ? -1.0 : -gbias,
                                                                       dRhs.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
        if (!wb) CU(cudaMemcpyAsync(ba.global_bias_base, dRhs.p, sizeof(T) * (size_t)ks, cudaMemcpyDeviceToHost, c.stream));
      } else {
        CU(cudaMemcpyAsync(dRhs.p, ba.global_bias_base, sizeof(T) * (size_t)ks, cudaMemcpyHostToDevice, c.stream));
      }
      o.rhs_init = dRhs.p;
    }
  }
  const T* dcnt = nullptr;
  if (o.feedback == B200ALS_EXPLICIT && o.dynamic_lambda && o.lambda > 0) {
    if (!cnt_X) return fail(B200ALS_EINVAL, "explicit feedback with dynamic_lambda needs cnt_X");
    CU(dCnt.ensure(sizeof(T) * n_src));
    CU(cudaMemcpyAsync(dCnt.p, cnt_X, sizeof(T) * n_src, cudaMemcpyHostToDevice, c.stream));
    dcnt = dCnt.template as<T>();
  }
  TRY(solve_rows<T>(c, D, Xs, Ys, G, nullptr, ks, o));
  if (wb) {
    unpack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dYc.template as<T>(), ks, (long long)n_tgt, dY.template as<T>(), rank, oo);
    LAUNCHED(); CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(Y, dY.p, sizeof(T) * k * n_tgt, cudaMemcpyDeviceToHost, c.stream));
  TRY(finish_loss<T>(c, dX.template as<T>(), rank, A->n_rows, dcnt, o, A->nnz, 0.0, false, loss));
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// 1b. pipelined stateless call: fp32, CG, rank 128, large inputs.  The solved rows are cut into blocks of
//     <= 512k rows / 64M non-zeros; block c+1 and c+2 travel host->device (copy engine) while block c is
//     classified, rotated, solved and rotated back on the compute stream and block c-1 returns device->host.
//     Device buffers are cached in the context between calls; no data is retained.
// ------------------------------------------------------------------------------------------------------
__global__ void diag_matrix_kernel(const float* __restrict__ d, float* __restrict__ G, int k) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < k * k) G[e] = ((e / k) == (e % k)) ? d[e / k] : 0.f;
}
struct PipeBuf {
  DevBuf ptr, idx, val64, val32, Y, short_list, long_list, counts;
  cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
  bool used = false;
};
struct PipeCtx {
  static constexpr int NB = 3;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  PipeBuf buf[NB];
  DevBuf X, G, G64, Vt, Q, Qt, Q64, diag, Gdiag, cnt;
  int init() {
    if (h2d) return B200ALS_OK;
    CU(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    for (auto& b : buf) {
      CU(cudaEventCreateWithFlags(&b.h2d_done, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.compute_done, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.d2h_done, cudaEventDisableTiming));
    }
    return B200ALS_OK;
  }
};
static PipeCtx& pipe_ctx() {
  static thread_local PipeCtx p;
  return p;
}
static int rotate_matrix(Ctx& c, float* M, long long n, const float* R);

static int stateless_pipelined(const b200als_csc* A, const float* X, float* Y, const float* XtX, const float* cnt_X,
                               const HalfOpts& o, double* loss) {
  Ctx& c = ctx();
  PipeCtx& pc = pipe_ctx();
  TRY(pc.init());
  const int k = kResK;
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  const int32_t* hp = A->ptr;
  // ---- block boundaries from the host row pointers ----
  int64_t kMaxRows = 512 * 1024;
  const int64_t kMaxNnz = 64ll * 1024 * 1024;
  if (const char* er = getenv("B200ALS_PIPELINE_ROWS")) kMaxRows = std::max<int64_t>(1, atoll(er));  // tests: force many blocks
  std::vector<int32_t> cuts{0};
  int64_t max_rows = 0, max_nnz = 0;
  while (cuts.back() < A->n_cols) {
    const int32_t b = cuts.back();
    int32_t e = (int32_t)std::min<int64_t>(A->n_cols, (int64_t)b + kMaxRows);
    while (e > b + 1 && (int64_t)hp[e] - hp[b] > kMaxNnz) e = b + std::max(1, (e - b) / 2);
    cuts.push_back(e);
    max_rows = std::max<int64_t>(max_rows, e - b);
    max_nnz = std::max<int64_t>(max_nnz, (int64_t)hp[e] - hp[b]);
  }
  const int n_chunks = (int)cuts.size() - 1;
  for (auto& b : pc.buf) {
    CU(b.ptr.ensure(sizeof(int32_t) * (size_t)(max_rows + 1)));
    CU(b.idx.ensure(sizeof(int32_t) * (size_t)max_nnz));
    if (A->val_f64) CU(b.val64.ensure(sizeof(double) * (size_t)max_nnz));
    CU(b.val32.ensure(sizeof(float) * (size_t)max_nnz));
    CU(b.Y.ensure(sizeof(float) * (size_t)max_rows * k));
    CU(b.short_list.ensure(sizeof(int32_t) * (size_t)max_rows));
    CU(b.long_list.ensure(sizeof(int32_t) * (size_t)max_rows));
    CU(b.counts.ensure(4 * sizeof(int)));
    b.used = false;
  }
  // ---- fixed matrix, Gram, eigenbasis (compute stream) ----
  const size_t xbytes = sizeof(float) * (size_t)k * (size_t)A->n_rows;
  CU(pc.X.ensure(xbytes));
  CU(cudaMemcpyAsync(pc.X.p, X, xbytes, cudaMemcpyHostToDevice, c.stream));
  const float* diag = nullptr;
  const float* Glong = nullptr;
  if (implicit) {
    CU(pc.G.ensure(sizeof(float) * k * k));
    CU(pc.G64.ensure(sizeof(double) * k * k));
    CU(pc.Vt.ensure(sizeof(double) * k * k));
    CU(pc.Q64.ensure(sizeof(double) * k * k));
    CU(pc.Q.ensure(sizeof(float) * k * k));
    CU(pc.Qt.ensure(sizeof(float) * k * k));
    CU(pc.diag.ensure(sizeof(float) * k));
    CU(pc.Gdiag.ensure(sizeof(float) * k * k));
    if (XtX) {
      CU(cudaMemcpyAsync(pc.G.p, XtX, sizeof(float) * k * k, cudaMemcpyHostToDevice, c.stream));
      convert_kernel<float, double><<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.G.f32(), pc.G64.f64(), k * k);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      TRY(run_gram<float>(c, pc.X.f32(), k, A->n_rows, o.lambda, pc.G.f32(), pc.G64.f64()));
    }
    const size_t jsm = sizeof(double) * (size_t)k * (k + 1);
    CU(cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
    jacobi_eig_kernel<<<1, kJacobiThreads, jsm, c.stream>>>(pc.G64.f64(), pc.Vt.f64(), k, pc.Q.f32(), pc.diag.f32(),
                                                            pc.Q64.f64(), 30, 1);
    LAUNCHED(); CU(cudaGetLastError());
    convert_kk_kernel<<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.Q64.f64(), pc.Qt.f32(), k, 1);
    LAUNCHED(); CU(cudaGetLastError());
    diag_matrix_kernel<<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.diag.f32(), pc.Gdiag.f32(), k);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, pc.X.f32(), A->n_rows, pc.Q.f32()));
    diag = pc.diag.f32();
    Glong = pc.Gdiag.f32();
  }
  const float* dcnt = nullptr;
  if (!implicit && o.dynamic_lambda && o.lambda > 0) {
    if (!cnt_X) return fail(B200ALS_EINVAL, "explicit feedback with dynamic_lambda needs cnt_X");
    CU(pc.cnt.ensure(sizeof(float) * (size_t)A->n_rows));
    CU(cudaMemcpyAsync(pc.cnt.p, cnt_X, sizeof(float) * (size_t)A->n_rows, cudaMemcpyHostToDevice, c.stream));
    dcnt = pc.cnt.f32();
  }
  CU(cudaMemsetAsync(c.loss_acc.p, 0, sizeof(double), c.stream));
  CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
  const int res_grid = c.sm_count * 3;
  const int gen_grid = c.sm_count * 4;
  CU(c.loss_partials.ensure(sizeof(double) * (size_t)c.sm_count * 8));
  const size_t res_smem = sizeof(ResidentSmem);
  CU(cudaFuncSetAttribute(als_cg_resident_kernel<false, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
  // ---- the pipeline ----
  for (int ci = 0; ci < n_chunks; ci++) {
    PipeBuf& b = pc.buf[ci % PipeCtx::NB];
    const int32_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
    const int64_t e0 = hp[r0], ne = (int64_t)hp[r1] - e0;
    // host -> device
    if (b.used) CU(cudaStreamWaitEvent(pc.h2d, b.d2h_done, 0));
    CU(cudaMemcpyAsync(b.ptr.p, hp + r0, sizeof(int32_t) * (size_t)(nr + 1), cudaMemcpyHostToDevice, pc.h2d));
    if (ne) {
      CU(cudaMemcpyAsync(b.idx.p, A->idx + e0, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
      if (A->val_f64) CU(cudaMemcpyAsync(b.val64.p, A->val_f64 + e0, sizeof(double) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
      else CU(cudaMemcpyAsync(b.val32.p, A->val_f32 + e0, sizeof(float) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
    }
    CU(cudaMemcpyAsync(b.Y.p, Y + (size_t)r0 * k, sizeof(float) * (size_t)nr * k, cudaMemcpyHostToDevice, pc.h2d));
    CU(cudaEventRecord(b.h2d_done, pc.h2d));
    // compute
    CU(cudaStreamWaitEvent(c.stream, b.h2d_done, 0));
    if (A->val_f64 && ne) {
      convert_kernel<double, float><<<(unsigned)((ne + 255) / 256), 256, 0, c.stream>>>(b.val64.f64(), b.val32.f32(), ne);
      LAUNCHED(); CU(cudaGetLastError());
    }
    CU(cudaMemsetAsync(b.counts.p, 0, 4 * sizeof(int), c.stream));
    classify_rows_kernel<<<(nr + 255) / 256, 256, 0, c.stream>>>(b.ptr.i32(), nr, kResMaxN, b.short_list.i32(), b.long_list.i32(), b.counts.i32());
    LAUNCHED(); CU(cudaGetLastError());
    zero_empty_rows_kernel<float><<<(unsigned)(((long long)nr * k + 255) / 256), 256, 0, c.stream>>>(b.ptr.i32(), nr, k, b.Y.f32());
    LAUNCHED(); CU(cudaGetLastError());
    if (implicit) TRY(rotate_matrix(c, b.Y.f32(), nr, pc.Q.f32()));
    ResidentParams R;
    R.ptr = b.ptr.i32(); R.idx = b.idx.i32(); R.val = b.val32.f32();
    R.X = pc.X.f32(); R.Y = b.Y.f32(); R.diag = diag; R.G = nullptr;
    R.feedback = o.feedback; R.cg_steps = o.cg_steps; R.dynamic_lambda = o.dynamic_lambda; R.lambda = (float)o.lambda;
    R.row_list = b.short_list.i32(); R.n_list = 0; R.n_list_dev = b.counts.i32(); R.ptr_base = (int)e0; R.row_begin = 0;
    R.loss_partials = c.loss_partials.f64();
    als_cg_resident_kernel<false, 1, 3><<<res_grid, kResThreads, res_smem, c.stream>>>(R);
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.loss_partials.f64(), res_grid, c.loss_acc.f64(), 1);
    LAUNCHED(); CU(cudaGetLastError());
    {  // rows longer than the register tile: streaming kernel on the same (rotated) data
      SolveParams<float> P{};
      P.one_minus_g = 1.f;
      P.ptr = b.ptr.i32(); P.idx = b.idx.i32(); P.val = b.val32.f32(); P.X = pc.X.f32(); P.Y = b.Y.f32();
      P.G = implicit ? Glong : nullptr; P.k = k; P.n_targets = nr; P.feedback = o.feedback; P.cg_steps = o.cg_steps;
      P.dynamic_lambda = o.dynamic_lambda; P.solver = 0; P.lambda = o.lambda; P.row_list = b.long_list.i32(); P.n_list = 0;
      P.n_list_dev = b.counts.i32() + 1; P.ptr_base = (int)e0; P.row_begin = 0; P.ticket = c.ticket.u64();
      P.loss_partials = c.loss_partials.f64(); P.status = c.status.i32();
      CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
      als_cg_generic_kernel<float, 4><<<gen_grid, 256, 0, c.stream>>>(P);
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.loss_partials.f64(), gen_grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (implicit) TRY(rotate_matrix(c, b.Y.f32(), nr, pc.Qt.f32()));
    CU(cudaEventRecord(b.compute_done, c.stream));
    // device -> host
    CU(cudaStreamWaitEvent(pc.d2h, b.compute_done, 0));
    CU(cudaMemcpyAsync(Y + (size_t)r0 * k, b.Y.p, sizeof(float) * (size_t)nr * k, cudaMemcpyDeviceToHost, pc.d2h));
    CU(cudaEventRecord(b.d2h_done, pc.d2h));
    b.used = true;
  }
  TRY(finish_loss<float>(c, pc.X.f32(), k, A->n_rows, dcnt, o, A->nnz, 0.0, false, loss));
  CU(cudaStreamSynchronize(pc.d2h));
  CU(cudaStreamSynchronize(pc.h2d));
  return B200ALS_OK;
}

// large fp32 CG problems at rank 128 take the pipelined path (B200ALS_PIPELINE=0 disables, =1 forces)
static bool use_pipelined(const b200als_csc* m, int rank, const float* X, const float* Y, const HalfOpts& o) {
  if (!m || !X || !Y || !m->ptr || rank != kResK || o.solver != B200ALS_CONJUGATE_GRADIENT) return false;
  if (ctx().init() != B200ALS_OK) return false;
  const char* env = getenv("B200ALS_PIPELINE");
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return m->n_cols > 0;
  return m->n_cols >= 200000;
}

// bias terms run on the generic kernels of the plain (non-pipelined) call
extern "C" int b200als_als_implicit_float(const b200als_csc* m, int rank, const float* X, float* Y, const float* XtX,
                                          double lambda, int, unsigned solver, unsigned cg_steps, int with_biases,
                                          int is_x_bias_last_row, double global_bias, float* global_bias_base,
                                          int initialize_bias_base, double* loss) {
  HalfOpts o{B200ALS_IMPLICIT, (int)solver, (int)cg_steps, 0, 0, lambda};
  const bool biased = with_biases || global_bias >= std::sqrt((double)std::numeric_limits<float>::epsilon());
  if (!biased && use_pipelined(m, rank, X, Y, o)) return stateless_pipelined(m, X, Y, XtX, nullptr, o, loss);
  BiasArgs<float> ba{with_biases, is_x_bias_last_row, global_bias, global_bias_base, initialize_bias_base};
  return stateless_half<float>(m, rank, X, Y, XtX, nullptr, o, loss, ba);
}
extern "C" int b200als_als_implicit_double(const b200als_csc* m, int rank, const double* X, double* Y, const double* XtX,
                                           double lambda, int, unsigned solver, unsigned cg_steps, int with_biases,
                                           int is_x_bias_last_row, double global_bias, double* global_bias_base,
                                           int initialize_bias_base, double* loss) {
  HalfOpts o{B200ALS_IMPLICIT, (int)solver, (int)cg_steps, 0, 0, lambda};
  BiasArgs<double> ba{with_biases, is_x_bias_last_row, global_bias, global_bias_base, initialize_bias_base};
  return stateless_half<double>(m, rank, X, Y, XtX, nullptr, o, loss, ba);
}
extern "C" int b200als_als_explicit_float(const b200als_csc* m, int rank, const float* X, float* Y, const float* cnt_X,
                                          double lambda, unsigned, unsigned solver, unsigned cg_steps, int dynamic_lambda,
                                          int with_biases, int is_x_bias_last_row, double* loss) {
  HalfOpts o{B200ALS_EXPLICIT, (int)solver, (int)cg_steps, dynamic_lambda != 0, 0, lambda};
  if (!with_biases && use_pipelined(m, rank, X, Y, o)) return stateless_pipelined(m, X, Y, nullptr, cnt_X, o, loss);
  BiasArgs<float> ba{with_biases, is_x_bias_last_row, 0.0, nullptr, 0};
  return stateless_half<float>(m, rank, X, Y, nullptr, cnt_X, o, loss, ba);
}
extern "C" int b200als_als_explicit_double(const b200als_csc* m, int rank, const double* X, double* Y, const double* cnt_X,
                                           double lambda, unsigned, unsigned solver, unsigned cg_steps, int dynamic_lambda,
                                           int with_biases, int is_x_bias_last_row, double* loss) {
  HalfOpts o{B200ALS_EXPLICIT, (int)solver, (int)cg_steps, dynamic_lambda != 0, 0, lambda};
  BiasArgs<double> ba{with_biases, is_x_bias_last_row, 0.0, nullptr, 0};
  return stateless_half<double>(m, rank, X, Y, nullptr, cnt_X, o, loss, ba);
}

// initialize_biases<T> (wrmf_utils.hpp:170-183; src/wrmf_init.cpp:6-34) -- see bias_init.cuh
template <typename T>
static int initialize_biases_impl(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr, const int32_t* csc_idx,
                                  double* csc_val, const int32_t* csr_ptr, const int32_t* csr_idx, double* csr_val,
                                  T* user_bias, T* item_bias, double lambda, int dynamic_lambda, int non_negative,
                                  int calculate_global_bias, int is_explicit, double* global_bias) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!csc_ptr || !csr_ptr || !user_bias || !item_bias || n_user < 0 || n_item < 0 || nnz < 0)
    return fail(B200ALS_EINVAL, "bad argument");
  if (nnz > 0 && (!csc_idx || !csc_val || !csr_idx || !csr_val)) return fail(B200ALS_EINVAL, "null matrix slots");
  DevBuf cp, ci, cv, rp, ri, rv, ub, ib, part, scal, um, ua, im, ia;
  const size_t e = (size_t)std::max<int64_t>(1, nnz);
  CU(cp.ensure(sizeof(int32_t) * ((size_t)n_item + 1)));
  CU(rp.ensure(sizeof(int32_t) * ((size_t)n_user + 1)));
  CU(ci.ensure(sizeof(int32_t) * e)); CU(ri.ensure(sizeof(int32_t) * e));
  CU(cv.ensure(sizeof(double) * e)); CU(rv.ensure(sizeof(double) * e));
  CU(ub.ensure(sizeof(T) * (size_t)std::max(1, n_user)));
  CU(ib.ensure(sizeof(T) * (size_t)std::max(1, n_item)));
  const int grid = c.sm_count * 4;
  CU(part.ensure(sizeof(double) * (size_t)grid));
  CU(scal.ensure(sizeof(double) * 4));
  cudaStream_t st = c.stream;
  CU(cudaMemcpyAsync(cp.p, csc_ptr, sizeof(int32_t) * ((size_t)n_item + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(rp.p, csr_ptr, sizeof(int32_t) * ((size_t)n_user + 1), cudaMemcpyHostToDevice, st));
  if (nnz) {
    CU(cudaMemcpyAsync(ci.p, csc_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ri.p, csr_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cv.p, csc_val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(rv.p, csr_val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  }
  if (n_user) CU(cudaMemcpyAsync(ub.p, user_bias, sizeof(T) * (size_t)n_user, cudaMemcpyHostToDevice, st));
  if (n_item) CU(cudaMemcpyAsync(ib.p, item_bias, sizeof(T) * (size_t)n_item, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(scal.p, 0, sizeof(double) * 4, st));
  double* d_scal = scal.f64();   // [0] sum of values, [1] sum(user_bias), [2] sum(item_bias)
  auto device_sum = [&](auto* v, long long n, double* out) -> int {
    using V = std::remove_pointer_t<decltype(v)>;
    sum_to_partials_kernel<std::remove_const_t<V>><<<grid, 256, 0, st>>>(v, n, part.f64());
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, st>>>(part.f64(), grid, out, 0);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  };
  double g = 0.0;
  const unsigned gi = (unsigned)std::max(1, (n_item + 127) / 128), gu = (unsigned)std::max(1, (n_user + 127) / 128);
  if (calculate_global_bias && nnz > 0) {
    TRY(device_sum((const double*)cv.f64(), (long long)nnz, d_scal));
    double s = 0.0;
    CU(cudaMemcpyAsync(&s, d_scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (is_explicit) {
      g = s / (double)nnz;                                                     // mean rating (wrmf_utils.hpp:40-43)
      shift_values_kernel<<<grid, 256, 0, st>>>(cv.f64(), rv.f64(), (long long)nnz, d_scal, 1.0 / (double)nnz);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      g = s / (s + (double)n_item * (double)n_user - (double)nnz);             // (:91-94)
    }
  }
  if (!is_explicit && non_negative) g = std::fmax(0.0, g);                       // (:95)
  if (is_explicit) {
    for (int iter = 0; iter < 5; iter++) {                                      // (:54-80)
      if (n_item) {
        bias_sweep_explicit_kernel<T><<<gi, 128, 0, st>>>(cp.i32(), ci.i32(), cv.f64(), n_item, ub.template as<T>(),
                                                          ib.template as<T>(), (T)lambda, dynamic_lambda, non_negative);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (n_user) {
        bias_sweep_explicit_kernel<T><<<gu, 128, 0, st>>>(rp.i32(), ri.i32(), rv.f64(), n_user, ib.template as<T>(),
                                                          ub.template as<T>(), (T)lambda, dynamic_lambda, non_negative);
        LAUNCHED(); CU(cudaGetLastError());
      }
    }
    if (calculate_global_bias && nnz > 0) {   // the reference shifts the caller's values in place (:48-51)
      CU(cudaMemcpyAsync(csc_val, cv.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(csr_val, rv.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    }
  } else {
    CU(um.ensure(sizeof(double) * (size_t)std::max(1, n_user))); CU(ua.ensure(sizeof(double) * (size_t)std::max(1, n_user)));
    CU(im.ensure(sizeof(double) * (size_t)std::max(1, n_item))); CU(ia.ensure(sizeof(double) * (size_t)std::max(1, n_item)));
    const double lam_t = (double)(T)lambda;   // `T lambda` in the reference's signature
    if (n_user) {
      bias_means_implicit_kernel<<<gu, 128, 0, st>>>(rp.i32(), rv.f64(), n_user, n_item, lam_t, um.f64(), ua.f64());
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (n_item) {
      bias_means_implicit_kernel<<<gi, 128, 0, st>>>(cp.i32(), cv.f64(), n_item, n_user, lam_t, im.f64(), ia.f64());
      LAUNCHED(); CU(cudaGetLastError());
    }
    for (int iter = 0; iter < 5; iter++) {                                      // (:130-162)
      if (iter > 0 && n_user) TRY(device_sum((const T*)ub.template as<T>(), (long long)n_user, d_scal + 1));
      if (n_item) {
        bias_sweep_implicit_kernel<T><<<gi, 128, 0, st>>>(cp.i32(), ci.i32(), cv.f64(), n_item, n_user, ub.template as<T>(),
                                                          (iter > 0 && n_user) ? d_scal + 1 : nullptr, im.f64(), ia.f64(), g,
                                                          non_negative, ib.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
        TRY(device_sum((const T*)ib.template as<T>(), (long long)n_item, d_scal + 2));
      }
      if (n_user) {
        bias_sweep_implicit_kernel<T><<<gu, 128, 0, st>>>(rp.i32(), ri.i32(), rv.f64(), n_user, n_item, ib.template as<T>(),
                                                          n_item ? d_scal + 2 : nullptr, um.f64(), ua.f64(), g, non_negative,
                                                          ub.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
      }
    }
  }
  if (n_user) CU(cudaMemcpyAsync(user_bias, ub.p, sizeof(T) * (size_t)n_user, cudaMemcpyDeviceToHost, st));
  if (n_item) CU(cudaMemcpyAsync(item_bias, ib.p, sizeof(T) * (size_t)n_item, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (global_bias) *global_bias = g;
  return B200ALS_OK;
}
extern "C" int b200als_initialize_biases_float(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                               const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                               const int32_t* csr_idx, double* csr_val, float* user_bias, float* item_bias,
                                               double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                               int is_explicit_feedback, double* global_bias) {
  return initialize_biases_impl<float>(n_user, n_item, nnz, csc_ptr, csc_idx, csc_val, csr_ptr, csr_idx, csr_val, user_bias,
                                       item_bias, lambda, dynamic_lambda, non_negative, calculate_global_bias,
                                       is_explicit_feedback, global_bias);
}
extern "C" int b200als_initialize_biases_double(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                                const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                                const int32_t* csr_idx, double* csr_val, double* user_bias, double* item_bias,
                                                double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                                int is_explicit_feedback, double* global_bias) {
  return initialize_biases_impl<double>(n_user, n_item, nnz, csc_ptr, csc_idx, csc_val, csr_ptr, csr_idx, csr_val, user_bias,
                                        item_bias, lambda, dynamic_lambda, non_negative, calculate_global_bias,
                                        is_explicit_feedback, global_bias);
}

extern "C" int b200als_gram_float(const float* X, int rank, int64_t n, double lambda, float* XtX) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!X || !XtX || rank <= 0 || n < 0) return fail(B200ALS_EINVAL, "bad argument");
  DevBuf dX, dG;
  CU(dX.ensure(sizeof(float) * (size_t)rank * (size_t)n));
  CU(dG.ensure(sizeof(float) * (size_t)rank * rank));
  CU(cudaMemcpyAsync(dX.p, X, sizeof(float) * (size_t)rank * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  TRY(run_gram<float>(c, dX.f32(), rank, n, lambda, dG.f32(), nullptr));
  CU(cudaMemcpyAsync(XtX, dG.p, sizeof(float) * (size_t)rank * rank, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

// pinned host memory + device timers for host programs without a CUDA binding (bench.py, the R shim)
extern "C" int b200als_host_alloc(size_t bytes, void** out) {
  TRY(ctx().init());
  if (!out) return fail(B200ALS_EINVAL, "null out");
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return B200ALS_OK;
}
extern "C" int b200als_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return B200ALS_OK;
}
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
extern "C" int b200als_timer_start(void) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!g_t0) { CU(cudaEventCreate(&g_t0)); CU(cudaEventCreate(&g_t1)); }
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(g_t0, c.stream));
  return B200ALS_OK;
}
extern "C" int b200als_timer_stop(float* ms) {
  Ctx& c = ctx();
  if (!g_t0 || !ms) return fail(B200ALS_EINVAL, "timer not started");
  CU(cudaEventRecord(g_t1, c.stream));
  CU(cudaEventSynchronize(g_t1));
  CU(cudaDeviceSynchronize());
  CU(cudaEventElapsedTime(ms, g_t0, g_t1));
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// top-k recommendation: `top_product` of src/matrix_top_product.cpp:20-102
// ------------------------------------------------------------------------------------------------------
__global__ void set_bits_kernel(const int32_t* __restrict__ ids_1based, int n, int n_item, uint32_t* __restrict__ bits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int i = ids_1based[e] - 1;   // R indices
  if (i >= 0 && i < n_item) atomicOr(&bits[i >> 5], 1u << (i & 31));
}
static int run_topk(Ctx& c, const float* dX, long long n_user, const float* dY, int n_item, int rank, int top_k,
                    const int32_t* h_nr_ptr, const int32_t* h_nr_idx, const int32_t* h_exclude, int n_exclude,
                    double glob_mean, int32_t* h_idx_out, double* h_scores_out) {
  if (rank > kTopMaxRank) return fail(B200ALS_EUNSUPPORTED, "top_product: rank > 128 is not supported");
  if (top_k < 1 || top_k > kTopMaxK) return fail(B200ALS_EUNSUPPORTED, "top_product: k must be in 1..128");
  if (n_user <= 0) return B200ALS_OK;
  DevBuf nr_ptr, nr_idx, excl, bits, d_idx, d_sc;
  TopkParams P;
  P.x = dX; P.y = dY; P.n_user = n_user; P.n_item = n_item; P.rank = rank; P.top_k = top_k;
  P.nr_ptr = nullptr; P.nr_idx = nullptr; P.exclude_bits = nullptr; P.glob_mean = glob_mean;
  if (h_nr_ptr) {
    const long long nnz = h_nr_ptr[n_user];
    if (nnz > 0) {   // src/matrix_top_product.cpp:33: an empty filter matrix is ignored
      if (!h_nr_idx) return fail(B200ALS_EINVAL, "top_product: not_recommend indices missing");
      CU(nr_ptr.ensure(sizeof(int32_t) * (size_t)(n_user + 1)));
      CU(nr_idx.ensure(sizeof(int32_t) * (size_t)nnz));
      CU(cudaMemcpyAsync(nr_ptr.p, h_nr_ptr, sizeof(int32_t) * (size_t)(n_user + 1), cudaMemcpyHostToDevice, c.stream));
      CU(cudaMemcpyAsync(nr_idx.p, h_nr_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c.stream));
      P.nr_ptr = nr_ptr.i32();
      P.nr_idx = nr_idx.i32();
    }
  }
  if (n_exclude > 0) {
    if (!h_exclude) return fail(B200ALS_EINVAL, "top_product: exclude list missing");
    const size_t words = ((size_t)n_item + 31) / 32;
    CU(excl.ensure(sizeof(int32_t) * (size_t)n_exclude));
    CU(bits.ensure(sizeof(uint32_t) * words));
    CU(cudaMemsetAsync(bits.p, 0, sizeof(uint32_t) * words, c.stream));
    CU(cudaMemcpyAsync(excl.p, h_exclude, sizeof(int32_t) * (size_t)n_exclude, cudaMemcpyHostToDevice, c.stream));
    set_bits_kernel<<<(n_exclude + 255) / 256, 256, 0, c.stream>>>(excl.i32(), n_exclude, n_item, (uint32_t*)bits.p);
    LAUNCHED(); CU(cudaGetLastError());
    P.exclude_bits = (const uint32_t*)bits.p;
  }
  CU(d_idx.ensure(sizeof(int32_t) * (size_t)n_user * top_k));
  CU(d_sc.ensure(sizeof(double) * (size_t)n_user * top_k));
  P.idx_out = d_idx.i32();
  P.score_out = d_sc.f64();
  const size_t smem = sizeof(TopkSmem);
  CU(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n_user + kTopUB - 1) / kTopUB;
  topk_kernel<<<(unsigned)blocks, 256, smem, c.stream>>>(P);
  LAUNCHED(); CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h_idx_out, d_idx.p, sizeof(int32_t) * (size_t)n_user * top_k, cudaMemcpyDeviceToHost, c.stream));
  if (h_scores_out)
    CU(cudaMemcpyAsync(h_scores_out, d_sc.p, sizeof(double) * (size_t)n_user * top_k, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_top_product(const float* user_emb, int64_t n_user, const float* item_emb, int32_t n_item, int rank,
                                   int top_k, const int32_t* not_recommend_ptr, const int32_t* not_recommend_idx,
                                   const int32_t* exclude, int n_exclude, double glob_mean, int32_t* idx_out,
                                   double* scores_out) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!user_emb || !item_emb || !idx_out || n_user < 0 || n_item <= 0 || rank <= 0) return fail(B200ALS_EINVAL, "bad argument");
  DevBuf dX, dY;
  CU(dX.ensure(sizeof(float) * (size_t)rank * (size_t)std::max<int64_t>(1, n_user)));
  CU(dY.ensure(sizeof(float) * (size_t)rank * (size_t)n_item));
  CU(cudaMemcpyAsync(dX.p, user_emb, sizeof(float) * (size_t)rank * (size_t)n_user, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(dY.p, item_emb, sizeof(float) * (size_t)rank * (size_t)n_item, cudaMemcpyHostToDevice, c.stream));
  return run_topk(c, dX.f32(), n_user, dY.f32(), n_item, rank, top_k, not_recommend_ptr, not_recommend_idx, exclude, n_exclude,
                  glob_mean, idx_out, scores_out);
}

// ------------------------------------------------------------------------------------------------------
// 3. communicator (one process per GPU)
// ------------------------------------------------------------------------------------------------------
// NCCL is bound at run time, not link time: a host process may already carry a different libnccl.so.2 (PyTorch
// bundles its own), and single-GPU users need none at all.  dlopen() returns whichever copy is already loaded.
struct NcclApi {
  void* h = nullptr;
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllReduce) AllReduce = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclBroadcast) Broadcast = nullptr;
  decltype(&::ncclGroupStart) GroupStart = nullptr;
  decltype(&::ncclGroupEnd) GroupEnd = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
  int load() {
    if (h) return B200ALS_OK;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(B200ALS_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define B200ALS_SYM(name)                                                                     \
    name = reinterpret_cast<decltype(name)>(dlsym(h, "nccl" #name));                        \
    if (!name) return fail(B200ALS_ENCCL, "libnccl.so.2 lacks nccl" #name)
    B200ALS_SYM(GetUniqueId); B200ALS_SYM(CommInitRank); B200ALS_SYM(CommDestroy); B200ALS_SYM(AllReduce);
    B200ALS_SYM(AllGather); B200ALS_SYM(Broadcast); B200ALS_SYM(GroupStart); B200ALS_SYM(GroupEnd);
    B200ALS_SYM(GetErrorString);
#undef B200ALS_SYM
    return B200ALS_OK;
  }
};
static NcclApi g_nccl;
struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
static Comm g_comm;

extern "C" int b200als_comm_unique_id(void* id_out) {
  static_assert(sizeof(ncclUniqueId) == B200ALS_UNIQUE_ID_BYTES, "ncclUniqueId size");
  if (!id_out) return fail(B200ALS_EINVAL, "null id");
  TRY(g_nccl.load());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return B200ALS_OK;
}
extern "C" int b200als_comm_init(const void* id, int rank, int world_size) {
  if (!id || world_size < 1 || rank < 0 || rank >= world_size) return fail(B200ALS_EINVAL, "bad communicator arguments");
  TRY(ctx().init());
  TRY(g_nccl.load());
  if (g_comm.comm) return fail(B200ALS_EINVAL, "communicator already initialised");
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  NC(g_nccl.CommInitRank(&g_comm.comm, world_size, uid, rank));
  g_comm.rank = rank;
  g_comm.world = world_size;
  return B200ALS_OK;
}
extern "C" int b200als_comm_destroy(void) {
  if (g_comm.comm) {
    g_nccl.CommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
  }
  g_comm.rank = 0;
  g_comm.world = 1;
  return B200ALS_OK;
}
extern "C" int b200als_comm_info(int* rank, int* world_size) {
  if (rank) *rank = g_comm.rank;
  if (world_size) *world_size = g_comm.world;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// 2. session
// ------------------------------------------------------------------------------------------------------
struct b200als_session {
  b200als_options opt;
  int k = 0;
  int32_t n_user = 0, n_item = 0;
  // orientation [B200ALS_ITEMS]: columns = items (local block), idx = users ; [B200ALS_USERS]: columns = users
  CscDev<float> csc[2];
  bool has[2] = {false, false};
  int32_t shard_begin[2] = {0, 0}, shard_end[2] = {0, 0};
  int64_t nnz_global[2] = {0, 0};
  DevBuf fac[2];   // full factor matrices (stored in basis B): [ITEMS] k x n_item, [USERS] k x n_user
  DevBuf cnt[2];   // cnt[w][j] = nnz of row j of factor matrix w (global), for the dynamic-lambda regulariser
  DevBuf G, G64, Vt, Q, Qt, diag, B64, Btmp, Bf, scratch;
  bool basis_identity = true;
  std::vector<int32_t> ranges[2];   // [3*world]: every rank's [begin, end, can_chunk) per orientation (multi-GPU)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_chunk[8] = {}, ev_comm_done = nullptr;
  // peer-memory exchange (multi-GPU): every rank maps every other rank's factor matrices (CUDA IPC) and pushes its
  // freshly solved rows straight into them with the copy engines over NVLink -- no SM is taken from the solve
  static constexpr int kMaxPeers = 16;
  int p2p_state = 0;                      // 0 not tried yet, 1 active, -1 unavailable (NCCL broadcasts instead)
  float* peer_fac[2][kMaxPeers] = {};
  cudaStream_t push_stream[kMaxPeers] = {};
  cudaEvent_t ev_push[kMaxPeers] = {};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float t_gram = 0, t_prep = 0, t_solve = 0, t_comm = 0;
};

extern "C" void b200als_default_options(b200als_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->feedback = B200ALS_IMPLICIT;
  o->solver = B200ALS_CONJUGATE_GRADIENT;
  o->cg_steps = 3;
  o->dynamic_lambda = 1;
  o->lambda = 0.0;
  o->kernel = 0;
}

static int session_alloc(b200als_session* s) {
  Ctx& c = ctx();
  const size_t k = (size_t)s->k;
  CU(s->fac[B200ALS_ITEMS].ensure(sizeof(float) * k * (size_t)s->n_item));
  CU(s->fac[B200ALS_USERS].ensure(sizeof(float) * k * (size_t)s->n_user));
  CU(cudaMemsetAsync(s->fac[0].p, 0, s->fac[0].bytes, c.stream));
  CU(cudaMemsetAsync(s->fac[1].p, 0, s->fac[1].bytes, c.stream));
  CU(s->G.ensure(sizeof(float) * k * k));
  CU(s->G64.ensure(sizeof(double) * k * k));
  CU(s->Vt.ensure(sizeof(double) * k * k));
  CU(s->Q.ensure(sizeof(float) * k * k));
  CU(s->Qt.ensure(sizeof(float) * k * k));
  CU(s->diag.ensure(sizeof(float) * k));
  CU(s->B64.ensure(sizeof(double) * k * k));
  CU(s->Btmp.ensure(sizeof(double) * k * k));
  CU(s->Bf.ensure(sizeof(float) * k * k));
  set_identity_kernel<<<(unsigned)((k * k + 255) / 256), 256, 0, c.stream>>>(s->B64.f64(), (int)k);
  LAUNCHED(); CU(cudaGetLastError());
  s->basis_identity = true;
  for (auto& e : s->ev) CU(cudaEventCreate(&e));
  for (auto& e : s->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s->ev_comm_done, cudaEventDisableTiming));
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));  // the exchange must get SM slots early
  return B200ALS_OK;
}

// cnt[which][j] += number of entries of row j seen in the local block of the *other* orientation
__global__ void count_idx_kernel(const int32_t* __restrict__ idx, long long nnz, float* __restrict__ cnt) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(&cnt[idx[e]], 1.0f);
}

static int session_counts(b200als_session* s) {
  // cnt_X for the dynamic-lambda regulariser (R/model_WRMF.R:305-315): nnz per row of the FIXED matrix, i.e. for
  // the user half (X = items) the nnz per item.  Counted from whichever orientation is present.
  Ctx& c = ctx();
  for (int w = 0; w < 2; w++) {
    const int32_t n = (w == B200ALS_ITEMS) ? s->n_item : s->n_user;
    CU(s->cnt[w].ensure(sizeof(float) * (size_t)std::max(1, n)));
    CU(cudaMemsetAsync(s->cnt[w].p, 0, sizeof(float) * (size_t)n, c.stream));
  }
  // orientation USERS has idx = items -> counts per item ; orientation ITEMS has idx = users -> counts per user
  for (int w = 0; w < 2; w++) {
    if (!s->has[w] || s->csc[w].nnz == 0) continue;
    const int other = 1 - w;
    count_idx_kernel<<<(unsigned)((s->csc[w].nnz + 255) / 256), 256, 0, c.stream>>>(s->csc[w].idx.i32(), s->csc[w].nnz,
                                                                                   s->cnt[other].f32());
    LAUNCHED(); CU(cudaGetLastError());
  }
  if (g_comm.world > 1) {
    for (int w = 0; w < 2; w++) {
      const int32_t n = (w == B200ALS_ITEMS) ? s->n_item : s->n_user;
      if (s->has[1 - w]) NC(g_nccl.AllReduce(s->cnt[w].p, s->cnt[w].p, (size_t)n, ncclFloat, ncclSum, g_comm.comm, c.stream));
    }
  }
  return B200ALS_OK;
}

extern "C" int b200als_create(b200als_session** out, const b200als_csc* c_ui, const b200als_csc* c_iu, int32_t n_user,
                              int32_t n_item, int rank, const b200als_options* opts) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!out || rank <= 0 || n_user < 0 || n_item < 0) return fail(B200ALS_EINVAL, "bad argument");
  if (rank > 256) return fail(B200ALS_EUNSUPPORTED, "rank > 256 is not supported");
  b200als_session* s = new b200als_session();
  if (opts) s->opt = *opts; else b200als_default_options(&s->opt);
  s->k = rank;
  s->n_user = n_user;
  s->n_item = n_item;
  int rc = session_alloc(s);
  if (rc == B200ALS_OK && c_ui) {
    rc = upload_csc<float>(c_ui, s->csc[B200ALS_ITEMS], c.stream);
    s->has[B200ALS_ITEMS] = true;
    s->shard_begin[B200ALS_ITEMS] = 0;
    s->shard_end[B200ALS_ITEMS] = c_ui->n_cols;
  }
  if (rc == B200ALS_OK && c_iu) {
    rc = upload_csc<float>(c_iu, s->csc[B200ALS_USERS], c.stream);
    s->has[B200ALS_USERS] = true;
    s->shard_begin[B200ALS_USERS] = 0;
    s->shard_end[B200ALS_USERS] = c_iu->n_cols;
  }
  if (rc == B200ALS_OK) rc = session_counts(s);
  for (int w = 0; w < 2 && rc == B200ALS_OK; w++) {
    long long nnz = s->has[w] ? s->csc[w].nnz : 0;
    if (g_comm.world > 1) {
      DevBuf t;
      if (t.ensure(sizeof(long long)) != cudaSuccess) { rc = fail(B200ALS_ECUDA, "alloc"); break; }
      cudaMemcpyAsync(t.p, &nnz, sizeof(nnz), cudaMemcpyHostToDevice, c.stream);
      if (g_nccl.AllReduce(t.p, t.p, 1, ncclInt64, ncclSum, g_comm.comm, c.stream) != ncclSuccess) { rc = fail(B200ALS_ENCCL, "allreduce nnz"); break; }
      cudaMemcpyAsync(&nnz, t.p, sizeof(nnz), cudaMemcpyDeviceToHost, c.stream);
      cudaStreamSynchronize(c.stream);
    }
    s->nnz_global[w] = nnz;
  }
  if (rc == B200ALS_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(B200ALS_ECUDA, "sync after upload");
  if (rc != B200ALS_OK) {
    b200als_destroy(s);
    return rc;
  }
  *out = s;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Format ingest on the device: build the other orientation of the sparse matrix, i.e. what
// `MatrixExtra::as.csr.matrix` / `t_shallow` do on the host in R/model_WRMF.R:184-189.  A stable LSD radix sort
// (CUB) of the entries by their row id keeps, inside every new column, the source order = ascending source
// column, so the result satisfies the dgCMatrix invariant and is bit-reproducible.
// ------------------------------------------------------------------------------------------------------
__global__ void expand_columns_kernel(const int32_t* __restrict__ ptr, int n_cols, int32_t* __restrict__ col_of) {
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= n_cols) return;
  for (int e = ptr[cidx]; e < ptr[cidx + 1]; e++) col_of[e] = cidx;
}
__global__ void iota_kernel(int32_t* __restrict__ a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
__global__ void gather_transposed_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ col_of,
                                         const float* __restrict__ val, long long nnz, int32_t* __restrict__ idx_out,
                                         float* __restrict__ val_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  const int32_t e = perm[t];
  idx_out[t] = col_of[e];
  val_out[t] = val[e];
}
// ptr_out[r] = first position in the sorted key array with key >= r  (r = 0 .. n_rows)
__global__ void row_starts_kernel(const int32_t* __restrict__ sorted_keys, long long nnz, int n_rows, int32_t* __restrict__ ptr_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < r) lo = mid + 1; else hi = mid;
  }
  ptr_out[r] = (int32_t)lo;
}
static int transpose_on_device(Ctx& c, const CscDev<float>& src, CscDev<float>& dst) {
  dst.n_rows = src.n_cols;
  dst.n_cols = src.n_rows;
  dst.nnz = src.nnz;
  dst.n_short = -1;
  const long long nnz = src.nnz;
  CU(dst.ptr.ensure(sizeof(int32_t) * ((size_t)dst.n_cols + 1)));
  CU(dst.idx.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(dst.val.ensure(sizeof(float) * (size_t)nnz));
  if (nnz == 0) {
    CU(cudaMemsetAsync(dst.ptr.p, 0, sizeof(int32_t) * ((size_t)dst.n_cols + 1), c.stream));
    return B200ALS_OK;
  }
  DevBuf col_of, iota, perm, keys_out, temp;
  CU(col_of.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(iota.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(perm.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(keys_out.ensure(sizeof(int32_t) * (size_t)nnz));
  expand_columns_kernel<<<(src.n_cols + 255) / 256, 256, 0, c.stream>>>(src.ptr.i32(), src.n_cols, col_of.i32());
  LAUNCHED(); CU(cudaGetLastError());
  iota_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c.stream>>>(iota.i32(), nnz);
  LAUNCHED(); CU(cudaGetLastError());
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)std::max(1, src.n_rows)) bits++;
  size_t temp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, src.idx.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz, 0,
                                     bits, c.stream));
  CU(temp.ensure(temp_bytes));
  CU(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, src.idx.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz, 0,
                                     bits, c.stream));
  LAUNCHED();
  gather_transposed_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c.stream>>>(perm.i32(), col_of.i32(), src.val.f32(), nnz,
                                                                              dst.idx.i32(), dst.val.f32());
  LAUNCHED(); CU(cudaGetLastError());
  row_starts_kernel<<<(dst.n_cols + 1 + 255) / 256, 256, 0, c.stream>>>(keys_out.i32(), nnz, dst.n_cols, dst.ptr.i32());
  LAUNCHED(); CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_build_missing_orientation(b200als_session* s) {
  Ctx& c = ctx();
  if (!s) return fail(B200ALS_EINVAL, "null session");
  if (s->has[0] && s->has[1]) return B200ALS_OK;
  if (!s->has[0] && !s->has[1]) return fail(B200ALS_EINVAL, "the session holds no sparse matrix");
  if (g_comm.world > 1) return fail(B200ALS_EUNSUPPORTED, "device-side transpose of a sharded matrix is not implemented");
  const int have = s->has[0] ? 0 : 1, need = 1 - have;
  TRY(transpose_on_device(c, s->csc[have], s->csc[need]));
  s->has[need] = true;
  s->shard_begin[need] = 0;
  s->shard_end[need] = s->csc[need].n_cols;
  s->nnz_global[need] = s->csc[need].nnz;
  s->ranges[need].clear();
  return session_counts(s);
}
// copy one orientation back to the host (tests / export): ptr[n_cols+1], idx[nnz], val[nnz]
extern "C" int b200als_get_orientation(b200als_session* s, int which, int32_t* ptr, int32_t* idx, float* val, int64_t* nnz_out) {
  Ctx& c = ctx();
  if (!s || which < 0 || which > 1 || !s->has[which]) return fail(B200ALS_EINVAL, "orientation not present");
  const CscDev<float>& A = s->csc[which];
  if (nnz_out) *nnz_out = A.nnz;
  if (ptr) CU(cudaMemcpyAsync(ptr, A.ptr.p, sizeof(int32_t) * ((size_t)A.n_cols + 1), cudaMemcpyDeviceToHost, c.stream));
  if (idx && A.nnz) CU(cudaMemcpyAsync(idx, A.idx.p, sizeof(int32_t) * (size_t)A.nnz, cudaMemcpyDeviceToHost, c.stream));
  if (val && A.nnz) CU(cudaMemcpyAsync(val, A.val.p, sizeof(float) * (size_t)A.nnz, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_destroy(b200als_session* s) {
  if (!s) return B200ALS_OK;
  for (auto& e : s->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : s->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (s->ev_comm_done) cudaEventDestroy(s->ev_comm_done);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  if (s->p2p_state == 1) {
    cudaDeviceSynchronize();
    for (int w = 0; w < 2; w++)
      for (int r = 0; r < b200als_session::kMaxPeers; r++)
        if (s->peer_fac[w][r]) cudaIpcCloseMemHandle(s->peer_fac[w][r]);
    for (int r = 0; r < b200als_session::kMaxPeers; r++) {
      if (s->push_stream[r]) cudaStreamDestroy(s->push_stream[r]);
      if (s->ev_push[r]) cudaEventDestroy(s->ev_push[r]);
    }
    // nobody frees a matrix a peer still has mapped: destroy is collective while the communicator lives
    if (g_comm.comm) {
      Ctx& c = ctx();
      cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream);
      g_nccl.AllReduce(c.status.p, c.status.p, 1, ncclInt32, ncclSum, g_comm.comm, c.stream);
      cudaStreamSynchronize(c.stream);
    }
  }
  delete s;
  return B200ALS_OK;
}

extern "C" int b200als_exchange_mode(b200als_session* s, int* mode) {
  if (!s || !mode) return fail(B200ALS_EINVAL, "bad argument");
  *mode = (g_comm.world <= 1 || s->p2p_state == 0) ? 0 : (s->p2p_state == 1 ? 1 : 2);
  return B200ALS_OK;
}

extern "C" int b200als_set_shard(b200als_session* s, int which, int32_t begin, int32_t end) {
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const int32_t n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  if (begin < 0 || end < begin || end > n || (s->has[which] && end - begin != s->csc[which].n_cols))
    return fail(B200ALS_EINVAL, "shard range does not match the uploaded block");
  s->shard_begin[which] = begin;
  s->shard_end[which] = end;
  s->ranges[which].clear();
  return B200ALS_OK;
}

static int rotate_matrix(Ctx& c, float* M, long long n, const float* R) {
  if (n <= 0) return B200ALS_OK;
  // large matrices: tcgen05 3xTF32 kernel (B200ALS_ROTATE=ffma forces the fp32 FMA kernel)
  const char* env = getenv("B200ALS_ROTATE");
  const bool force_tc = env && (env[0] == 't' || env[0] == 'T');   // tests
  const bool tc = force_tc || (!(env && (env[0] == 'f' || env[0] == 'F')) && n >= 65536);
  if (tc) {
    CU(c.rot_rt.ensure(sizeof(float) * kTcK * kTcK));
    transpose_128_kernel<<<(kTcK * kTcK + 255) / 256, 256, 0, c.stream>>>(R, c.rot_rt.f32());
    LAUNCHED(); CU(cudaGetLastError());
    const size_t smem = sizeof(RotTcSmem);
    CU(cudaFuncSetAttribute(rotate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (n + 127) / 128;
    const int grid = (int)std::min<long long>(tiles, c.sm_count);
    rotate_tc_kernel<<<grid, 128, smem, c.stream>>>(M, M, c.rot_rt.f32(), n);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  }
  const size_t smem = sizeof(RotSmem);
  CU(cudaFuncSetAttribute(rotate_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n + kRotRows - 1) / kRotRows;
  const int grid = (int)std::min<long long>(blocks, c.sm_count * 2);
  rotate_rows_kernel<<<grid, 256, smem, c.stream>>>(M, M, R, n);
  LAUNCHED(); CU(cudaGetLastError());
  return B200ALS_OK;
}

// true = stored * B'  <=> stored = true * B.  Export / import copies through a scratch buffer.
extern "C" int b200als_set_factors(b200als_session* s, int which, const float* host) {
  Ctx& c = ctx();
  if (!s || !host || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  CU(cudaMemcpyAsync(s->fac[which].p, host, sizeof(float) * (size_t)s->k * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  if (!s->basis_identity) {
    convert_kk_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->Bf.f32(), s->k, 0);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, s->fac[which].f32(), n, s->Bf.f32()));
  }
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}
static int export_rotated(b200als_session* s, const float* dev, long long n, float* host) {
  Ctx& c = ctx();
  const size_t bytes = sizeof(float) * (size_t)s->k * (size_t)n;
  if (s->basis_identity) {
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
  } else {
    CU(s->scratch.ensure(bytes));
    CU(cudaMemcpyAsync(s->scratch.p, dev, bytes, cudaMemcpyDeviceToDevice, c.stream));
    convert_kk_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->Bf.f32(), s->k, 1);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, s->scratch.f32(), n, s->Bf.f32()));
    CU(cudaMemcpyAsync(host, s->scratch.p, bytes, cudaMemcpyDeviceToHost, c.stream));
  }
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}
extern "C" int b200als_get_factors(b200als_session* s, int which, float* host) {
  if (!s || !host || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  return export_rotated(s, s->fac[which].f32(), n, host);
}

extern "C" int b200als_init_factors(b200als_session* s, uint64_t seed) {
  Ctx& c = ctx();
  if (!s) return fail(B200ALS_EINVAL, "null session");
  const long long nu = (long long)s->k * s->n_user, ni = (long long)s->k * s->n_item;
  if (nu) init_normal_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, c.stream>>>(s->fac[B200ALS_USERS].f32(), nu, seed, 0.01f);
  LAUNCHED(); CU(cudaGetLastError());
  if (s->opt.solver == B200ALS_CONJUGATE_GRADIENT) {
    CU(cudaMemsetAsync(s->fac[B200ALS_ITEMS].p, 0, sizeof(float) * (size_t)ni, c.stream));  // R/model_WRMF.R:217-230
  } else if (ni) {
    init_normal_kernel<<<(unsigned)((ni + 255) / 256), 256, 0, c.stream>>>(s->fac[B200ALS_ITEMS].f32(), ni,
                                                                           seed ^ 0xA5A5A5A5ull, 0.01f);
    LAUNCHED(); CU(cudaGetLastError());
  }
  set_identity_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->k);
  LAUNCHED(); CU(cudaGetLastError());
  s->basis_identity = true;
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_randomize_factors(b200als_session* s, int which, uint64_t seed, float scale, float decay) {
  Ctx& c = ctx();
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (long long)s->k * ((which == B200ALS_ITEMS) ? s->n_item : s->n_user);
  if (n) init_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(s->fac[which].f32(), n, seed, scale, s->k, decay);
  LAUNCHED(); CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

// every rank learns every rank's [begin, end) of the solved matrix (cached until set_shard)
static int gather_ranges(b200als_session* s, int which) {
  if (!s->ranges[which].empty()) return B200ALS_OK;
  Ctx& c = ctx();
  TRY(classify_rows(c, s->csc[which]));
  std::vector<int32_t> ranges(3 * g_comm.world);
  DevBuf d;
  CU(d.ensure(sizeof(int32_t) * 3 * g_comm.world));
  int32_t mine[3] = {s->shard_begin[which], s->shard_end[which],
                     (s->csc[which].all_short && s->csc[which].n_cols >= 8 * 4096) ? 1 : 0};
  CU(cudaMemcpyAsync(d.i32() + 3 * g_comm.rank, mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllGather(d.i32() + 3 * g_comm.rank, d.p, 3, ncclInt32, g_comm.comm, c.stream));
  CU(cudaMemcpyAsync(ranges.data(), d.p, sizeof(int32_t) * 3 * g_comm.world, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  s->ranges[which] = ranges;
  return B200ALS_OK;
}
// exchange of the freshly solved rows: chunk `ch` of `n_ch` of every rank's block, one broadcast per owner
// (unequal block sizes allowed), grouped, on stream `st`.
static int exchange_chunk(b200als_session* s, int which, int ch, int n_ch, cudaStream_t st) {
  if (g_comm.world <= 1) return B200ALS_OK;
  const std::vector<int32_t>& ranges = s->ranges[which];
  float* M = s->fac[which].f32();
  NC(g_nccl.GroupStart());
  for (int r = 0; r < g_comm.world; r++) {
    const long long rb = ranges[3 * r], len = ranges[3 * r + 1] - rb;
    const long long cb = rb + len * ch / n_ch, ce = rb + len * (ch + 1) / n_ch;
    if (ce <= cb) continue;
    float* p = M + (size_t)cb * s->k;
    NC(g_nccl.Broadcast(p, p, (size_t)(ce - cb) * (size_t)s->k, ncclFloat, r, g_comm.comm, st));
  }
  NC(g_nccl.GroupEnd());
  return B200ALS_OK;
}
// Maps the peers' factor matrices.  Collective: every rank calls it at the same point.  Falls back to NCCL (state -1)
// unless every rank could open every handle (B200ALS_EXCHANGE=nccl forces the fallback, =p2p makes failure an error).
static int p2p_setup(b200als_session* s) {
  if (s->p2p_state != 0) return B200ALS_OK;
  Ctx& c = ctx();
  const char* ev = getenv("B200ALS_EXCHANGE");
  const bool force_nccl = ev && !strcmp(ev, "nccl"), force_p2p = ev && !strcmp(ev, "p2p");
  const int W = g_comm.world, me = g_comm.rank;
  int ok = (!force_nccl && W <= b200als_session::kMaxPeers) ? 1 : 0;
  cudaIpcMemHandle_t mine[2];
  std::memset(mine, 0, sizeof(mine));
  if (ok)
    for (int w = 0; w < 2; w++)
      if (cudaIpcGetMemHandle(&mine[w], s->fac[w].p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  const size_t hb = sizeof(mine);
  DevBuf d, flag;
  CU(d.ensure(hb * (size_t)W));
  CU(flag.ensure(sizeof(int)));
  CU(cudaMemcpyAsync((char*)d.p + hb * me, mine, hb, cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllGather((char*)d.p + hb * me, d.p, hb, ncclChar, g_comm.comm, c.stream));
  std::vector<cudaIpcMemHandle_t> all(2 * (size_t)W);
  CU(cudaMemcpyAsync(all.data(), d.p, hb * (size_t)W, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, g_comm.comm, c.stream));
  CU(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  if (ok) {
    for (int r = 0; r < W && ok; r++) {
      if (r == me) continue;
      for (int w = 0; w < 2; w++) {
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[2 * (size_t)r + w], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = 0;
          cudaGetLastError();
          break;
        }
        s->peer_fac[w][r] = (float*)q;
      }
    }
    CU(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
    NC(g_nccl.AllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, g_comm.comm, c.stream));
    CU(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
  }
  if (!ok) {
    for (int w = 0; w < 2; w++)
      for (int r = 0; r < b200als_session::kMaxPeers; r++)
        if (s->peer_fac[w][r]) { cudaIpcCloseMemHandle(s->peer_fac[w][r]); s->peer_fac[w][r] = nullptr; }
    s->p2p_state = -1;
    if (force_p2p) return fail(B200ALS_ECUDA, "B200ALS_EXCHANGE=p2p: peer mapping of the factor matrices failed on some rank");
    return B200ALS_OK;
  }
  for (int r = 0; r < W; r++) {
    if (r == me) continue;
    CU(cudaStreamCreateWithFlags(&s->push_stream[r], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s->ev_push[r], cudaEventDisableTiming));
  }
  s->p2p_state = 1;
  return B200ALS_OK;
}
// pushes chunk `ch` of `n_ch` of this rank's block into every peer's copy of the matrix, one copy-engine stream per
// peer, after `ready` (the chunk's solve).  Completion is collected by p2p_join().
static int p2p_push_chunk(b200als_session* s, int which, int ch, int n_ch, cudaEvent_t ready) {
  const std::vector<int32_t>& ranges = s->ranges[which];
  const int me = g_comm.rank;
  const long long rb = ranges[3 * me], len = ranges[3 * me + 1] - rb;
  const long long cb = rb + len * ch / n_ch, ce = rb + len * (ch + 1) / n_ch;
  if (ce <= cb) return B200ALS_OK;
  const size_t off = (size_t)cb * s->k, bytes = sizeof(float) * (size_t)(ce - cb) * s->k;
  const float* src = s->fac[which].f32() + off;
  for (int i = 1; i < g_comm.world; i++) {
    const int r = (me + i) % g_comm.world;   // staggered start: no two ranks open on the same destination
    CU(cudaStreamWaitEvent(s->push_stream[r], ready, 0));
    CU(cudaMemcpyAsync(s->peer_fac[which][r] + off, src, bytes, cudaMemcpyDeviceToDevice, s->push_stream[r]));
  }
  return B200ALS_OK;
}
static int p2p_join(b200als_session* s, cudaStream_t st) {
  for (int r = 0; r < g_comm.world; r++) {
    if (r == g_comm.rank) continue;
    CU(cudaEventRecord(s->ev_push[r], s->push_stream[r]));
    CU(cudaStreamWaitEvent(st, s->ev_push[r], 0));
  }
  return B200ALS_OK;
}
__global__ void finalize_gram_kernel(double* __restrict__ G64, float* __restrict__ G, int k, double lambda) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  const double v = G64[e] + (((e / k) == (e % k)) ? lambda : 0.0);
  G64[e] = v;
  G[e] = (float)v;
}

// solve for `which`; Yout == nullptr: in place into the session's factors.  Yout != nullptr (transform_):
// solve into Yout (device, local block) starting from zeros.
static int session_half(b200als_session* s, int which, int solver, float* Yout, double* loss) {
  Ctx& c = ctx();
  if (!s->has[which]) return fail(B200ALS_EINVAL, "the orientation needed for this half-iteration was not supplied");
  const int fixed = 1 - which;
  const long long n_fixed = (fixed == B200ALS_ITEMS) ? s->n_item : s->n_user;
  float* X = s->fac[fixed].f32();
  float* Yfull = s->fac[which].f32();
  float* Y = Yout ? Yout : (Yfull + (size_t)s->shard_begin[which] * s->k);
  HalfOpts o{s->opt.feedback, solver, s->opt.cg_steps, s->opt.dynamic_lambda, s->opt.kernel, s->opt.lambda, s->opt.reserved[0], s->opt.reserved[1]};
  CscDev<float>& A = s->csc[which];
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  CU(cudaEventRecord(s->ev[0], c.stream));
  const float* G = nullptr;
  const float* diag = nullptr;
  if (g_comm.world > 1) TRY(gather_ranges(s, which));
  if (g_comm.world > 1 && !Yout) TRY(p2p_setup(s));
  const bool p2p = (g_comm.world > 1 && !Yout && s->p2p_state == 1);
  if (p2p && !implicit) {
    // one-sided pushes need every rank to have finished its earlier writes to the matrix (set_factors, randomize, the
    // previous half-iteration) before any peer writes into it: implicit feedback gets that from the Gram all-reduce
    // below, explicit feedback from this one-word all-reduce
    CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
    NC(g_nccl.AllReduce(c.status.p, c.status.p, 1, ncclInt32, ncclSum, g_comm.comm, c.stream));
  }
  if (implicit) {
    if (g_comm.world > 1) {
      // each rank reduces its 1/world slice of the fixed matrix; the k x k partials are summed over NVLink
      const long long b = n_fixed * g_comm.rank / g_comm.world, e = n_fixed * (g_comm.rank + 1) / g_comm.world;
      TRY(run_gram<float>(c, X + (size_t)b * s->k, s->k, e - b, 0.0, s->G.f32(), s->G64.f64()));
      NC(g_nccl.AllReduce(s->G64.p, s->G64.p, (size_t)s->k * s->k, ncclDouble, ncclSum, g_comm.comm, c.stream));
      finalize_gram_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->G64.f64(), s->G.f32(), s->k, o.lambda);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      TRY(run_gram<float>(c, X, s->k, n_fixed, o.lambda, s->G.f32(), s->G64.f64()));
    }
    G = s->G.f32();
  }
  CU(cudaEventRecord(s->ev[1], c.stream));
  // eigenbasis path: implicit CG, rank 128, resident kernel, enough rows to amortise the rotation
  bool use_diag = implicit && solver == B200ALS_CONJUGATE_GRADIENT && s->k == kResK && o.kernel != 1 && o.kernel != 2 && !Yout;
  if (use_diag) {
    TRY(classify_rows(c, A));
    if (o.kernel != 3 && (A.n_long > 0 || (long long)A.n_cols * g_comm.world < 50000)) use_diag = false;
  }
  if (use_diag) {
    const size_t jsm = sizeof(double) * (size_t)s->k * (s->k + 1);
    const int a_in_smem = (jsm + 8192 <= c.smem_optin) ? 1 : 0;
    if (a_in_smem) CU(cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
    jacobi_eig_kernel<<<1, kJacobiThreads, a_in_smem ? jsm : 0, c.stream>>>(s->G64.f64(), s->Vt.f64(), s->k, s->Q.f32(),
                                                                          s->diag.f32(), s->Btmp.f64(), 30, a_in_smem);
    LAUNCHED(); CU(cudaGetLastError());
    // fixed <- fixed Q (whole matrix), solved slice <- slice Q, B <- B Q
    TRY(rotate_matrix(c, X, n_fixed, s->Q.f32()));
    TRY(rotate_matrix(c, Y, A.n_cols, s->Q.f32()));
    matmul_kk_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->Btmp.f64(),
                                                                     s->Vt.f64(), s->k);
    LAUNCHED(); CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->B64.p, s->Vt.p, sizeof(double) * (size_t)s->k * s->k, cudaMemcpyDeviceToDevice, c.stream));
    s->basis_identity = false;
    diag = s->diag.f32();
    G = nullptr;
  }
  CU(cudaEventRecord(s->ev[2], c.stream));
  if (g_comm.world > 1 && !Yout) {
    // the block is solved in chunks; chunk c travels to the other ranks (priority stream) while chunk c+1 is solved
    // every rank must take the same decision: chunk only if every block qualifies.  More chunks = shorter
    // exposed tail of the exchange (only the last chunk's broadcast is not hidden behind a solve)
    int n_ch = 8;
    if (const char* ev = getenv("B200ALS_EXCHANGE_CHUNKS")) n_ch = std::max(1, std::min(8, atoi(ev)));
    for (int r = 0; r < g_comm.world; r++)
      if (!s->ranges[which][3 * r + 2]) n_ch = 1;
    for (int ch = 0; ch < n_ch; ch++) {
      HalfOpts oc = o;
      if (n_ch > 1) {
        oc.row_begin = (int)((long long)A.n_cols * ch / n_ch);
        oc.row_count = (int)((long long)A.n_cols * (ch + 1) / n_ch) - oc.row_begin;
      }
      oc.reset_loss = (ch == 0);
      TRY(solve_rows<float>(c, A, X, Y, G, diag, s->k, oc));
      CU(cudaEventRecord(s->ev_chunk[ch], c.stream));
      if (p2p) {
        TRY(p2p_push_chunk(s, which, ch, n_ch, s->ev_chunk[ch]));
      } else {
        CU(cudaStreamWaitEvent(s->comm_stream, s->ev_chunk[ch], 0));
        TRY(exchange_chunk(s, which, ch, n_ch, s->comm_stream));
      }
    }
    CU(cudaEventRecord(s->ev[3], c.stream));
    if (p2p) {
      // own pushes done; the loss all-reduce below completes only when every rank got here, i.e. when every push
      // into this rank's matrix has landed
      TRY(p2p_join(s, c.stream));
    } else {
      CU(cudaEventRecord(s->ev_comm_done, s->comm_stream));
      CU(cudaStreamWaitEvent(c.stream, s->ev_comm_done, 0));
    }
  } else {
    TRY(solve_rows<float>(c, A, X, Y, G, diag, s->k, o));
    CU(cudaEventRecord(s->ev[3], c.stream));
  }
  CU(cudaEventRecord(s->ev[4], c.stream));
  // loss: local row sums -> global
  double rows_sum = 0.0;
  {
    double h = 0.0;
    if (g_comm.world > 1) NC(g_nccl.AllReduce(c.loss_acc.p, c.loss_acc.p, 1, ncclDouble, ncclSum, g_comm.comm, c.stream));
    CU(cudaMemcpyAsync(&h, c.loss_acc.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    rows_sum = h;
  }
  TRY(finish_loss<float>(c, X, s->k, n_fixed, s->cnt[fixed].f32(), o, s->nnz_global[which], rows_sum, true, loss));
  cudaEventElapsedTime(&s->t_gram, s->ev[0], s->ev[1]);
  cudaEventElapsedTime(&s->t_prep, s->ev[1], s->ev[2]);
  cudaEventElapsedTime(&s->t_solve, s->ev[2], s->ev[3]);
  cudaEventElapsedTime(&s->t_comm, s->ev[3], s->ev[4]);
  return B200ALS_OK;
}

extern "C" int b200als_half_iteration(b200als_session* s, int which, int solver_override, double* loss) {
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  return session_half(s, which, solver_override >= 0 ? solver_override : s->opt.solver, nullptr, loss);
}

extern "C" int b200als_fit(b200als_session* s, int n_iter, double convergence_tol, double* loss_trace, int* n_iter_done) {
  if (!s || n_iter < 0) return fail(B200ALS_EINVAL, "bad argument");
  double loss_prev = INFINITY;
  int done = 0;
  for (int i = 0; i < n_iter; i++) {  // R/model_WRMF.R:318-338
    double li = 0, lu = 0;
    TRY(session_half(s, B200ALS_ITEMS, s->opt.solver, nullptr, &li));
    TRY(session_half(s, B200ALS_USERS, s->opt.solver, nullptr, &lu));
    if (loss_trace) { loss_trace[2 * i] = li; loss_trace[2 * i + 1] = lu; }
    done = i + 1;
    if (loss_prev / lu - 1 < convergence_tol) break;
    loss_prev = lu;
  }
  if (n_iter_done) *n_iter_done = done;
  return B200ALS_OK;
}

extern "C" int b200als_transform(b200als_session* s, float* host_out, double* loss) {
  Ctx& c = ctx();
  if (!s || !host_out) return fail(B200ALS_EINVAL, "bad argument");
  if (!s->has[B200ALS_USERS]) return fail(B200ALS_EINVAL, "transform needs the users orientation");
  CscDev<float>& A = s->csc[B200ALS_USERS];
  DevBuf res;
  const size_t bytes = sizeof(float) * (size_t)s->k * (size_t)A.n_cols;
  CU(res.ensure(bytes));
  CU(cudaMemsetAsync(res.p, 0, bytes, c.stream));  // res = zeros (R/model_WRMF.R:423-427)
  const int solver = (s->opt.solver == B200ALS_CONJUGATE_GRADIENT) ? B200ALS_CHOLESKY : s->opt.solver;  // avoid_cg (:112)
  TRY(session_half(s, B200ALS_USERS, solver, res.f32(), loss));
  return export_rotated(s, res.f32(), A.n_cols, host_out);
}

extern "C" int b200als_last_timing(b200als_session* s, float* gram_ms, float* prep_ms, float* solve_ms, float* comm_ms) {
  if (!s) return fail(B200ALS_EINVAL, "null session");
  if (gram_ms) *gram_ms = s->t_gram;
  if (prep_ms) *prep_ms = s->t_prep;
  if (solve_ms) *solve_ms = s->t_solve;
  if (comm_ms) *comm_ms = s->t_comm;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// 4. synthetic workloads
// ------------------------------------------------------------------------------------------------------
extern "C" int b200als_synth_csr_host(int32_t n_rows, int32_t n_cols, int32_t nnz_per_row, uint64_t seed, int explicit_values,
                                      int64_t row_offset, int32_t* ptr, int32_t* idx, float* val_f32, double* val_f64) {
  if (n_rows < 0 || n_cols <= 0 || nnz_per_row <= 0 || nnz_per_row > n_cols || !ptr || !idx)
    return fail(B200ALS_EINVAL, "bad synthetic shape");
  if ((long long)n_rows * nnz_per_row > 2147483647LL) return fail(B200ALS_EINVAL, "nnz exceeds 32-bit row pointers");
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=]() {
      const int64_t r0 = (int64_t)n_rows * t / nt, r1 = (int64_t)n_rows * (t + 1) / nt;
      for (int64_t r = r0; r < r1; r++) {
        ptr[r] = (int32_t)(r * nnz_per_row);
        for (int j = 0; j < nnz_per_row; j++) {
          int32_t col; float v;
          synth_entry(r + row_offset, j, n_cols, nnz_per_row, seed, explicit_values, &col, &v);
          const int64_t e = r * nnz_per_row + j;
          idx[e] = col;
          if (val_f32) val_f32[e] = v;
          if (val_f64) val_f64[e] = (double)v;
        }
      }
    });
  for (auto& x : th) x.join();
  ptr[n_rows] = (int32_t)((int64_t)n_rows * nnz_per_row);
  return B200ALS_OK;
}

extern "C" int b200als_create_synthetic(b200als_session** out, int32_t n_user_local, int64_t user_offset, int32_t n_user_global,
                                        int32_t n_item, int32_t nnz_per_row, uint64_t seed, int rank,
                                        const b200als_options* opts) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!out || rank <= 0 || n_user_local < 0 || n_item <= 0 || nnz_per_row <= 0 || nnz_per_row > n_item)
    return fail(B200ALS_EINVAL, "bad argument");
  if ((long long)n_user_local * nnz_per_row > 2147483647LL) return fail(B200ALS_EINVAL, "local nnz exceeds 32-bit row pointers");
  b200als_session* s = new b200als_session();
  if (opts) s->opt = *opts; else b200als_default_options(&s->opt);
  s->k = rank;
  s->n_user = n_user_global;
  s->n_item = n_item;
  int rc = session_alloc(s);
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  CscDev<float>& A = s->csc[B200ALS_USERS];
  A.n_rows = n_item;
  A.n_cols = n_user_local;
  A.nnz = (int64_t)n_user_local * nnz_per_row;
  auto bail = [&](int code, const char* m) { b200als_destroy(s); return fail(code, m); };
  if (A.ptr.ensure(sizeof(int32_t) * ((size_t)n_user_local + 1)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc ptr");
  if (A.idx.ensure(sizeof(int32_t) * (size_t)A.nnz) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc idx");
  if (A.val.ensure(sizeof(float) * (size_t)A.nnz) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc val");
  const long long total = std::max<long long>(A.nnz, n_user_local + 1);
  synth_csr_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(n_user_local, n_item, nnz_per_row, seed,
                                                                         s->opt.feedback == B200ALS_EXPLICIT, user_offset,
                                                                         A.ptr.i32(), A.idx.i32(), A.val.f32());
  LAUNCHED(); if (cudaGetLastError() != cudaSuccess) return bail(B200ALS_ECUDA, "synth kernel launch");
  s->has[B200ALS_USERS] = true;
  s->shard_begin[B200ALS_USERS] = (int32_t)user_offset;
  s->shard_end[B200ALS_USERS] = (int32_t)user_offset + n_user_local;
  s->nnz_global[B200ALS_USERS] = (int64_t)n_user_global * nnz_per_row;
  rc = session_counts(s);
  if (rc == B200ALS_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(B200ALS_ECUDA, "sync");
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  *out = s;
  return B200ALS_OK;
}
