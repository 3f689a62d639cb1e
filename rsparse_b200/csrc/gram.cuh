// gram.cuh -- XtX = tcrossprod(X) + lambda*I  (reference: R/model_WRMF.R:474-486 and :347-353, which
// call BLAS through R).  fp32 FMA parity path: each CTA owns one 128 x 128 output tile and a slice
// of the n rows; 256 threads hold an 8 x 8 register block each, rows are staged through shared memory
// 32 at a time with coalesced 128-bit loads, the per-thread float accumulators are flushed into a
// double accumulator every kFlushRows rows (bounds the fp32 summation length, like a blocked syrk),
// and a second kernel adds the per-CTA partials in a fixed order => bit-reproducible.
// Algorithmic work: 2*n*k^2 flop, n*k*sizeof(T) bytes read.
#pragma once
#include "common.cuh"

namespace b200als {

constexpr int kGramTile = 128;
constexpr int kGramRows = 32;
constexpr int kGramFlushRows = 512;

// partials: [gridDim.x][n_tiles][128*128] doubles; tile t covers (ta, tb) with tb <= ta
template <typename T>
__global__ void __launch_bounds__(256) gram_partial_kernel(const T* __restrict__ X, int k, long long n,
                                                           long long rows_per_cta, double* __restrict__ partials,
                                                           int n_tiles_1d) {
  constexpr int kRows = (sizeof(T) == 4) ? kGramRows : kGramRows / 2;  // 32 KB of static smem either way
  __shared__ __align__(16) T sa[kRows][kGramTile];
  __shared__ __align__(16) T sb[kRows][kGramTile];
  const int tile = blockIdx.y;
  // decode lower-triangular tile index
  int ta = 0, acc_t = 0;
  while (acc_t + ta + 1 <= tile) { acc_t += ta + 1; ta++; }
  const int tb = tile - acc_t;
  const int a0 = ta * kGramTile, b0 = tb * kGramTile;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = min(n, r_begin + rows_per_cta);
  // thread (ty, tx) owns output rows  {4ty..4ty+3} u {64+4ty..64+4ty+3}  and the same pattern of columns
  // for tx: both shared-memory operand reads are then conflict-free 128-bit loads.
  constexpr bool kIsF32 = (sizeof(T) == 4);
  T acc[8][8];
  double dacc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[i][j] = T(0); dacc[i][j] = 0.0; }
  int since_flush = 0;
  for (long long r0 = r_begin; r0 < r_end; r0 += kRows) {
    const int cnt = (int)min((long long)kRows, r_end - r0);
    __syncthreads();
    for (int t = tid; t < kRows * kGramTile; t += 256) {
      const int rr = t / kGramTile, c = t % kGramTile;
      T va = T(0), vb = T(0);
      if (rr < cnt) {
        const T* xr = X + (size_t)(r0 + rr) * k;
        if (a0 + c < k) va = __ldg(xr + a0 + c);
        if (b0 + c < k) vb = __ldg(xr + b0 + c);
      }
      sa[rr][c] = va;
      sb[rr][c] = vb;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < kRows; rr++) {
      T av[8], bv[8];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        av[i] = sa[rr][ty * 4 + i]; av[4 + i] = sa[rr][64 + ty * 4 + i];
        bv[i] = sb[rr][tx * 4 + i]; bv[4 + i] = sb[rr][64 + tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] += av[i] * bv[j];
    }
    if (kIsF32) {
      since_flush += kRows;
      if (since_flush >= kGramFlushRows) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) { dacc[i][j] += (double)acc[i][j]; acc[i][j] = T(0); }
        since_flush = 0;
      }
    }
  }
  double* out = partials + ((size_t)blockIdx.x * gridDim.y + tile) * (kGramTile * kGramTile);
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int a = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
      const int b = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
      out[a * kGramTile + b] = (kIsF32 ? dacc[i][j] : 0.0) + (double)acc[i][j];
    }
}

// G[a][b] = sum over CTAs (fixed order) + lambda on the diagonal; writes T copy and (optional) double copy.
template <typename T>
__global__ void gram_reduce_kernel(const double* __restrict__ partials, int n_cta, int n_tiles, int k, double lambda,
                                   T* __restrict__ G, double* __restrict__ G64) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  int a = e / k, b = e % k;
  const int aa = max(a, b), bb = min(a, b);  // lower-triangular tiles only; mirror
  const int ta = aa / kGramTile, tb = bb / kGramTile;
  const int tile = ta * (ta + 1) / 2 + tb;
  const int off = (aa % kGramTile) * kGramTile + (bb % kGramTile);
  double s = 0.0;
  for (int c = 0; c < n_cta; c++) s += partials[((size_t)c * n_tiles + tile) * (kGramTile * kGramTile) + off];
  if (a == b) s += lambda;
  G[e] = (T)s;
  if (G64) G64[e] = s;
}

}  // namespace b200als
