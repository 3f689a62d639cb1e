// als_chol_tile.cuh -- Cholesky branch of the half-iteration for the common shapes (rank 64 / 128, fp32,
// rows with 1..80 non-zeros); rows outside that go to als_chol_generic_kernel.
// Reference: lhs = XtX + X_nnz diag(c-1) X_nnz', rhs = X_nnz c, solve(lhs, rhs)   (wrmf_implicit.hpp:207-236)
//            lhs = X_nnz X_nnz' + lambda_u I, rhs = X_nnz r, solve(lhs, rhs)      (wrmf_explicit.hpp:103-108)
// One CTA of 256 threads per row (static interleave over a row list), several CTAs per SM:
//   1. the gathered tile X_nnz (n x K) is staged once into shared memory with 16-byte cp.async;
//   2. Gram: every thread owns a (K/16) x (K/16) register block of the K x K system (interleaved 4-wide groups so
//      both operand reads are conflict-free 128-bit loads), one rank-1 update per gathered row, packed FFMA2;
//      the rhs rides along as an extra row;  XtX (or lambda_u I) is added when the block is written to smem;
//   3. right-looking Cholesky on the unscaled matrix: A[a][b] -= A[a][j] A[b][j] / d_j for j < b <= a -- column j is
//      only read, so ONE __syncthreads per column; the appended rhs row is forward-substituted for free;
//   4. back substitution by one warp, loss by warp-per-gathered-row dots on the staged tile.
// Algorithmic work per row: 2nK^2 (Gram, computed symmetric-full = 2x) + K^3/3 flop; bytes as the CG path.
#pragma once
#include "als_generic.cuh"

namespace b200als {

constexpr int kCholMaxN = 80;

template <int K>
struct alignas(16) CholTileSmem {
  static constexpr int LDA = K + 1;     // odd: column reads of the factorisation are conflict-free
  float A[(K + 1) * LDA];               // lower triangle (+ rhs in row K)
  alignas(16) float tile[kCholMaxN * K];   // 16-byte cp.async destinations / float4 reads
  float cs[kCholMaxN], ws[kCholMaxN];
  int idx[kCholMaxN];
  float rs[K];                          // 1 / sqrt(d_j)
  alignas(16) float zz[K];
  alignas(8) double red[32];
  int fail;
};

template <int K>
__global__ void __launch_bounds__(256) als_chol_tile_kernel(SolveParams<float> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = CholTileSmem<K>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDA = SM::LDA;
  constexpr int G = K / 64;            // 4-wide groups per thread and dimension (1 at K = 64, 2 at K = 128)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const bool implicit = (P.feedback == 0);
  const int total = P.n_list_dev ? __ldg(P.n_list_dev) : P.n_list;
  double cta_loss = 0.0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int row = P.row_list ? __ldg(P.row_list + t) : t + P.row_begin;
    const int p1 = __ldg(P.ptr + row) - P.ptr_base, n = __ldg(P.ptr + row + 1) - P.ptr_base - p1;
    __syncthreads();   // previous row fully consumed
    if (tid < n) {
      S.idx[tid] = __ldg(P.idx + p1 + tid);
      const float c = __ldg(P.val + p1 + tid);
      S.cs[tid] = c;
      S.ws[tid] = implicit ? (c - 1.0f) : 1.0f;
    }
    if (tid == 0) S.fail = 0;
    __syncthreads();
    for (int e = tid; e < n * (K / 4); e += 256) {
      const int j = e / (K / 4), c4 = e - j * (K / 4);
      cp_async_16(&S.tile[j * K + c4 * 4], P.X + (size_t)S.idx[j] * K + c4 * 4);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- Gram + rhs ----------------------------------------------------------------------------------
    float2 acc[G * 4][G * 2];   // rows: group gi, element i ; cols: group gj, pairs
    float racc[G * 4];
#pragma unroll
    for (int i = 0; i < G * 4; i++) {
      racc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < G * 2; j++) acc[i][j] = make_float2(0.f, 0.f);
    }
    for (int j = 0; j < n; j++) {
      const float wj = S.ws[j], cj = S.cs[j];
      float a[G * 4];
      float2 b[G * 2];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float4 av = *reinterpret_cast<const float4*>(&S.tile[j * K + g * 64 + ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&S.tile[j * K + g * 64 + tx * 4]);
        a[g * 4 + 0] = av.x; a[g * 4 + 1] = av.y; a[g * 4 + 2] = av.z; a[g * 4 + 3] = av.w;
        b[g * 2 + 0] = make_float2(bv.x, bv.y);
        b[g * 2 + 1] = make_float2(bv.z, bv.w);
      }
#pragma unroll
      for (int i = 0; i < G * 4; i++) {
        const float aw = a[i] * wj;
        const float2 aw2 = make_float2(aw, aw);
#pragma unroll
        for (int jj = 0; jj < G * 2; jj++) acc[i][jj] = __ffma2_rn(aw2, b[jj], acc[i][jj]);
        racc[i] = fmaf(cj, a[i], racc[i]);
      }
    }
    // ---- block -> shared memory (+ XtX, or lambda_u on the diagonal) ---------------------------------------
    const float lam_use = implicit ? 0.0f : (float)(P.lambda * (P.dynamic_lambda ? (double)(float)n : 1.));
#pragma unroll
    for (int i = 0; i < G * 4; i++) {
      const int ar = (i / 4) * 64 + ty * 4 + (i % 4);
#pragma unroll
      for (int jj = 0; jj < G * 2; jj++) {
        const int bc = (jj / 2) * 64 + tx * 4 + (jj % 2) * 2;
        float v0 = acc[i][jj].x, v1 = acc[i][jj].y;
        if (implicit) {
          const float2 g = __ldg(reinterpret_cast<const float2*>(P.G + (size_t)ar * K + bc));
          v0 += g.x; v1 += g.y;
        } else {
          if (ar == bc) v0 += lam_use;
          if (ar == bc + 1) v1 += lam_use;
        }
        S.A[ar * LDA + bc] = v0;
        S.A[ar * LDA + bc + 1] = v1;
      }
      if (tx == 0) S.A[K * LDA + ar] = racc[i];
    }
    // ---- right-looking Cholesky, one barrier per column ---------------------------------------------------
    for (int j = 0; j < K; j++) {
      __syncthreads();
      const float d = S.A[j * LDA + j];
      if (!(d > 0.0f)) {
        if (tid == 0) { S.fail = 1; atomicExch(P.status, 1); }
        break;
      }
      const float dinv = 1.0f / d;
      if (tid == 0) S.rs[j] = 1.0f / sqrtf(d);
      for (int a = j + 1 + ty; a <= K; a += 16) {
        const float aj = S.A[a * LDA + j] * dinv;
        const int bmax = (a < K) ? a : (K - 1);
        for (int b = j + 1 + tx; b <= bmax; b += 16) S.A[a * LDA + b] = fmaf(-aj, S.A[b * LDA + j], S.A[a * LDA + b]);
      }
    }
    __syncthreads();
    if (S.fail) continue;   // Y row untouched; status reports B200ALS_ENOTSPD
    // ---- back substitution (warp 0): L = A[:, j] * rs[j], z_j = A[K][j] * rs[j] --------------------------------
    if (warp == 0) {
      for (int f = lane; f < K; f += 32) S.zz[f] = S.A[K * LDA + f] * S.rs[f];
      __syncwarp();
      for (int i = K - 1; i >= 0; i--) {
        const float yi = S.zz[i] * S.rs[i];
        __syncwarp();
        if (lane == 0) S.zz[i] = yi;
        for (int l = lane; l < i; l += 32) S.zz[l] = fmaf(-S.A[i * LDA + l] * S.rs[l], yi, S.zz[l]);
        __syncwarp();
      }
    }
    __syncthreads();
    float* y = P.Y + (size_t)row * K;
    if (tid < K / 4) *reinterpret_cast<float4*>(y + tid * 4) = *reinterpret_cast<const float4*>(&S.zz[tid * 4]);
    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132) on the staged tile --------------------
    float l = 0.0f;
    for (int j = warp; j < n; j += 8) {
      float dsum = 0.0f;
      for (int f = lane; f < K; f += 32) dsum = fmaf(S.tile[j * K + f], S.zz[f], dsum);
      dsum = warp_sum(dsum);
      const float c = S.cs[j];
      const float tt = implicit ? (1.0f - dsum) : (c - dsum);
      if (lane == 0) l += implicit ? tt * tt * c : tt * tt;
    }
    if (warp == 0) {
      float yy = 0.0f;
      for (int f = lane; f < K; f += 32) yy = fmaf(S.zz[f], S.zz[f], yy);
      yy = warp_sum(yy);
      if (lane == 0) l += (implicit ? (float)P.lambda : lam_use) * yy;
    }
    cta_loss += block_sum_double((double)l, S.red);
  }
  if (tid == 0) P.loss_partials[blockIdx.x] = cta_loss;
}

}  // namespace b200als
