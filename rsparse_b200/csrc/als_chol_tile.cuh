// als_chol_tile.cuh -- Cholesky branch of the half-iteration for the common shapes (rank 64 / 128, fp32,
// rows with 1..80 non-zeros); rows outside that go to als_chol_generic_kernel.
// Reference: lhs = XtX + X_nnz diag(c-1) X_nnz', rhs = X_nnz c, solve(lhs, rhs)   (wrmf_implicit.hpp:207-236)
//            lhs = X_nnz X_nnz' + lambda_u I, rhs = X_nnz r, solve(lhs, rhs)      (wrmf_explicit.hpp:103-108)
// One CTA of 160 threads per row (static interleave over a row list), several CTAs per SM.  Only the 136 lower-
// triangular (K/16) x (K/16) blocks of the symmetric system are kept (one per thread, triangular index decode), so
// no issue slot is spent on mirror blocks:
//   1. the gathered tile X_nnz (n x K) is staged once into shared memory with 16-byte cp.async;
//   2. Gram: every thread owns a (K/16) x (K/16) register block of the K x K system (interleaved 4-wide groups so
//      both operand reads are conflict-free 128-bit loads), one rank-1 update per gathered row, packed FFMA2;
//      the rhs rides along as an extra row;  XtX (or lambda_u I) is added when the block is written to smem;
//   3. right-looking Cholesky with the matrix STILL IN THOSE REGISTERS: at column j the owning threads publish the
//      (unscaled) column into row j of a K x K shared array, one __syncthreads, then every thread applies
//      A[a][b] -= A[a][j] A[b][j] / d_j to its whole register block (symmetric-full, rows/cols <= j masked to zero)
//      with packed FFMA2 -- no address arithmetic, no shared-memory read-modify-write; the rhs is carried along
//      (forward substitution for free);
//   4. back substitution by one warp on the published columns, loss by warp-per-gathered-row dots on the tile.
// Algorithmic work per row: 2nK^2 (Gram, computed symmetric-full = 2x) + K^3/3 flop; bytes as the CG path.
#pragma once
#include "als_generic.cuh"

namespace b200als {

constexpr int kCholMaxN = 80;
constexpr int kCholThreads = 160;   // 136 lower-triangular 16 x 16 block positions + 24 idle lanes
constexpr int kCholWarps = kCholThreads / 32;

template <int K>
struct alignas(16) CholTileSmem {
  static constexpr int LDT = K + 4;     // rows stay 16-byte aligned, consecutive rows are 4 banks apart
  alignas(16) float Lt[K * LDT];        // row j = unscaled column j of the factor (entries a > j), published at step j
  alignas(16) float tile[kCholMaxN * K];   // 16-byte cp.async destinations / float4 reads
  float cs[kCholMaxN], ws[kCholMaxN];
  int idx[kCholMaxN];
  float rs[K];                          // 1 / sqrt(d_j)
  float rj[2];                          // rhs entry of the current column (double buffered)
  alignas(16) float zz[K];
  alignas(8) double red[32];
  int fail;
};

template <int K>
__global__ void __launch_bounds__(kCholThreads, (K == 64) ? 5 : 2) als_chol_tile_kernel(SolveParams<float> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = CholTileSmem<K>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDT = SM::LDT;
  constexpr int G = K / 64;            // 4-wide groups per thread and dimension (1 at K = 64, 2 at K = 128)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // thread -> block (ty, tx) of the lower triangle, row-major: t = ty (ty + 1) / 2 + tx, tx <= ty < 16
  int ty = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
  while ((ty + 1) * (ty + 2) / 2 <= tid) ty++;
  while (ty * (ty + 1) / 2 > tid) ty--;
  const int tx = tid - ty * (ty + 1) / 2;
  const bool has_block = (tid < 136);
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
    for (int e = tid; e < n * (K / 4); e += kCholThreads) {
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
    for (int j = 0; j < (has_block ? n : 0); j++) {
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
    // ---- + XtX (or lambda_u on the diagonal), still in registers ------------------------------------------------
    const float lam_use = implicit ? 0.0f : (float)(P.lambda * (P.dynamic_lambda ? (double)(float)n : 1.));
#pragma unroll
    for (int i = 0; i < G * 4; i++) {
      const int ar = (i / 4) * 64 + (has_block ? ty : 0) * 4 + (i % 4);
#pragma unroll
      for (int jj = 0; jj < G * 2; jj++) {
        const int bc = (jj / 2) * 64 + (has_block ? tx : 0) * 4 + (jj % 2) * 2;
        if (!has_block) continue;
        if (implicit) {
          const float2 g = __ldg(reinterpret_cast<const float2*>(P.G + (size_t)ar * K + bc));
          acc[i][jj].x += g.x;
          acc[i][jj].y += g.y;
        } else {
          if (ar == bc) acc[i][jj].x += lam_use;
          if (ar == bc + 1) acc[i][jj].y += lam_use;
        }
      }
    }
    // ---- right-looking Cholesky on the register blocks, one barrier per column --------------------------------
    bool failed = false;
#pragma unroll
    for (int g = 0; g < G; g++) {
      for (int J = 0; J < 16 && !failed; J++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int j = g * 64 + J * 4 + c;
          if (has_block && tx == J) {   // owners of column j publish it (all their rows; readers mask rows <= j)
#pragma unroll
            for (int gi = 0; gi < G; gi++) {
              float4 col;
              col.x = (c % 2 == 0) ? acc[gi * 4 + 0][g * 2 + c / 2].x : acc[gi * 4 + 0][g * 2 + c / 2].y;
              col.y = (c % 2 == 0) ? acc[gi * 4 + 1][g * 2 + c / 2].x : acc[gi * 4 + 1][g * 2 + c / 2].y;
              col.z = (c % 2 == 0) ? acc[gi * 4 + 2][g * 2 + c / 2].x : acc[gi * 4 + 2][g * 2 + c / 2].y;
              col.w = (c % 2 == 0) ? acc[gi * 4 + 3][g * 2 + c / 2].x : acc[gi * 4 + 3][g * 2 + c / 2].y;
              *reinterpret_cast<float4*>(&S.Lt[j * LDT + gi * 64 + ty * 4]) = col;
            }
          }
          if (G > 1 && has_block && ty == J && tx < J) {
            // interleaved groups (K = 128): entries (a, j) with a in an EARLIER 16-block than j but a later group are
            // kept transposed, as row j of block (J, tx) -- publish those too (disjoint from the columns above)
#pragma unroll
            for (int gj = 0; gj < G; gj++) {
              const float2 r0 = acc[g * 4 + c][gj * 2 + 0], r1 = acc[g * 4 + c][gj * 2 + 1];
              *reinterpret_cast<float4*>(&S.Lt[j * LDT + gj * 64 + tx * 4]) = make_float4(r0.x, r0.y, r1.x, r1.y);
            }
          }
          if (has_block && tx == 0 && ty == J) S.rj[j & 1] = racc[g * 4 + c];
          __syncthreads();
          const float d = S.Lt[j * LDT + j];
          if (!(d > 0.0f)) {   // same value in every thread
            if (tid == 0) { S.fail = 1; atomicExch(P.status, 1); }
            failed = true;
            break;
          }
          const float dinv = 1.0f / d;
          const float rjv = S.rj[j & 1];
          if (tid == 0) {
            const float r = 1.0f / sqrtf(d);
            S.rs[j] = r;
            S.zz[j] = rjv * r;   // z_j = (L^-1 rhs)_j
          }
          // my block needs updating only if it has a row > j and a column > j
          const int max_row = (G - 1) * 64 + ty * 4 + 3, max_col = (G - 1) * 64 + tx * 4 + 3;
          if (has_block && max_row > j && (max_col > j || tx == 0)) {   // tx == 0 threads also carry the rhs row
            float ca[G * 4];
            float2 cb[G * 2];
#pragma unroll
            for (int gi = 0; gi < G; gi++) {
              const float4 av = *reinterpret_cast<const float4*>(&S.Lt[j * LDT + gi * 64 + ty * 4]);
              const float4 bv = *reinterpret_cast<const float4*>(&S.Lt[j * LDT + gi * 64 + tx * 4]);
              const int a0 = gi * 64 + ty * 4, b0 = gi * 64 + tx * 4;
              ca[gi * 4 + 0] = (a0 + 0 > j) ? av.x : 0.f; ca[gi * 4 + 1] = (a0 + 1 > j) ? av.y : 0.f;
              ca[gi * 4 + 2] = (a0 + 2 > j) ? av.z : 0.f; ca[gi * 4 + 3] = (a0 + 3 > j) ? av.w : 0.f;
              cb[gi * 2 + 0] = make_float2((b0 + 0 > j) ? bv.x : 0.f, (b0 + 1 > j) ? bv.y : 0.f);
              cb[gi * 2 + 1] = make_float2((b0 + 2 > j) ? bv.z : 0.f, (b0 + 3 > j) ? bv.w : 0.f);
            }
#pragma unroll
            for (int i = 0; i < G * 4; i++) {
              const float sneg = -ca[i] * dinv;
              const float2 s2 = make_float2(sneg, sneg);
#pragma unroll
              for (int jj = 0; jj < G * 2; jj++) acc[i][jj] = __ffma2_rn(s2, cb[jj], acc[i][jj]);
              racc[i] = fmaf(-rjv * dinv, ca[i], racc[i]);   // rhs row (meaningful in the tx == 0 threads)
            }
          }
        }
      }
    }
    __syncthreads();
    if (S.fail) continue;   // Y row untouched; status reports B200ALS_ENOTSPD
    // ---- blocked back substitution  L' y = z,  L[l][i] = Lt[i][l] * rs_i  (l > i) --------------------------------
    // 32-row blocks from the bottom: (1) all warps subtract the already solved part (row dots + shuffle reductions); (2) warp 0 solves the 32 x 32 triangle with one broadcast per step.
    for (int b0 = K - 32; b0 >= 0; b0 -= 32) {
      if (b0 + 32 < K) {
        for (int i = b0 + warp; i < b0 + 32; i += kCholWarps) {
          float part = 0.f;
          for (int l = b0 + 32 + lane; l < K; l += 32) part = fmaf(S.Lt[i * LDT + l], S.zz[l], part);
          part = warp_sum(part);
          if (lane == 0) S.zz[i] -= S.rs[i] * part;
        }
        __syncthreads();
      }
      if (warp == 0) {
        const int i = b0 + lane;
        const float ri = S.rs[i];
        float zi = S.zz[i];
#pragma unroll 8
        for (int sidx = 31; sidx >= 0; sidx--) {
          const float ys = __shfl_sync(kFull, zi * ri, sidx);            // y_s, final once step s is reached
          if (lane < sidx) zi = fmaf(-S.Lt[i * LDT + b0 + sidx] * ri, ys, zi);
        }
        S.zz[i] = zi * ri;
      }
      __syncthreads();
    }
    float* y = P.Y + (size_t)row * K;
    if (tid < K / 4) *reinterpret_cast<float4*>(y + tid * 4) = *reinterpret_cast<const float4*>(&S.zz[tid * 4]);
    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132) on the staged tile --------------------
    float l = 0.0f;
    for (int j = warp; j < n; j += kCholWarps) {
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
