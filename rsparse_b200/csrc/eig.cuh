// eig.cuh -- change of basis that turns the k x k product XtX * p of the implicit-feedback CG
// (reference: inst/include/wrmf_implicit.hpp:16,22 -- `XtX * x`, `XtX * p`, 2k^2 flop per CG step and
// 44 % of the arithmetic at k = 128) into a 4-FMA diagonal scale.
//
// XtX = Q diag(d) Q' (symmetric eigendecomposition).  With X~ = X Q and y~ = Q' y the per-row system
//   (XtX + X_nnz D X_nnz') y = X_nnz c      becomes      (diag(d) + X~_nnz D X~_nnz') y~ = X~_nnz c
// and every CG iterate of the rotated system is Q' times the iterate of the original one (r~ = Q' r,
// p~ = Q' p; alpha, beta, the rsnew < CG_TOL test and the loss are rotation invariant), so the solver
// can run entirely in the rotated coordinates.  The engine therefore keeps both factor matrices in a
// common orthonormal basis B (true = stored * B'), re-diagonalises once per implicit CG half-iteration
// and rotates back only when factors leave the device.
//
//   jacobi_eig_kernel   one CTA, parallel cyclic two-sided Jacobi in fp64 on the k x k Gram (k <= 256)
//   rotate_rows_kernel  Z = X * R for n x 128 row blocks (fp32 FMA, R in shared memory), in place
#pragma once
#include "common.cuh"

namespace b200als {

constexpr int kJacobiThreads = 1024;

// A: k x k symmetric (row-major, overwritten: ends diagonal), Vt: k x k, row e = eigenvector e.
// Outputs: Qf[f*k + e] = V[f][e] (fp32), df[e] = eigenvalue (fp32), Q64 likewise in double.
// When k*k doubles fit (k <= 160) the working copy of the Gram lives in shared memory: its column
// rotations are strided accesses, which cost an L2 round trip each from global memory.
__global__ void __launch_bounds__(kJacobiThreads) jacobi_eig_kernel(double* __restrict__ Ag, double* __restrict__ Vt,
                                                                    int k, float* __restrict__ Qf,
                                                                    float* __restrict__ df, double* __restrict__ Q64,
                                                                    int max_sweeps, int a_in_smem) {
  extern __shared__ __align__(16) unsigned char jacobi_smem[];
  double* A = a_in_smem ? reinterpret_cast<double*>(jacobi_smem) : Ag;
  const int lda = a_in_smem ? (k + 1) : k;   // odd leading dimension: the column rotations are bank-conflict free
  __shared__ double s_c[128], s_s[128];
  __shared__ int s_p[128], s_q[128];
  __shared__ double s_red[32];
  __shared__ double s_off, s_diag;
  const int tid = threadIdx.x;
  const int np = (k + 1) / 2;       // pairs per step
  const int npad = 2 * np;          // even number of players
  for (int e = tid; e < k * k; e += kJacobiThreads) {
    Vt[e] = ((e / k) == (e % k)) ? 1.0 : 0.0;
    if (a_in_smem) A[(e / k) * lda + (e % k)] = Ag[e];
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    // convergence: off(A)^2 <= 1e-26 * diag(A)^2 (off/diag <= 1e-13: far below the fp32 consumers' precision)
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < k * k; e += kJacobiThreads) {
      const double v = A[(e / k) * lda + (e % k)];
      if ((e / k) == (e % k)) dg += v * v; else off += v * v;
    }
    const double toff = block_sum_double(off, s_red);
    const double tdg = block_sum_double(dg, s_red);
    if (tid == 0) { s_off = toff; s_diag = tdg; }
    __syncthreads();
    if (s_off <= 1e-26 * s_diag) break;
    for (int step = 0; step < npad - 1; step++) {
      // rotation parameters for the np disjoint pairs of this step (round-robin tournament)
      if (tid < np) {
        int p, q;
        if (tid == 0) { p = npad - 1; q = step; }
        else { p = (step + tid) % (npad - 1); q = (step - tid + (npad - 1)) % (npad - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < k) {
          const double apq = A[p * lda + q];
          if (apq != 0.0) {
            const double tau = (A[q * lda + q] - A[p * lda + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = -1;  // padding player
        }
        s_p[tid] = p; s_q[tid] = q; s_c[tid] = c; s_s[tid] = s;
      }
      __syncthreads();
      // A <- A J  (columns p, q of every row)
      for (int item = tid; item < np * k; item += kJacobiThreads) {
        const int pr = item / k, i = item - pr * k;
        const int p = s_p[pr], q = s_q[pr];
        if (q < 0) continue;
        const double c = s_c[pr], s = s_s[pr];
        const double ap = A[i * lda + p], aq = A[i * lda + q];
        A[i * lda + p] = c * ap - s * aq;
        A[i * lda + q] = s * ap + c * aq;
      }
      __syncthreads();
      // A <- J' A (rows p, q) ;  Vt <- J' Vt  (i.e. V <- V J)
      for (int item = tid; item < np * k; item += kJacobiThreads) {
        const int pr = item / k, j = item - pr * k;
        const int p = s_p[pr], q = s_q[pr];
        if (q < 0) continue;
        const double c = s_c[pr], s = s_s[pr];
        const double ap = A[p * lda + j], aq = A[q * lda + j];
        A[p * lda + j] = c * ap - s * aq;
        A[q * lda + j] = s * ap + c * aq;
        const double vp = Vt[p * k + j], vq = Vt[q * k + j];
        Vt[p * k + j] = c * vp - s * vq;
        Vt[q * k + j] = s * vp + c * vq;
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < k * k; e += kJacobiThreads) {
    const int f = e / k, ev = e % k;
    const double v = Vt[ev * k + f];
    Qf[e] = (float)v;
    if (Q64) Q64[e] = v;
  }
  for (int e = tid; e < k; e += kJacobiThreads) df[e] = (float)A[e * lda + e];
}

// C = A * B (k x k, row-major, double) -- basis bookkeeping B <- B Q ; trivially small.
// transA != 0: C = A' * B
__global__ void matmul_kk_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int k, int transA = 0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  const int i = e / k, j = e % k;
  double s = 0.0;
  if (transA) for (int l = 0; l < k; l++) s += A[l * k + i] * B[l * k + j];
  else for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * k + j];
  C[e] = s;
}
// out(fp32)[i][j] = transpose ? in[j][i] : in[i][j]
__global__ void convert_kk_kernel(const double* __restrict__ in, float* __restrict__ out, int k, int transpose) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  const int i = e / k, j = e % k;
  out[e] = (float)(transpose ? in[j * k + i] : in[e]);
}
__global__ void set_identity_kernel(double* __restrict__ B, int k) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  B[e] = ((e / k) == (e % k)) ? 1.0 : 0.0;
}

// Z[r][e] = sum_f X[r][f] R[f][e] for rows [0, n), k = 128, in place allowed (Z == X): each CTA stages its
// 64-row block in shared memory before writing.  256 threads, thread tile 4 rows x 8 columns.
constexpr int kRotK = 128;
constexpr int kRotRows = 64;
struct RotSmem {
  float R[kRotK][kRotK];           // 64 KB
  float Xt[kRotK][kRotRows + 4];   // the row block TRANSPOSED: Xt[f][r]; one 128-bit read = 4 rows of feature f
};
// X and Z may be the SAME buffer (in-place rotation): neither is __restrict__; a block reads its rows before writing them.
__global__ void __launch_bounds__(256) rotate_rows_kernel(const float* X, float* Z, const float* __restrict__ R, long long n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RotSmem& S = *reinterpret_cast<RotSmem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int t = tid; t < kRotK * kRotK / 4; t += 256)
    reinterpret_cast<float4*>(&S.R[0][0])[t] = __ldg(reinterpret_cast<const float4*>(R) + t);
  const int tx = tid & 15, ty = tid >> 4;  // columns {4tx..4tx+3} u {64+4tx..}, rows ty*4 .. ty*4+3
  const long long n_blocks = (n + kRotRows - 1) / kRotRows;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long r0 = blk * kRotRows;
    __syncthreads();
    for (int t = tid; t < kRotRows * (kRotK / 4); t += 256) {
      const int c4 = t / kRotRows, rr = t % kRotRows;   // consecutive threads -> consecutive rows: conflict-free stores
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + rr < n) v = *(reinterpret_cast<const float4*>(X + (size_t)(r0 + rr) * kRotK) + c4);   // plain load: Z may alias X
      S.Xt[c4 * 4 + 0][rr] = v.x;
      S.Xt[c4 * 4 + 1][rr] = v.y;
      S.Xt[c4 * 4 + 2][rr] = v.z;
      S.Xt[c4 * 4 + 3][rr] = v.w;
    }
    __syncthreads();
    float2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int f = 0; f < kRotK; f++) {
      const float4 b0 = *reinterpret_cast<const float4*>(&S.R[f][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&S.R[f][64 + tx * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&S.Xt[f][ty * 4]);
      const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 a2 = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(a2, bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const long long r = r0 + ty * 4 + i;
      if (r < n) {
        *reinterpret_cast<float4*>(Z + (size_t)r * kRotK + tx * 4) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
        *reinterpret_cast<float4*>(Z + (size_t)r * kRotK + 64 + tx * 4) = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
      }
    }
  }
}

// Z = X * R for any rank k <= KPAD with k % 4 == 0 (fp32 FFMA2; the rank-128 kernels above / rotate_tc.cuh are the
// fast paths).  One CTA of 256 threads per 64-row block: the block is staged transposed (Xt[f][r]) once, R streams
// through shared memory in chunks of 32 input features (from L2: k * k * 4 bytes per block), thread tile = 4 rows x
// (4 columns per 64-column block).  In place allowed (Z == X): a block reads all of its rows before it writes them.
template <int KPAD>
struct RotAnySmem {
  float Xt[KPAD][kRotRows + 4];
  float Rc[32][KPAD];
};
template <int KPAD>
__global__ void __launch_bounds__(256) rotate_any_kernel(const float* X, float* Z, const float* __restrict__ R, long long n, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RotAnySmem<KPAD>& S = *reinterpret_cast<RotAnySmem<KPAD>*>(smem_raw);
  constexpr int NB = (KPAD + 63) / 64;           // 64-column blocks per thread
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k4 = k / 4;
  const long long n_blocks = (n + kRotRows - 1) / kRotRows;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long r0 = blk * kRotRows;
    __syncthreads();
    for (int t = tid; t < kRotRows * k4; t += 256) {
      const int c4 = t / kRotRows, rr = t % kRotRows;   // consecutive threads -> consecutive rows: conflict-free stores
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + rr < n) v = *(reinterpret_cast<const float4*>(X + (size_t)(r0 + rr) * k) + c4);
      S.Xt[c4 * 4 + 0][rr] = v.x;
      S.Xt[c4 * 4 + 1][rr] = v.y;
      S.Xt[c4 * 4 + 2][rr] = v.z;
      S.Xt[c4 * 4 + 3][rr] = v.w;
    }
    float2 acc[4][NB][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int m = 0; m < NB; m++) acc[i][m][0] = acc[i][m][1] = make_float2(0.f, 0.f);
    for (int f0 = 0; f0 < k; f0 += 32) {
      const int fc = min(32, k - f0);
      __syncthreads();   // Xt staged (first chunk) / previous chunk of R consumed
      for (int t = tid; t < 32 * (KPAD / 4); t += 256) {
        const int ff = t / (KPAD / 4), c4 = t % (KPAD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ff < fc && c4 < k4) v = __ldg(reinterpret_cast<const float4*>(R + (size_t)(f0 + ff) * k) + c4);
        *reinterpret_cast<float4*>(&S.Rc[ff][c4 * 4]) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int ff = 0; ff < fc; ff++) {
        const float4 a4 = *reinterpret_cast<const float4*>(&S.Xt[f0 + ff][ty * 4]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int m = 0; m < NB; m++) {
          if (m * 64 + tx * 4 < KPAD) {
            const float4 b = *reinterpret_cast<const float4*>(&S.Rc[ff][m * 64 + tx * 4]);
            const float2 b0 = make_float2(b.x, b.y), b1 = make_float2(b.z, b.w);
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float2 a2 = make_float2(av[i], av[i]);
              acc[i][m][0] = __ffma2_rn(a2, b0, acc[i][m][0]);
              acc[i][m][1] = __ffma2_rn(a2, b1, acc[i][m][1]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const long long r = r0 + ty * 4 + i;
      if (r < n) {
#pragma unroll
        for (int m = 0; m < NB; m++) {
          const int col = m * 64 + tx * 4;
          if (col < k)
            *reinterpret_cast<float4*>(Z + (size_t)r * k + col) = make_float4(acc[i][m][0].x, acc[i][m][0].y, acc[i][m][1].x, acc[i][m][1].y);
        }
      }
    }
  }
}

}  // namespace b200als
