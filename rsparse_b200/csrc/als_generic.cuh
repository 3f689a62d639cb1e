// als_generic.cuh -- the shape-agnostic ALS half-iteration kernels (any rank <= 32*KPL, any row
// length, T = float or double).  These are the faithful restatement of the reference's per-column
// solve on the GPU and the fallback for rows / ranks the register-resident kernel does not cover:
//   als_cg_generic_kernel    warp per row, fixed-step CG, factor rows re-gathered from L2/HBM per pass
//       reference: cg_solver_implicit inst/include/wrmf_implicit.hpp:8-32,
//                  cg_solver_explicit inst/include/wrmf_explicit.hpp:8-31,
//                  loop body         wrmf_implicit.hpp:175-282 / wrmf_explicit.hpp:71-146
//   als_chol_generic_kernel  CTA per row, per-row Gram + Cholesky in shared memory
//       reference: wrmf_implicit.hpp:207-208,231,236 / wrmf_explicit.hpp:103-108
//       the same kernel runs solver = "nnls": c_nnls / scd_ls_update, inst/include/nnls.hpp:10-48, on the system it
//       has just assembled (wrmf_implicit.hpp:233-234, wrmf_explicit.hpp:110)
// Scheduling is a persistent grid with an atomic row ticket -- the GPU analogue of the reference's
// `omp for schedule(dynamic)` (wrmf_implicit.hpp:172-174).
#pragma once
#include "common.cuh"

namespace b200als {

template <typename T>
struct SolveParams {
  const int32_t* ptr;  // [n_targets + 1]
  const int32_t* idx;  // [nnz]
  const T* val;        // [nnz] (converted from R's double once at upload: wrmf_implicit.hpp:182-183)
  const T* X;          // rank x n_src   (fixed factor matrix)
  T* Y;                // rank x n_targets (solved in place)
  const T* G;          // rank x rank, XtX + lambda*I (implicit) ; nullptr for explicit
  const T* diag;       // [rank] CG only: X and Y are stored in the eigenbasis of XtX (eig.cuh), XtX v = diag (.) v; G unused
  int k;
  int n_targets;
  int feedback;        // 0 implicit, 1 explicit
  int cg_steps;
  int dynamic_lambda;
  int solver;          // als_chol_generic_kernel only: 0 Cholesky, 2 sequential coordinate-wise NNLS (wrmf.hpp:16-18)
  double lambda;
  const int32_t* row_list;  // optional subset of rows to solve
  int n_list;
  const int* n_list_dev;    // when non-null the list length is read from device memory (pipelined calls)
  int ptr_base;             // ptr[] values are offsets into a buffer that starts at this absolute offset
  int row_begin;            // without a row list: solve rows [row_begin, row_begin + n_targets)
  unsigned long long* ticket;
  double* loss_partials;  // [gridDim.x]
  int* status;            // != 0: some system was not positive definite
  // ---- bias terms (with_user_item_bias / with_global_bias).  The engine hands these kernels *compact* matrices:
  // X without its bias row, Y = the rows being solved, both k wide (see stateless_half, engine_stateless.inl) ----
  const T* xbias;         // [n_src] x_biases (wrmf_implicit.hpp:115-119, wrmf_explicit.hpp:58-63) or nullptr
  const T* rhs_init;      // [k] implicit only: rhs_init (wrmf_implicit.hpp:143-157) or nullptr
  T gbias;                // global_bias (implicit), 0 when unused
  T one_minus_g;          // (T)(1 - global_bias): the target of the implicit loss (wrmf_implicit.hpp:259-270)
  int solve_empty;        // implicit with biases / global bias: rows without entries are solved too (:179)
};

enum PassMode { kR0Implicit = 0, kApImplicit = 1, kR0Explicit = 2, kApExplicit = 3, kLossImplicit = 4, kLossExplicit = 5 };

// One fused sweep over the row's gathered factor rows x_j:
//   u_j = x_j . vec ;  w_j = f(c_j, u_j) ;  acc += w_j x_j        (modes 0-3)
//   loss += c_j (1 - u_j)^2  or  (c_j - u_j)^2                    (modes 4-5, acc untouched)
template <typename T, int KPL>
__device__ __forceinline__ T fused_pass(const SolveParams<T>& P, int p1, int n, const T (&vec)[KPL], T (&acc)[KPL],
                                        int mode) {
  const int lane = lane_id();
  const int k = P.k;
  T loss = T(0);
#pragma unroll
  for (int e = 0; e < KPL; e++) acc[e] = T(0);
  for (int base = 0; base < n; base += 32) {
    const int cnt = min(32, n - base);
    int my_idx = 0;
    T my_c = T(0), my_b = T(0);
    if (lane < cnt) {
      my_idx = __ldg(P.idx + p1 + base + lane);
      my_c = __ldg(P.val + p1 + base + lane);
      if (P.xbias) my_b = __ldg(P.xbias + my_idx);
    }
    for (int q = 0; q < cnt; q += 4) {
      T xr[4][KPL];
      T cj[4], bj[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = __shfl_sync(kFull, my_idx, q + u);
        cj[u] = __shfl_sync(kFull, my_c, q + u);
        bj[u] = __shfl_sync(kFull, my_b, q + u);
        ok[u] = (q + u) < cnt;
        const T* xj = P.X + (size_t)j * k;
#pragma unroll
        for (int e = 0; e < KPL; e++) {
          const int f = e * 32 + lane;
          xr[u][e] = (f < k) ? __ldg(xj + f) : T(0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        T d = T(0);
#pragma unroll
        for (int e = 0; e < KPL; e++) d += xr[u][e] * vec[e];
        d = warp_sum(d);
        T w;
        switch (mode) {
          // bj = gbias = 0 without bias terms: adding them is exact, the plain formulas of :16 / :15 result
          case kR0Implicit: w = cj[u] - (cj[u] - T(1)) * ((d + bj[u]) + P.gbias); break;   // wrmf_implicit.hpp:46,72
          case kApImplicit: w = (cj[u] - T(1)) * d; break;
          case kR0Explicit: w = (cj[u] - bj[u]) - d; break;                                // wrmf_explicit.hpp:89
          case kApExplicit: w = d; break;
          case kLossImplicit: { const T t = (P.one_minus_g - d) - bj[u]; w = T(0); if (ok[u]) loss += t * t * cj[u]; } break;
          default: { const T t = (cj[u] - bj[u]) - d; w = T(0); if (ok[u]) loss += t * t; } break;
        }
        if (!ok[u]) w = T(0);
#pragma unroll
        for (int e = 0; e < KPL; e++) acc[e] += w * xr[u][e];
      }
    }
  }
  return loss;
}

// out = G * v, G symmetric rank x rank in global memory (L1/L2 resident), v distributed over lanes.
template <typename T, int KPL>
__device__ __forceinline__ void gemv_sym(const T* __restrict__ G, int k, const T (&v)[KPL], T (&out)[KPL]) {
  const int lane = lane_id();
#pragma unroll
  for (int e = 0; e < KPL; e++) out[e] = T(0);
#pragma unroll
  for (int jj = 0; jj < KPL; jj++) {
    const int lim = min(32, k - jj * 32);
    for (int l = 0; l < lim; l++) {
      const T vj = __shfl_sync(kFull, v[jj], l);
      const T* g = G + (size_t)(jj * 32 + l) * k;
#pragma unroll
      for (int e = 0; e < KPL; e++) {
        const int f = e * 32 + lane;
        if (f < k) out[e] += __ldg(g + f) * vj;
      }
    }
  }
}

template <typename T, int KPL>
__device__ __forceinline__ T dot_lanes(const T (&a)[KPL], const T (&b)[KPL]) {
  T d = T(0);
#pragma unroll
  for (int e = 0; e < KPL; e++) d += a[e] * b[e];
  return warp_sum(d);
}

template <typename T, int KPL>
__global__ void __launch_bounds__(256) als_cg_generic_kernel(SolveParams<T> P) {
  __shared__ double s_red[32];
  const int lane = lane_id();
  const int k = P.k;
  const bool implicit = (P.feedback == 0);
  const unsigned long long total = P.n_list_dev ? (unsigned long long)__ldg(P.n_list_dev)
                                   : (P.row_list ? (unsigned long long)P.n_list : (unsigned long long)P.n_targets);
  double warp_loss = 0.0;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(P.ticket, 1ULL);
    t = __shfl_sync(kFull, t, 0);
    if (t >= total) break;
    const int row = P.row_list ? P.row_list[t] : (int)t + P.row_begin;
    const int p1 = P.ptr[row] - P.ptr_base, p2 = P.ptr[row + 1] - P.ptr_base;
    T* y = P.Y + (size_t)row * k;
    if (p1 >= p2 && !P.solve_empty) {  // wrmf_implicit.hpp:281 / wrmf_explicit.hpp:144
#pragma unroll
      for (int e = 0; e < KPL; e++) {
        const int f = e * 32 + lane;
        if (f < k) y[f] = T(0);
      }
      continue;
    }
    const int n = p2 - p1;
    T x[KPL], r[KPL], p[KPL], Ap[KPL], acc[KPL];
#pragma unroll
    for (int e = 0; e < KPL; e++) {
      const int f = e * 32 + lane;
      x[e] = (f < k) ? y[f] : T(0);  // init = Y.col(i)
    }
    // lambda_use: wrmf_explicit.hpp:78 (computed in T)
    const T lam_use = implicit ? T(0) : (T)(P.lambda * (P.dynamic_lambda ? (double)static_cast<T>(n) : 1.));
    // r = X_nnz (c - (c-1) % X_nnz' x) - XtX x      (wrmf_implicit.hpp:16)
    // r = X_nnz (c - X_nnz' x) - lambda x            (wrmf_explicit.hpp:15)
    fused_pass<T, KPL>(P, p1, n, x, acc, implicit ? kR0Implicit : kR0Explicit);
    T dgv[KPL];
#pragma unroll
    for (int e = 0; e < KPL; e++) {
      const int f = e * 32 + lane;
      dgv[e] = (P.diag && f < k) ? __ldg(P.diag + f) : T(0);
    }
    if (implicit) {
      if (P.diag) {
#pragma unroll
        for (int e = 0; e < KPL; e++) Ap[e] = dgv[e] * x[e];
      } else {
        gemv_sym<T, KPL>(P.G, k, x, Ap);
      }
#pragma unroll
      for (int e = 0; e < KPL; e++) {
        r[e] = acc[e] - Ap[e];
        const int f = e * 32 + lane;
        if (P.rhs_init && f < k) r[e] += __ldg(P.rhs_init + f);   // + rhs_init / global_bias_base (:47, :73)
      }
    } else {
#pragma unroll
      for (int e = 0; e < KPL; e++) r[e] = acc[e] - lam_use * x[e];
    }
#pragma unroll
    for (int e = 0; e < KPL; e++) p[e] = r[e];
    double rsold = (double)dot_lanes<T, KPL>(r, r);
    // guard absent from the reference: rsold == 0 (row converged exactly) would give alpha = 0/0 = NaN
    const int n_steps = (rsold > 0.0) ? P.cg_steps : 0;
    for (int it = 0; it < n_steps; it++) {
      // Ap = XtX p + X_nnz ((c-1) % X_nnz' p)        (wrmf_implicit.hpp:22)
      // Ap = X_nnz (X_nnz' p) + lambda p             (wrmf_explicit.hpp:21)
      fused_pass<T, KPL>(P, p1, n, p, acc, implicit ? kApImplicit : kApExplicit);
      if (implicit) {
        if (P.diag) {
#pragma unroll
          for (int e = 0; e < KPL; e++) Ap[e] = dgv[e] * p[e];
        } else {
          gemv_sym<T, KPL>(P.G, k, p, Ap);
        }
#pragma unroll
        for (int e = 0; e < KPL; e++) Ap[e] += acc[e];
      } else {
#pragma unroll
        for (int e = 0; e < KPL; e++) Ap[e] = acc[e] + lam_use * p[e];
      }
      const double pAp = (double)dot_lanes<T, KPL>(p, Ap);
      const double alpha = (pAp != 0.0) ? rsold / pAp : 0.0;
      const T a = (T)alpha;
#pragma unroll
      for (int e = 0; e < KPL; e++) {
        x[e] += a * p[e];
        r[e] -= a * Ap[e];
      }
      const double rsnew = (double)dot_lanes<T, KPL>(r, r);
      if (rsnew < B200ALS_CG_TOL) break;
      const T b = (T)(rsnew / rsold);
#pragma unroll
      for (int e = 0; e < KPL; e++) p[e] = r[e] + p[e] * b;
      rsold = rsnew;
    }
#pragma unroll
    for (int e = 0; e < KPL; e++) {
      const int f = e * 32 + lane;
      if (f < k) y[f] = x[e];
    }
    // loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132)
    T l = fused_pass<T, KPL>(P, p1, n, x, acc, implicit ? kLossImplicit : kLossExplicit);
    const T yy = dot_lanes<T, KPL>(x, x);
    l += (implicit ? (T)P.lambda : lam_use) * yy;
    warp_loss += (double)l;
  }
  // every lane of a warp carries the same warp_loss; keep lane 0's
  double v = (lane == 0) ? warp_loss : 0.0;
  const double tot = block_sum_double(v, s_red);
  if (threadIdx.x == 0) P.loss_partials[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------------
// Cholesky path: one CTA (256 threads as 16 x 16) per row.
//   shared: A[(k+1) x ks]  lower triangle of lhs with the rhs appended as row k (so the right-looking
//           factorisation also performs the forward substitution), tile[TN x k] staged factor rows.
// ---------------------------------------------------------------------------------------------------
constexpr int kCholTN = 16;

template <typename T>
__global__ void __launch_bounds__(256) als_chol_generic_kernel(SolveParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[32];
  __shared__ int s_row;
  __shared__ int s_fail;
  __shared__ double s_jit;
  const int k = P.k;
  const int ks = k + 1;  // odd-ish stride: breaks bank conflicts between rows
  T* A = reinterpret_cast<T*>(smem_raw);          // (k+1) * ks
  T* tile = A + (size_t)(k + 1) * ks;             // kCholTN * k
  T* colj = tile + (size_t)kCholTN * k;           // k + 1
  T* wts = colj + (k + 1);                        // kCholTN
  T* cs = wts + kCholTN;                          // kCholTN
  T* Bm = cs + kCholTN;                           // k * ks, NNLS only: XtX = lhs' * lhs
  T* muv = Bm + ((P.solver == 2) ? (size_t)k * ks : 0);  // k, NNLS only
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool implicit = (P.feedback == 0);
  const int total = P.n_list_dev ? __ldg(P.n_list_dev) : (P.row_list ? P.n_list : P.n_targets);
  double cta_loss = 0.0;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const unsigned long long t = atomicAdd(P.ticket, 1ULL);
      s_row = (t < (unsigned long long)total) ? (P.row_list ? P.row_list[t] : (int)t + P.row_begin) : -1;
      s_fail = 0;
    }
    __syncthreads();
    const int row = s_row;
    if (row < 0) break;
    const int p1 = P.ptr[row] - P.ptr_base, p2 = P.ptr[row + 1] - P.ptr_base;
    T* y = P.Y + (size_t)row * k;
    if (p1 >= p2 && !P.solve_empty) {
      for (int f = tid; f < k; f += 256) y[f] = T(0);
      continue;
    }
    const int n = max(p2 - p1, 0);
    const T lam_use = implicit ? T(0) : (T)(P.lambda * (P.dynamic_lambda ? (double)static_cast<T>(n) : 1.));
    // A singular / indefinite system (the reference's default lambda = 0 on explicit feedback makes every row with fewer
    // entries than the rank singular; its arma::solve(..., fast) then falls back to an approximate solution,
    // wrmf_explicit.hpp:108) is not an error of the whole call: the row is re-assembled and re-factored with a relative
    // diagonal shift (1e-6, then 1e-3 of the mean diagonal in fp32; 1e-12 / 1e-8 in fp64).  Only a row that still fails
    // sets the status word (B200ALS_ENOTSPD).  The shifted solution is NOT the reference's pseudo-inverse-like answer.
    T jitter = T(0);
    int failed = 0;
    for (int attempt = 0;; attempt++) {
    // lhs = XtX (already + lambda I)      wrmf_implicit.hpp:207 ; or lambda_use I   wrmf_explicit.hpp:103-104
    for (int a = ty; a <= k; a += 16)
      for (int b = tx; b < k; b += 16) {
        T v = T(0);
        if (a < k && b <= a) v = (implicit ? __ldg(P.G + (size_t)a * k + b) : ((a == b) ? lam_use : T(0))) + ((a == b) ? jitter : T(0));
        if (a == k && P.rhs_init) v = __ldg(P.rhs_init + b);   // rhs = rhs_init + ...   (wrmf_implicit.hpp:226,229)
        A[a * ks + b] = v;
      }
    __syncthreads();
    for (int base = 0; base < n; base += kCholTN) {
      const int cnt = min(kCholTN, n - base);
      for (int t = tid; t < cnt * k; t += 256) {
        const int j = t / k, f = t - j * k;
        tile[j * k + f] = __ldg(P.X + (size_t)__ldg(P.idx + p1 + base + j) * k + f);
      }
      if (tid < cnt) {
        const T c = __ldg(P.val + p1 + base + tid);
        const T bj = P.xbias ? __ldg(P.xbias + __ldg(P.idx + p1 + base + tid)) : T(0);
        // rhs weights: X_nnz c (:231) ; with biases X_nnz (c - x_biases % (c - 1)) (:226) ; explicit c - x_biases (:89)
        cs[tid] = implicit ? (P.xbias ? c - bj * (c - T(1)) : c) : (c - bj);
        wts[tid] = implicit ? (c - T(1)) : T(1);
      }
      __syncthreads();
      // lhs += X_nnz diag(w) X_nnz'  (lower triangle) ; rhs (row k) += X_nnz c
      for (int a = ty; a <= k; a += 16)
        for (int b = tx; b < k; b += 16) {
          if (a < k && b > a) continue;
          T s = T(0);
          if (a < k) {
            for (int j = 0; j < cnt; j++) s += wts[j] * tile[j * k + a] * tile[j * k + b];
          } else {
            for (int j = 0; j < cnt; j++) s += cs[j] * tile[j * k + b];
          }
          A[a * ks + b] += s;
        }
      __syncthreads();
    }
    if (P.solver == 2) {
      // ---- c_nnls (nnls.hpp:37-48): X = lhs (symmetric, lower triangle + mirror), y = rhs (row k of A) -----------
      auto lhs = [&](int a, int b) -> T { return (b <= a) ? A[a * ks + b] : A[b * ks + a]; };
      for (int a = ty; a < k; a += 16)
        for (int b = tx; b < k; b += 16) {
          T sacc = T(0);
          for (int m = 0; m < k; m++) sacc += lhs(m, a) * lhs(m, b);    // XtX = Xt * X
          if (a == b) sacc += (T)1e-16;                                // XtX.diag() += EPS
          Bm[a * ks + b] = sacc;
        }
      for (int f = tid; f < k; f += 256) colj[f] = y[f];                // res = initial = Y.col(i)
      __syncthreads();
      for (int a = tid; a < k; a += 256) {                               // mu = XtX * init - Xt * y
        T sacc = T(0);
        for (int m = 0; m < k; m++) sacc += Bm[a * ks + m] * colj[m] - lhs(m, a) * A[k * ks + m];
        muv[a] = sacc;
      }
      __syncthreads();
      // ---- scd_ls_update (nnls.hpp:10-34): sequential over coordinates, warp 0, mu updated lane-parallel -------
      if (tid < 32) {
        for (int t = 0; t < 10000; t++) {                               // SCD_MAX_ITER
          T rel_diff = T(0);
          for (int c = 0; c < k; c++) {
            const T old_value = colj[c];
            T new_value = old_value - muv[c] / Bm[c * ks + c];
            if (new_value < T(0)) new_value = T(0);
            const T diff = new_value - old_value;
            __syncwarp();
            if (diff != T(0)) {                                         // same value in every lane
              if (tid == 0) colj[c] = new_value;
              for (int l = tid; l < k; l += 32) muv[l] += diff * Bm[c * ks + l];   // XtX.unsafe_col(c), symmetric
              const T step_err = (T)(fabs((double)diff) / (fabs((double)old_value) + 1e-16));
              if (step_err > rel_diff) rel_diff = step_err;
            }
            __syncwarp();
          }
          if (rel_diff <= (T)1e-4) break;                                // SCD_TOL
        }
      }
      __syncthreads();
      break;
    }
    // right-looking Cholesky on rows 0..k (row k = rhs => ends up holding z = L^-1 rhs)
    for (int j = 0; j < k; j++) {
      const T d = A[j * ks + j];
      if (!(d > T(0))) {
        if (tid == 0) s_fail = 1;
        break;
      }
      const T inv = T(1) / sqrt(d);
      for (int a = j + 1 + tid; a <= k; a += 256) colj[a] = A[a * ks + j] * inv;
      __syncthreads();
      if (tid == 0) A[j * ks + j] = sqrt(d);
      for (int a = j + 1 + tid; a <= k; a += 256) A[a * ks + j] = colj[a];
      for (int a = j + 1 + ty; a <= k; a += 16) {
        const T ca = colj[a];
        const int bmax = (a < k) ? a : (k - 1);
        for (int b = j + 1 + tx; b <= bmax; b += 16) A[a * ks + b] -= ca * colj[b];
      }
      __syncthreads();
    }
    __syncthreads();
    failed = s_fail;
    __syncthreads();
    if (failed && attempt < 2) {
      // mean diagonal of the unshifted system = (sum of the gathered rows' weighted squared norms + tr(XtX) or k lambda) / k;
      // recomputed cheaply from the data that is still at hand: the first pivot-free quantity is the rhs-independent trace
      if (tid == 0) s_fail = 0;
      T tr = T(0);
      for (int j = warp_id(); j < n; j += 8) {
        const int src = __ldg(P.idx + p1 + j);
        const T* xj = P.X + (size_t)src * k;
        T d = T(0);
        for (int f = lane_id(); f < k; f += 32) { const T v = __ldg(xj + f); d += v * v; }
        d = warp_sum(d);
        const T cc = __ldg(P.val + p1 + j);
        if (lane_id() == 0) tr += d * (implicit ? (cc - T(1)) : T(1));
      }
      double trd = block_sum_double((double)tr, s_red);
      if (tid == 0) {
        if (implicit) for (int a = 0; a < k; a++) trd += (double)__ldg(P.G + (size_t)a * k + a);
        else trd += (double)lam_use * k;
        s_jit = fabs(trd) / k;
      }
      __syncthreads();
      const double rel = (sizeof(T) == 4) ? (attempt == 0 ? 1e-6 : 1e-3) : (attempt == 0 ? 1e-12 : 1e-8);
      jitter = (T)fmax(s_jit * rel, (sizeof(T) == 4) ? 1e-30 : 1e-200);
      continue;
    }
    if (failed && tid == 0) atomicExch(P.status, 1);
    break;
    }  // attempts
    if (failed) continue;   // leave Y untouched for this row; status reports B200ALS_ENOTSPD
    if (P.solver != 2) {
    // back substitution L' y = z by warp 0 (z = row k of A); result into colj[0..k)
    if (tid < 32) {
      for (int f = tid; f < k; f += 32) colj[f] = A[k * ks + f];
      __syncwarp();
      for (int i = k - 1; i >= 0; i--) {
        const T yi = colj[i] / A[i * ks + i];
        __syncwarp();
        if (tid == 0) colj[i] = yi;
        for (int l = tid; l < i; l += 32) colj[l] -= A[i * ks + l] * yi;
        __syncwarp();
      }
    }
    __syncthreads();
    }
    for (int f = tid; f < k; f += 256) y[f] = colj[f];
    // loss: warp per gathered row
    T l = T(0);
    for (int j = warp_id(); j < n; j += 8) {
      const int src = __ldg(P.idx + p1 + j);
      const T* xj = P.X + (size_t)src * k;
      T d = T(0);
      for (int f = lane_id(); f < k; f += 32) d += __ldg(xj + f) * colj[f];
      d = warp_sum(d);
      const T c = __ldg(P.val + p1 + j);
      const T bj = P.xbias ? __ldg(P.xbias + src) : T(0);
      const T t = implicit ? ((P.one_minus_g - d) - bj) : ((c - bj) - d);
      if (lane_id() == 0) l += implicit ? t * t * c : t * t;
    }
    if (warp_id() == 0) {
      T yy = T(0);
      for (int f = lane_id(); f < k; f += 32) yy += colj[f] * colj[f];
      yy = warp_sum(yy);
      if (lane_id() == 0) l += (implicit ? (T)P.lambda : lam_use) * yy;
    }
    cta_loss += block_sum_double((double)l, s_red);
  }
  if (tid == 0) P.loss_partials[blockIdx.x] = cta_loss;
}

template <typename T>
inline size_t chol_generic_smem_bytes(int k, int solver = 0) {
  const size_t nnls = (solver == 2) ? ((size_t)k * (k + 1) + k) : 0;
  return sizeof(T) * ((size_t)(k + 1) * (k + 1) + (size_t)kCholTN * k + (k + 1) + 2 * kCholTN + nnls) + 16;
}

}  // namespace b200als
