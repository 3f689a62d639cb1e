// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rsparse_b200/, the
// C-ABI library) may include, link or call this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or
// the timed CPU baseline.
//
// A dependency-free C++17/OpenMP restatement of the reference's WRMF ALS half-iteration,
// following the reference's loop structure line by line (citations relative to
// /root/reference/):
//   als_implicit<T>        inst/include/wrmf_implicit.hpp:90-305   (no-bias branches; als_implicit_bias: all branches)
//   cg_solver_implicit<T>  inst/include/wrmf_implicit.hpp:8-32
//   als_explicit<T>        inst/include/wrmf_explicit.hpp:33-174   (no-bias branches; als_explicit_bias: all branches)
//   initialize_biases<T>   inst/include/wrmf_utils.hpp:32-183
//   cg_solver_explicit<T>  inst/include/wrmf_explicit.hpp:8-31
//   c_nnls / scd_ls_update inst/include/nnls.hpp:10-48
//   XtX = tcrossprod(X)+lambda*I   R/model_WRMF.R:474-486
//   constants               inst/include/wrmf.hpp:14-22
// Armadillo expressions are replaced by hand-written dense loops in the same numeric type T;
// `rsold/rsnew/alpha` stay double as in wrmf_implicit.hpp:18.  OpenMP scheduling matches the
// reference: `parallel` + `for schedule(dynamic) reduction(+:loss)` (implicit, :162-174) and
// `parallel for schedule(dynamic, GRAIN_SIZE)` (explicit, wrmf_explicit.hpp:68-70).
//
// PINNING: this port is checked (tests/test_oracle.py) against golden vectors produced by the
// reference's *own* header files compiled unmodified against oracle/mini_arma
// (oracle/_ref/libref_wrmf.so, recipe oracle/build_ref.sh, generator tests/golden/make_golden.py).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GRAIN_SIZE 100              // wrmf.hpp:14
#define CHOLESKY 0                  // wrmf.hpp:16
#define CONJUGATE_GRADIENT 1        // wrmf.hpp:17
#define SEQ_COORDINATE_WISE_NNLS 2  // wrmf.hpp:18
#define SCD_MAX_ITER 10000          // wrmf.hpp:20
#define SCD_TOL 1e-4                // wrmf.hpp:21
#define CG_TOL 1e-10                // wrmf.hpp:22
#define NNLS_EPS 1e-16              // nnls.hpp:8

namespace {

template <class T>
inline T dotT(const T* a, const T* b, int n) {
  T acc = T(0);
#pragma omp simd reduction(+ : acc)
  for (int i = 0; i < n; i++) acc += a[i] * b[i];
  return acc;
}

// y = A * x   (A column-major k x k)
template <class T>
inline void symv(const T* A, const T* x, T* y, int k) {
  for (int i = 0; i < k; i++) y[i] = T(0);
  for (int j = 0; j < k; j++) {
    const T s = x[j];
    const T* a = A + (size_t)j * k;
#pragma omp simd
    for (int i = 0; i < k; i++) y[i] += a[i] * s;
  }
}

// u = Xnnz^T * x  (Xnnz is k x n column-major: column j = gathered factor row j)
template <class T>
inline void gemv_t(const T* Xn, const T* x, T* u, int k, int n) {
  for (int j = 0; j < n; j++) u[j] = dotT(Xn + (size_t)j * k, x, k);
}
// y (+)= Xnnz * w
template <class T>
inline void gemv_n(const T* Xn, const T* w, T* y, int k, int n, bool accumulate) {
  if (!accumulate)
    for (int i = 0; i < k; i++) y[i] = T(0);
  for (int j = 0; j < n; j++) {
    const T s = w[j];
    const T* a = Xn + (size_t)j * k;
#pragma omp simd
    for (int i = 0; i < k; i++) y[i] += a[i] * s;
  }
}

template <class T>
struct Scratch {
  std::vector<T> Xn, conf, conf1, u, w, x, r, p, Ap, lhs, rhs, mu;
  void size(int k, int n) {
    if ((size_t)k * n > Xn.size()) Xn.resize((size_t)k * n);
    if ((size_t)n > conf.size()) { conf.resize(n); conf1.resize(n); u.resize(n); w.resize(n); }
    if ((size_t)k > x.size()) {
      x.resize(k); r.resize(k); p.resize(k); Ap.resize(k); rhs.resize(k); mu.resize(k);
      lhs.resize((size_t)k * k);
    }
  }
};

// wrmf_implicit.hpp:8-32
template <class T>
void cg_solver_implicit(const T* Xn, const T* conf, const T* conf1, T* x /*in: x_old, out*/,
                        int n_iter, const T* XtX, int k, int n, Scratch<T>& s) {
  T* r = s.r.data(); T* p = s.p.data(); T* Ap = s.Ap.data(); T* u = s.u.data(); T* w = s.w.data();
  // r = X_nnz * (confidence - (confidence_1 % (X_nnz.t() * x))) - XtX * x        (:16)
  gemv_t(Xn, x, u, k, n);
  for (int j = 0; j < n; j++) w[j] = conf[j] - conf1[j] * u[j];
  gemv_n(Xn, w, r, k, n, false);
  symv(XtX, x, Ap, k);
  for (int i = 0; i < k; i++) { r[i] -= Ap[i]; p[i] = r[i]; }
  double rsold, rsnew, alpha;
  rsold = dotT(r, r, k);
  for (int it = 0; it < n_iter; it++) {
    // Ap = XtX * p + X_nnz * (confidence_1 % (X_nnz.t() * p))                    (:22)
    symv(XtX, p, Ap, k);
    gemv_t(Xn, p, u, k, n);
    for (int j = 0; j < n; j++) w[j] = conf1[j] * u[j];
    gemv_n(Xn, w, Ap, k, n, true);
    alpha = rsold / dotT(p, Ap, k);                                              // (:23)
    const T a = (T)alpha;
    for (int i = 0; i < k; i++) { x[i] += a * p[i]; r[i] -= a * Ap[i]; }         // (:24-25)
    rsnew = dotT(r, r, k);
    if (rsnew < CG_TOL) break;                                                   // (:27)
    const T b = (T)(rsnew / rsold);
    for (int i = 0; i < k; i++) p[i] = r[i] + p[i] * b;                          // (:28)
    rsold = rsnew;
  }
}

// wrmf_explicit.hpp:8-31
template <class T>
void cg_solver_explicit(const T* Xn, const T* conf, T* x, T lambda, int n_iter, int k, int n,
                        Scratch<T>& s) {
  T* r = s.r.data(); T* p = s.p.data(); T* Ap = s.Ap.data(); T* u = s.u.data(); T* w = s.w.data();
  gemv_t(Xn, x, u, k, n);
  for (int j = 0; j < n; j++) w[j] = conf[j] - u[j];
  gemv_n(Xn, w, r, k, n, false);
  for (int i = 0; i < k; i++) { r[i] -= lambda * x[i]; p[i] = r[i]; }            // (:15)
  double rsold, rsnew, alpha;
  rsold = dotT(r, r, k);
  for (int it = 0; it < n_iter; it++) {
    gemv_t(Xn, p, u, k, n);
    gemv_n(Xn, u, Ap, k, n, false);
    for (int i = 0; i < k; i++) Ap[i] += lambda * p[i];                          // (:21)
    alpha = rsold / dotT(p, Ap, k);
    const T a = (T)alpha;
    for (int i = 0; i < k; i++) { x[i] += a * p[i]; r[i] -= a * Ap[i]; }
    rsnew = dotT(r, r, k);
    if (rsnew < CG_TOL) break;
    const T b = (T)(rsnew / rsold);
    for (int i = 0; i < k; i++) p[i] = r[i] + p[i] * b;
    rsold = rsnew;
  }
}

// solve(lhs, rhs, fast [+ likely_sympd]): LAPACK posv semantics, LU (gesv) fallback when a
// pivot is not positive -- what arma::solve does for a symmetric positive-definite system
// (wrmf_implicit.hpp:236, wrmf_explicit.hpp:108).  A (column-major, full) is destroyed.
template <class T>
bool chol_solve(T* A, T* b, int n) {
  for (int j = 0; j < n; j++) {
    // right-looking: column j already holds A[:,j] minus the contributions of columns < j
    T d = A[(size_t)j * n + j];
    if (!(d > T(0))) return false;
    d = std::sqrt(d);
    T* col_j = A + (size_t)j * n;
    col_j[j] = d;
    const T inv = T(1) / d;
    for (int i = j + 1; i < n; i++) col_j[i] *= inv;
    for (int c = j + 1; c < n; c++) {
      const T s = col_j[c];
      T* col_c = A + (size_t)c * n;
#pragma omp simd
      for (int i = c; i < n; i++) col_c[i] -= col_j[i] * s;
    }
  }
  for (int i = 0; i < n; i++) {  // L z = b
    const T* col = A + (size_t)i * n;
    b[i] /= col[i];
    const T bi = b[i];
    for (int l = i + 1; l < n; l++) b[l] -= col[l] * bi;
  }
  for (int i = n - 1; i >= 0; i--) {  // L^T y = z
    const T* col = A + (size_t)i * n;
    T s = b[i];
    for (int l = i + 1; l < n; l++) s -= col[l] * b[l];
    b[i] = s / col[i];
  }
  return true;
}
template <class T>
bool lu_solve(T* A, T* b, int n) {
  for (int j = 0; j < n; j++) {
    int piv = j; T best = std::abs(A[(size_t)j * n + j]);
    for (int i = j + 1; i < n; i++)
      if (std::abs(A[(size_t)j * n + i]) > best) { best = std::abs(A[(size_t)j * n + i]); piv = i; }
    if (best == T(0)) return false;
    if (piv != j) {
      for (int c = 0; c < n; c++) std::swap(A[(size_t)c * n + j], A[(size_t)c * n + piv]);
      std::swap(b[j], b[piv]);
    }
    for (int i = j + 1; i < n; i++) {
      const T f = A[(size_t)j * n + i] / A[(size_t)j * n + j];
      for (int c = j + 1; c < n; c++) A[(size_t)c * n + i] -= f * A[(size_t)c * n + j];
      b[i] -= f * b[j];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    T s = b[i];
    for (int l = i + 1; l < n; l++) s -= A[(size_t)l * n + i] * b[l];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}
template <class T>
void sympd_solve(T* lhs, T* rhs, int k, std::vector<T>& spare) {
  spare.assign(lhs, lhs + (size_t)k * k);
  std::vector<T> b0(rhs, rhs + k);
  if (chol_solve(lhs, rhs, k)) return;
  std::copy(b0.begin(), b0.end(), rhs);
  lu_solve(spare.data(), rhs, k);
}

// lhs += sum_j w_j x_j x_j^T  over the gathered columns (full symmetric matrix is written)
template <class T>
void weighted_gram(const T* Xn, const T* w, T* lhs, int k, int n) {
  for (int j = 0; j < n; j++) {
    const T* xj = Xn + (size_t)j * k;
    const T wj = w[j];
    for (int c = 0; c < k; c++) {
      const T s = xj[c] * wj;
      T* col = lhs + (size_t)c * k;
#pragma omp simd
      for (int i = 0; i < k; i++) col[i] += xj[i] * s;
    }
  }
}

// nnls.hpp:10-48 (c_nnls + scd_ls_update), X = lhs (k x k), y = rhs
template <class T>
void c_nnls(const T* lhs, const T* rhs, T* x /*in: init, out*/, int k, Scratch<T>& s) {
  std::vector<T> XtX((size_t)k * k, T(0));
  for (int c = 0; c < k; c++)
    for (int r_ = 0; r_ < k; r_++)
      XtX[(size_t)c * k + r_] = dotT(lhs + (size_t)r_ * k, lhs + (size_t)c * k, k);
  for (int i = 0; i < k; i++) XtX[(size_t)i * k + i] += (T)NNLS_EPS;
  T* mu = s.mu.data();
  symv(XtX.data(), x, mu, k);
  for (int i = 0; i < k; i++) mu[i] -= dotT(lhs + (size_t)i * k, rhs, k);
  for (int t = 0; t < SCD_MAX_ITER; t++) {
    T rel_diff = 0;
    for (int c = 0; c < k; c++) {
      const T old_value = x[c];
      T new_value = old_value - mu[c] / XtX[(size_t)c * k + c];
      if (new_value < 0) new_value = 0;
      const T diff = new_value - old_value;
      if (diff != 0) {
        x[c] = new_value;
        const T* col = XtX.data() + (size_t)c * k;
        for (int i = 0; i < k; i++) mu[i] += diff * col[i];
        const auto step_err = std::abs(diff) / (std::abs(old_value) + NNLS_EPS);
        if (step_err > rel_diff) rel_diff = (T)step_err;
      }
    }
    if (rel_diff <= SCD_TOL) break;
  }
}

template <class T>
double als_implicit(int nc, size_t nnz_total, const int* col_ptrs, const int* row_indices,
                    const double* values, const T* X, int k, int n_src, T* Y, const T* XtX,
                    double lambda, int n_threads, int solver, int cg_steps) {
  double loss = 0;
#pragma omp parallel num_threads(n_threads)
  {
    Scratch<T> s;
    std::vector<T> spare;
#pragma omp for schedule(dynamic) reduction(+ : loss)
    for (int i = 0; i < nc; i++) {
      const int p1 = col_ptrs[i], p2 = col_ptrs[i + 1];
      T* y = Y + (size_t)i * k;
      if (p1 < p2) {                                                             // (:179)
        const int n = p2 - p1;
        s.size(k, n);
        T* Xn = s.Xn.data(); T* conf = s.conf.data(); T* conf1 = s.conf1.data();
        for (int j = 0; j < n; j++) {
          conf[j] = (T)values[p1 + j];                                           // (:182-183)
          conf1[j] = conf[j] - T(1.0);
          std::memcpy(Xn + (size_t)j * k, X + (size_t)row_indices[p1 + j] * k, sizeof(T) * k);  // (:184)
        }
        T* ynew = s.x.data();
        std::memcpy(ynew, y, sizeof(T) * k);                                     // init = Y.col(i) (:185)
        if (solver == CONJUGATE_GRADIENT) {
          cg_solver_implicit<T>(Xn, conf, conf1, ynew, cg_steps, XtX, k, n, s);  // (:197)
        } else {
          T* lhs = s.lhs.data(); T* rhs = s.rhs.data();
          std::memcpy(lhs, XtX, sizeof(T) * (size_t)k * k);
          weighted_gram(Xn, conf1, lhs, k, n);                                   // (:207-208)
          gemv_n(Xn, conf, rhs, k, n, false);                                    // (:231)
          if (solver == SEQ_COORDINATE_WISE_NNLS) {
            c_nnls<T>(lhs, rhs, ynew, k, s);                                     // (:233-234)
          } else {
            sympd_solve(lhs, rhs, k, spare);                                     // (:236)
            std::memcpy(ynew, rhs, sizeof(T) * k);
          }
        }
        std::memcpy(y, ynew, sizeof(T) * k);                                     // (:254)
        // loss += dot(square(1 - Y_new' X_nnz), confidence) + lambda * dot(Y_new, Y_new)  (:259-261)
        T* u = s.u.data();
        gemv_t(Xn, ynew, u, k, n);
        T acc = T(0);
        for (int j = 0; j < n; j++) { const T d = T(1) - u[j]; acc += d * d * conf[j]; }
        loss += acc + lambda * dotT(ynew, ynew, k);
      } else {
        for (int c = 0; c < k; c++) y[c] = T(0);                                 // (:281)
      }
    }
  }
  if (lambda > 0) {                                                              // (:286-302)
    T acc = T(0);
    const size_t tot = (size_t)k * n_src;
    for (size_t t = 0; t < tot; t++) acc += X[t] * X[t];
    loss += lambda * acc;
  }
  return (double)(T)(loss / nnz_total);                                          // (:304) returns T
}

template <class T>
double als_explicit(int nc, size_t nnz_total, const int* col_ptrs, const int* row_indices,
                    const double* values, const T* X, int k, int n_src, T* Y, const T* cnt_X,
                    double lambda, int n_threads, int solver, int cg_steps, bool dynamic_lambda) {
  double loss = 0;
#pragma omp parallel num_threads(n_threads)
  {
    Scratch<T> s;
    std::vector<T> spare;
#pragma omp for schedule(dynamic, GRAIN_SIZE) reduction(+ : loss)
    for (int i = 0; i < nc; i++) {
      const int p1 = col_ptrs[i], p2 = col_ptrs[i + 1];
      T* y = Y + (size_t)i * k;
      if (p1 < p2) {
        const int n = p2 - p1;
        s.size(k, n);
        const T lambda_use = (T)(lambda * (dynamic_lambda ? static_cast<T>(p2 - p1) : 1.));  // (:78)
        T* Xn = s.Xn.data(); T* conf = s.conf.data();
        for (int j = 0; j < n; j++) {
          conf[j] = (T)values[p1 + j];
          std::memcpy(Xn + (size_t)j * k, X + (size_t)row_indices[p1 + j] * k, sizeof(T) * k);
        }
        T* ynew = s.x.data();
        std::memcpy(ynew, y, sizeof(T) * k);
        if (solver == CONJUGATE_GRADIENT) {
          cg_solver_explicit<T>(Xn, conf, ynew, lambda_use, cg_steps, k, n, s);  // (:101)
        } else {
          T* lhs = s.lhs.data(); T* rhs = s.rhs.data();
          std::fill(lhs, lhs + (size_t)k * k, T(0));
          std::vector<T>& ones = s.w;
          for (int j = 0; j < n; j++) ones[j] = T(1);
          weighted_gram(Xn, ones.data(), lhs, k, n);                             // (:103)
          for (int c = 0; c < k; c++) lhs[(size_t)c * k + c] += lambda_use;      // (:104)
          gemv_n(Xn, conf, rhs, k, n, false);                                    // (:105)
          if (solver == CHOLESKY) {
            sympd_solve(lhs, rhs, k, spare);                                     // (:108)
            std::memcpy(ynew, rhs, sizeof(T) * k);
          } else {
            c_nnls<T>(lhs, rhs, ynew, k, s);                                     // (:110)
          }
        }
        std::memcpy(y, ynew, sizeof(T) * k);
        T* u = s.u.data();
        gemv_t(Xn, ynew, u, k, n);
        T acc = T(0);
        for (int j = 0; j < n; j++) { const T d = conf[j] - u[j]; acc += d * d; }  // (:131)
        loss += acc + lambda_use * dotT(ynew, ynew, k);                          // (:132)
      } else {
        for (int c = 0; c < k; c++) y[c] = T(0);
      }
    }
  }
  if (lambda > 0) {                                                              // (:147-172)
    if (!dynamic_lambda) {
      T acc = T(0);
      const size_t tot = (size_t)k * n_src;
      for (size_t t = 0; t < tot; t++) acc += X[t] * X[t];
      loss += lambda * acc;
    } else {
      // accu((X % X) * cnt_X): squares weighted per source column by cnt_X
      std::vector<T> rows(k, T(0));
      for (int j = 0; j < n_src; j++) {
        const T c = cnt_X[j];
        const T* x = X + (size_t)j * k;
        for (int r_ = 0; r_ < k; r_++) rows[r_] += x[r_] * x[r_] * c;
      }
      T acc = T(0);
      for (int r_ = 0; r_ < k; r_++) acc += rows[r_];
      loss += lambda * acc;
    }
  }
  return (double)(T)(loss / nnz_total);
}

// ---------------------------------------------------------------------------------------------------
// Bias variants (with_user_item_bias / with_global_bias).  Kept apart from the hot no-bias functions
// above (those are the timed CPU baseline).  Layout, wrmf_implicit.hpp:96-101 / wrmf_explicit.hpp:38-52:
//   is_x_bias_last_row:  X = [1, ..., x_bias]   Y = [y_bias, ..., 1]
//   otherwise:           X = [x_bias, ..., 1]   Y = [1, ..., y_bias]
// `kf` = rows of X / Y (rank + 2 with biases), k = kf - with_biases = size of the solved system.
struct BiasLayout {
  int k, xo, xb, io, oo;
  BiasLayout(int kf, bool with_biases, bool is_last) {
    k = kf - (with_biases ? 1 : 0);
    xo = (with_biases && !is_last) ? 1 : 0;   // drop_row(X_nnz, is_x_bias_last_row)      (:190 / :88)
    xb = is_last ? kf - 1 : 0;                // x_biases = last / first row of X           (:116-119)
    io = (with_biases && is_last) ? 1 : 0;    // init = drop_row(init, !is_x_bias_last_row) (:191 / :90) -- sic
    oo = (with_biases && !is_last) ? 1 : 0;   // Y.head(rank-1) / Y.tail(rank-1)            (:246-252)
  }
};

// cg_solver_implicit_global_bias (:35-60) and cg_solver_implicit_user_item_bias (:62-88) in one body:
//   r = X_nnz (c - (c-1) % (X_nnz' x + x_biases + global_bias)) - XtX x + rhs_init
template <class T>
void cg_solver_implicit_bias(const T* Xn, const T* conf, const T* conf1, const T* xbn, T gbias, const T* rhs_init,
                             T* x, int n_iter, const T* XtX, int k, int n, Scratch<T>& s) {
  T* r = s.r.data(); T* p = s.p.data(); T* Ap = s.Ap.data(); T* u = s.u.data(); T* w = s.w.data();
  gemv_t(Xn, x, u, k, n);
  for (int j = 0; j < n; j++) w[j] = conf[j] - conf1[j] * ((xbn ? u[j] + xbn[j] : u[j]) + gbias);
  gemv_n(Xn, w, r, k, n, false);
  symv(XtX, x, Ap, k);
  for (int i = 0; i < k; i++) { r[i] = r[i] - Ap[i] + rhs_init[i]; p[i] = r[i]; }
  double rsold, rsnew, alpha;
  rsold = dotT(r, r, k);
  for (int it = 0; it < n_iter; it++) {
    symv(XtX, p, Ap, k);
    gemv_t(Xn, p, u, k, n);
    for (int j = 0; j < n; j++) w[j] = conf1[j] * u[j];
    gemv_n(Xn, w, Ap, k, n, true);
    alpha = rsold / dotT(p, Ap, k);
    const T a = (T)alpha;
    for (int i = 0; i < k; i++) { x[i] += a * p[i]; r[i] -= a * Ap[i]; }
    rsnew = dotT(r, r, k);
    if (rsnew < CG_TOL) break;
    const T b = (T)(rsnew / rsold);
    for (int i = 0; i < k; i++) p[i] = r[i] + p[i] * b;
    rsold = rsnew;
  }
}

// als_implicit<T> with with_biases and/or global_bias (wrmf_implicit.hpp:90-305, all branches).
// Deviation, on purpose: the reference's CONJUGATE_GRADIENT + with_biases branch drops a row of `init` twice
// (:191 and :199) and fails with a dimension error; here `init` is dropped once, like for the other solvers.
template <class T>
double als_implicit_bias(int nc, size_t nnz_total, const int* col_ptrs, const int* row_indices, const double* values,
                         const T* X, int kf, int n_src, T* Y, const T* XtX, double lambda, int n_threads, int solver,
                         int cg_steps, bool with_biases, bool is_last, double global_bias, T* global_bias_base,
                         bool initialize_bias_base) {
  const BiasLayout L(kf, with_biases, is_last);
  const int k = L.k;
  if (global_bias < std::sqrt(std::numeric_limits<T>::epsilon())) global_bias = 0;          // (:108-109)
  std::vector<T> rhs_init(k, T(0));
  if (global_bias && initialize_bias_base && !with_biases) {                                // (:111-112)
    for (int f = 0; f < k; f++) {
      T acc = T(0);
      for (int j = 0; j < n_src; j++) acc += X[(size_t)j * kf + f];
      global_bias_base[f] = acc * (T)(-global_bias);
    }
  }
  if (with_biases) {                                                                        // (:114-154)
    for (int j = 0; j < n_src; j++) {
      const T* xj = X + (size_t)j * kf;
      const T w = global_bias ? (T)(xj[L.xb] + (T)global_bias) : xj[L.xb];
      for (int f = 0; f < k; f++) rhs_init[f] += xj[L.xo + f] * w;
    }
    for (int f = 0; f < k; f++) rhs_init[f] = -rhs_init[f];
  } else if (global_bias) {
    for (int f = 0; f < k; f++) rhs_init[f] = global_bias_base[f];                          // (:155-157)
  }
  const T gb = (T)global_bias;
  double loss = 0;
#pragma omp parallel num_threads(n_threads)
  {
    Scratch<T> s;
    std::vector<T> spare, xbn;
#pragma omp for schedule(dynamic) reduction(+ : loss)
    for (int i = 0; i < nc; i++) {
      const int p1 = col_ptrs[i], p2 = col_ptrs[i + 1];
      T* y = Y + (size_t)i * kf;
      if (with_biases || global_bias || p1 < p2) {                                          // (:179)
        const int n = p2 - p1;
        s.size(k, n > 0 ? n : 1);
        xbn.resize(n > 0 ? n : 1);
        T* Xn = s.Xn.data(); T* conf = s.conf.data(); T* conf1 = s.conf1.data();
        for (int j = 0; j < n; j++) {
          const T* xj = X + (size_t)row_indices[p1 + j] * kf;
          conf[j] = (T)values[p1 + j];
          conf1[j] = conf[j] - T(1.0);
          xbn[j] = with_biases ? xj[L.xb] : T(0);
          std::memcpy(Xn + (size_t)j * k, xj + L.xo, sizeof(T) * k);
        }
        T* ynew = s.x.data();
        std::memcpy(ynew, y + L.io, sizeof(T) * k);
        if (solver == CONJUGATE_GRADIENT) {
          cg_solver_implicit_bias<T>(Xn, conf, conf1, with_biases ? xbn.data() : nullptr, gb, rhs_init.data(), ynew,
                                     cg_steps, XtX, k, n, s);                               // (:200-205)
        } else {
          T* lhs = s.lhs.data(); T* rhs = s.rhs.data();
          std::memcpy(lhs, XtX, sizeof(T) * (size_t)k * k);
          weighted_gram(Xn, conf1, lhs, k, n);                                              // (:207-208)
          T* w = s.w.data();
          if (with_biases) {
            for (int j = 0; j < n; j++) w[j] = conf[j] - xbn[j] * conf1[j];                  // (:226)
          } else {
            for (int j = 0; j < n; j++) w[j] = conf[j];                                     // (:229)
          }
          gemv_n(Xn, w, rhs, k, n, false);
          for (int f = 0; f < k; f++) rhs[f] += rhs_init[f];
          if (solver == SEQ_COORDINATE_WISE_NNLS) {
            c_nnls<T>(lhs, rhs, ynew, k, s);
          } else {
            sympd_solve(lhs, rhs, k, spare);
            std::memcpy(ynew, rhs, sizeof(T) * k);
          }
        }
        std::memcpy(y + L.oo, ynew, sizeof(T) * k);                                         // (:240-255)
        T acc = T(0);                                                                       // (:257-270)
        if (p1 < p2) {
          T* u = s.u.data();
          gemv_t(Xn, ynew, u, k, n);
          const T one_g = (T)(1 - global_bias);
          for (int j = 0; j < n; j++) { const T d = one_g - u[j] - xbn[j]; acc += d * d * conf[j]; }
        }
        loss += acc + lambda * dotT(ynew, ynew, k);
      } else {
        for (int c = 0; c < k; c++) y[L.oo + c] = T(0);                                     // (:272-282)
      }
    }
  }
  if (lambda > 0) {                                                                         // (:286-302)
    // with biases: every learned row of X, i.e. all but the row of ones (first row when is_x_bias_last_row)
    const int lo = with_biases ? (is_last ? 1 : 0) : 0, hi = with_biases ? (is_last ? kf : kf - 1) : kf;
    T acc = T(0);
    for (int j = 0; j < n_src; j++)
      for (int f = lo; f < hi; f++) acc += X[(size_t)j * kf + f] * X[(size_t)j * kf + f];
    loss += lambda * acc;
  }
  return (double)(T)(loss / nnz_total);
}

// als_explicit<T> with with_biases (wrmf_explicit.hpp:33-174, all branches).
template <class T>
double als_explicit_bias(int nc, size_t nnz_total, const int* col_ptrs, const int* row_indices, const double* values,
                         const T* X, int kf, int n_src, T* Y, const T* cnt_X, double lambda, int n_threads, int solver,
                         int cg_steps, bool dynamic_lambda, bool with_biases, bool is_last) {
  const BiasLayout L(kf, with_biases, is_last);
  const int k = L.k;
  double loss = 0;
#pragma omp parallel num_threads(n_threads)
  {
    Scratch<T> s;
    std::vector<T> spare;
#pragma omp for schedule(dynamic, GRAIN_SIZE) reduction(+ : loss)
    for (int i = 0; i < nc; i++) {
      const int p1 = col_ptrs[i], p2 = col_ptrs[i + 1];
      T* y = Y + (size_t)i * kf;
      if (p1 < p2) {
        const int n = p2 - p1;
        s.size(k, n);
        const T lambda_use = (T)(lambda * (dynamic_lambda ? static_cast<T>(p2 - p1) : 1.));
        T* Xn = s.Xn.data(); T* conf = s.conf.data();
        for (int j = 0; j < n; j++) {
          const T* xj = X + (size_t)row_indices[p1 + j] * kf;
          conf[j] = (T)values[p1 + j];
          if (with_biases) conf[j] -= xj[L.xb];                                             // (:89)
          std::memcpy(Xn + (size_t)j * k, xj + L.xo, sizeof(T) * k);
        }
        T* ynew = s.x.data();
        std::memcpy(ynew, y + L.io, sizeof(T) * k);                                         // (:90)
        if (solver == CONJUGATE_GRADIENT) {
          cg_solver_explicit<T>(Xn, conf, ynew, lambda_use, cg_steps, k, n, s);
        } else {
          T* lhs = s.lhs.data(); T* rhs = s.rhs.data();
          std::fill(lhs, lhs + (size_t)k * k, T(0));
          std::vector<T>& ones = s.w;
          for (int j = 0; j < n; j++) ones[j] = T(1);
          weighted_gram(Xn, ones.data(), lhs, k, n);
          for (int c = 0; c < k; c++) lhs[(size_t)c * k + c] += lambda_use;
          gemv_n(Xn, conf, rhs, k, n, false);
          if (solver == CHOLESKY) {
            sympd_solve(lhs, rhs, k, spare);
            std::memcpy(ynew, rhs, sizeof(T) * k);
          } else {
            c_nnls<T>(lhs, rhs, ynew, k, s);
          }
        }
        std::memcpy(y + L.oo, ynew, sizeof(T) * k);                                         // (:115-129)
        T* u = s.u.data();
        gemv_t(Xn, ynew, u, k, n);
        T acc = T(0);
        for (int j = 0; j < n; j++) { const T d = conf[j] - u[j]; acc += d * d; }
        loss += acc + lambda_use * dotT(ynew, ynew, k);
      } else {
        for (int c = 0; c < k; c++) y[L.oo + c] = T(0);                                     // (:134-145)
      }
    }
  }
  if (lambda > 0) {                                                                         // (:148-172)
    const int lo = with_biases ? (is_last ? 1 : 0) : 0, hi = with_biases ? (is_last ? kf : kf - 1) : kf;
    T acc = T(0);
    for (int j = 0; j < n_src; j++) {
      const T c = dynamic_lambda ? cnt_X[j] : T(1);
      for (int f = lo; f < hi; f++) acc += X[(size_t)j * kf + f] * X[(size_t)j * kf + f] * c;
    }
    loss += lambda * acc;
  }
  return (double)(T)(loss / nnz_total);
}

// initialize_biases_explicit (wrmf_utils.hpp:32-82): csc = users x items by item column, csr = the same entries
// by user; `csc_val` / `csr_val` are modified in place when calculate_global_bias.
template <class T>
double initialize_biases_explicit(int n_items, int n_users, size_t nnz, const int* cp, const int* ci, double* cv,
                                  const int* rp, const int* ri, double* rv, T* user_bias, T* item_bias, T lambda,
                                  bool dynamic_lambda, bool non_negative, bool calculate_global_bias) {
  double global_bias = 0;
  if (calculate_global_bias) {
    for (size_t ix = 0; ix < nnz; ix++) global_bias += (cv[ix] - global_bias) / (double)(ix + 1);
    for (size_t ix = 0; ix < nnz; ix++) { cv[ix] -= global_bias; rv[ix] -= global_bias; }
  }
  for (int iter = 0; iter < 5; iter++) {
    for (int col = 0; col < n_items; col++) {
      item_bias[col] = 0;
      const T lambda_use = lambda * (dynamic_lambda ? static_cast<T>(cp[col + 1] - cp[col]) : 1.);
      for (int ix = cp[col]; ix < cp[col + 1]; ix++) item_bias[col] += cv[ix] - user_bias[ci[ix]];
      item_bias[col] /= lambda_use + static_cast<T>(cp[col + 1] - cp[col]);
      if (non_negative) item_bias[col] = std::fmax((T)0, item_bias[col]);
    }
    for (int row = 0; row < n_users; row++) {
      user_bias[row] = 0;
      const T lambda_use = lambda * (dynamic_lambda ? static_cast<T>(rp[row + 1] - rp[row]) : 1.);
      for (int ix = rp[row]; ix < rp[row + 1]; ix++) user_bias[row] += rv[ix] - item_bias[ri[ix]];
      user_bias[row] /= lambda_use + static_cast<T>(rp[row + 1] - rp[row]);
      if (non_negative) user_bias[row] = std::fmax((T)0, user_bias[row]);
    }
  }
  return global_bias;
}

// initialize_biases_implicit (wrmf_utils.hpp:84-167)
template <class T>
double initialize_biases_implicit(int n_items, int n_users, size_t nnz, const int* cp, const int* ci, const double* cv,
                                  const int* rp, const int* ri, const double* rv, T* user_bias, T* item_bias, T lambda,
                                  bool calculate_global_bias, bool non_negative) {
  double global_bias = 0;
  if (calculate_global_bias) {                                                              // Kahan sum (:12-30)
    long double err = 0., res = 0.;
    for (size_t ix = 0; ix < nnz; ix++) {
      const long double diff = rv[ix] - err, temp = res + diff;
      err = (temp - res) - diff;
      res = temp;
    }
    global_bias = res / (res + (long double)n_items * (long double)n_users - (long double)nnz);
  }
  if (non_negative) global_bias = std::fmax(0., global_bias);
  std::vector<double> um(n_users), im(n_items), ua(n_users), ia(n_items);
  auto means = [&](int n_rows, int n_other, const int* p, const double* v, std::vector<double>& m, std::vector<double>& a) {
    for (int row = 0; row < n_rows; row++) {
      const int cnt = p[row + 1] - p[row];
      if (cnt > 0) {
        for (int ix = p[row]; ix < p[row + 1]; ix++) a[row] += v[ix];
        m[row] = a[row] / (a[row] + (double)(n_other - cnt));
        a[row] += (double)(n_other - cnt);
        a[row] /= a[row] + lambda;
      } else {
        m[row] = 0;
        a[row] = (double)n_other / ((double)n_other + lambda);
      }
    }
  };
  means(n_users, n_items, rp, rv, um, ua);
  means(n_items, n_users, cp, cv, im, ia);
  double bias_mean, bias_this, wsum;
  for (int iter = 0; iter < 5; iter++) {
    bias_mean = 0;
    if (iter > 0)
      for (int row = 0; row < n_users; row++) bias_mean += (user_bias[row] - bias_mean) / (T)(row + 1);
    for (int col = 0; col < n_items; col++) {
      wsum = n_users;
      bias_this = bias_mean;
      for (int ix = cp[col]; ix < cp[col + 1]; ix++)
        bias_this += ((cv[ix] - 1) * (user_bias[ci[ix]] - bias_this)) / (wsum += (cv[ix] - 1));
      item_bias[col] = (im[col] - bias_this - global_bias) * ia[col];
    }
    if (non_negative)
      for (int col = 0; col < n_items; col++) item_bias[col] = std::fmax((T)0, item_bias[col]);
    bias_mean = 0;
    for (int col = 0; col < n_items; col++) bias_mean += (item_bias[col] - bias_mean) / (T)(col + 1);
    for (int row = 0; row < n_users; row++) {
      wsum = n_items;
      bias_this = bias_mean;
      for (int ix = rp[row]; ix < rp[row + 1]; ix++)
        bias_this += ((rv[ix] - 1) * (item_bias[ri[ix]] - bias_this)) / (wsum += (rv[ix] - 1));
      user_bias[row] = (um[row] - bias_this - global_bias) * ua[row];
    }
    if (non_negative)
      for (int row = 0; row < n_users; row++) user_bias[row] = std::fmax((T)0, user_bias[row]);
  }
  return global_bias;
}

// XtX = tcrossprod(X) + diag(lambda)   (R/model_WRMF.R:474-486); BLAS syrk/gemm in R.
template <class T>
void gram(const T* X, int k, int n_src, double lambda, T* XtX, int n_threads) {
  std::vector<double> acc((size_t)k * k, 0.0);
#pragma omp parallel num_threads(n_threads)
  {
    std::vector<T> loc((size_t)k * k, T(0));
    const int CH = 256;  // blocked accumulation in T, combined in double across blocks
#pragma omp for schedule(static)
    for (int b = 0; b < (n_src + CH - 1) / CH; b++) {
      std::fill(loc.begin(), loc.end(), T(0));
      const int j1 = std::min(n_src, (b + 1) * CH);
      for (int j = b * CH; j < j1; j++) {
        const T* x = X + (size_t)j * k;
        for (int c = 0; c < k; c++) {
          const T s = x[c];
          T* col = loc.data() + (size_t)c * k;
#pragma omp simd
          for (int i = 0; i < k; i++) col[i] += x[i] * s;
        }
      }
#pragma omp critical
      for (size_t t = 0; t < acc.size(); t++) acc[t] += (double)loc[t];
    }
  }
  for (size_t t = 0; t < acc.size(); t++) XtX[t] = (T)acc[t];
  for (int i = 0; i < k; i++) XtX[(size_t)i * k + i] = (T)(acc[(size_t)i * k + i] + lambda);
}

}  // namespace

extern "C" {

int oracle_max_threads(void) {
#ifdef _OPENMP
  // src/utils.cpp:84-91 (omp_thread_count)
  int a = omp_get_max_threads(), b = omp_get_thread_limit();
  return a < b ? a : b;
#else
  return 1;
#endif
}


// Synthetic CSR of bench.py / BASELINE.md section 2 (row r draws one ascending column id per equal-width stratum of
// [0, n_cols) from a counter-based splitmix64 hash; values 1 + floor(10 u^2) or ratings 1..5).  Restated here so that
// bench.py's CPU arms (`--impl reference`, `cpu_baseline`) generate their inputs WITHOUT loading the product library;
// tests/test_abi.py checks that it equals the product's generator entry for entry.
static inline uint64_t synth_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
int oracle_synth_csr(int n_rows, int n_cols, int nnz_per_row, unsigned long long seed, int explicit_values,
                     long long row_offset, int* ptr, int* idx, double* val, int n_threads) {
  if (n_rows < 0 || n_cols <= 0 || nnz_per_row <= 0 || nnz_per_row > n_cols) return 1;
  if ((long long)n_rows * nnz_per_row > 2147483647LL) return 2;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long r = 0; r < (long long)n_rows; r++) {
    ptr[r] = (int)(r * nnz_per_row);
    for (int j = 0; j < nnz_per_row; j++) {
      const uint64_t h = synth_mix((uint64_t)seed * 0x100000001B3ull + (uint64_t)(r + row_offset) * (uint64_t)nnz_per_row + (uint64_t)j);
      const long long lo = ((long long)j * n_cols) / nnz_per_row, hi = ((long long)(j + 1) * n_cols) / nnz_per_row;
      const long long e = r * nnz_per_row + j;
      idx[e] = (int)(lo + (long long)(h % (uint64_t)(hi - lo)));
      const float u = (float)((h >> 40) & 0xFFFFFF) / 16777216.0f;
      if (val) val[e] = (double)(explicit_values ? (1.0f + floorf(u * 5.0f)) : (1.0f + floorf(10.0f * u * u)));
    }
  }
  ptr[n_rows] = (int)((long long)n_rows * nnz_per_row);
  return 0;
}

double oracle_als_implicit_f32(int nc, size_t nnz, const int* p, const int* idx, const double* v,
                               const float* X, int k, int n_src, float* Y, const float* XtX,
                               double lambda, int n_threads, int solver, int cg_steps) {
  return als_implicit<float>(nc, nnz, p, idx, v, X, k, n_src, Y, XtX, lambda, n_threads, solver, cg_steps);
}
double oracle_als_implicit_f64(int nc, size_t nnz, const int* p, const int* idx, const double* v,
                               const double* X, int k, int n_src, double* Y, const double* XtX,
                               double lambda, int n_threads, int solver, int cg_steps) {
  return als_implicit<double>(nc, nnz, p, idx, v, X, k, n_src, Y, XtX, lambda, n_threads, solver, cg_steps);
}
double oracle_als_explicit_f32(int nc, size_t nnz, const int* p, const int* idx, const double* v,
                               const float* X, int k, int n_src, float* Y, const float* cnt_X,
                               double lambda, int n_threads, int solver, int cg_steps, int dynamic_lambda) {
  return als_explicit<float>(nc, nnz, p, idx, v, X, k, n_src, Y, cnt_X, lambda, n_threads, solver,
                             cg_steps, dynamic_lambda != 0);
}
double oracle_als_explicit_f64(int nc, size_t nnz, const int* p, const int* idx, const double* v,
                               const double* X, int k, int n_src, double* Y, const double* cnt_X,
                               double lambda, int n_threads, int solver, int cg_steps, int dynamic_lambda) {
  return als_explicit<double>(nc, nnz, p, idx, v, X, k, n_src, Y, cnt_X, lambda, n_threads, solver,
                              cg_steps, dynamic_lambda != 0);
}
void oracle_gram_f32(const float* X, int k, int n_src, double lambda, float* XtX, int n_threads) {
  gram<float>(X, k, n_src, lambda, XtX, n_threads);
}
void oracle_gram_f64(const double* X, int k, int n_src, double lambda, double* XtX, int n_threads) {
  gram<double>(X, k, n_src, lambda, XtX, n_threads);
}

#define ORACLE_BIAS_EXPORTS(T, SFX)                                                                                    \
  double oracle_als_implicit_bias_##SFX(int nc, size_t nnz, const int* p, const int* idx, const double* v, const T* X,  \
                                        int kf, int n_src, T* Y, const T* XtX, double lambda, int n_threads,           \
                                        int solver, int cg_steps, int with_biases, int is_last, double global_bias,    \
                                        T* gbb, int init_base) {                                                       \
    return als_implicit_bias<T>(nc, nnz, p, idx, v, X, kf, n_src, Y, XtX, lambda, n_threads, solver, cg_steps,         \
                                with_biases != 0, is_last != 0, global_bias, gbb, init_base != 0);                     \
  }                                                                                                                    \
  double oracle_als_explicit_bias_##SFX(int nc, size_t nnz, const int* p, const int* idx, const double* v, const T* X,  \
                                        int kf, int n_src, T* Y, const T* cnt_X, double lambda, int n_threads,         \
                                        int solver, int cg_steps, int dynamic_lambda, int with_biases, int is_last) {  \
    return als_explicit_bias<T>(nc, nnz, p, idx, v, X, kf, n_src, Y, cnt_X, lambda, n_threads, solver, cg_steps,       \
                                dynamic_lambda != 0, with_biases != 0, is_last != 0);                                  \
  }                                                                                                                    \
  double oracle_initialize_biases_##SFX(int n_items, int n_users, size_t nnz, const int* cp, const int* ci, double* cv, \
                                        const int* rp, const int* ri, double* rv, T* user_bias, T* item_bias,          \
                                        double lambda, int dynamic_lambda, int non_negative, int calc_global,          \
                                        int is_explicit) {                                                             \
    if (is_explicit)                                                                                                   \
      return initialize_biases_explicit<T>(n_items, n_users, nnz, cp, ci, cv, rp, ri, rv, user_bias, item_bias,        \
                                           (T)lambda, dynamic_lambda != 0, non_negative != 0, calc_global != 0);       \
    return initialize_biases_implicit<T>(n_items, n_users, nnz, cp, ci, cv, rp, ri, rv, user_bias, item_bias,          \
                                         (T)lambda, calc_global != 0, non_negative != 0);                              \
  }
ORACLE_BIAS_EXPORTS(float, f32)
ORACLE_BIAS_EXPORTS(double, f64)

}  // extern "C"
