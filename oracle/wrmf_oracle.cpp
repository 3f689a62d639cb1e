// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rsparse_b200/, the
// C-ABI library) may include, link or call this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or
// the timed CPU baseline.
//
// A dependency-free C++17/OpenMP restatement of the reference's WRMF ALS half-iteration,
// following the reference's loop structure line by line (citations relative to
// /root/reference/):
//   als_implicit<T>        inst/include/wrmf_implicit.hpp:90-305   (no-bias branches only)
//   cg_solver_implicit<T>  inst/include/wrmf_implicit.hpp:8-32
//   als_explicit<T>        inst/include/wrmf_explicit.hpp:33-174   (no-bias branches only)
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

}  // extern "C"
