// TEST INFRASTRUCTURE ONLY. Thin extern "C" entry points around the reference's own,
// unmodified template headers, which are compiled *in place* from /root/reference
// (never copied into this repo) against oracle/mini_arma/armadillo:
//   als_implicit<T>  -> /root/reference/inst/include/wrmf_implicit.hpp:90-305
//   als_explicit<T>  -> /root/reference/inst/include/wrmf_explicit.hpp:33-174
//   initialize_biases<T> -> /root/reference/inst/include/wrmf_utils.hpp:170-183
// Mirrors what src/wrmf_implicit.cpp:5-31 and src/wrmf_explicit.cpp:5-27 do with Rcpp:
// wrap caller memory in MappedCSC / arma::Mat views (no copy) and call the template.
// Output goes to oracle/_ref/libref_wrmf.so (git-ignored, travels to the GPU box).
// The reference's nnls.hpp / wrmf_utils.hpp carry no include guards and its own build keeps
// the two paths in separate translation units (src/wrmf_implicit.cpp, src/wrmf_explicit.cpp);
// build_ref.sh therefore compiles this file twice (-DREF_IMPLICIT, -DREF_EXPLICIT).
#if defined(REF_IMPLICIT)
#include "wrmf_implicit.hpp"
#elif defined(REF_EXPLICIT)
#include "wrmf_explicit.hpp"
#else
#error "compile with -DREF_IMPLICIT or -DREF_EXPLICIT"
#endif

#ifdef REF_IMPLICIT
template <class T>
static double run_implicit(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                           T* X, int k, int n_src, T* Y, T* XtX, int k_xtx, double lambda, int n_threads,
                           unsigned solver, unsigned cg_steps, int with_biases,
                           int is_x_bias_last_row, double global_bias, T* gbb, int gbb_len,
                           int initialize_bias_base) {
  const dMappedCSC Conf((arma::uword)n_rows, (arma::uword)n_cols, nnz, (arma::uword*)ri,
                        (arma::uword*)cp, vals);
  arma::Mat<T> Xm(X, (arma::uword)k, (arma::uword)n_src, false, true);
  arma::Mat<T> Ym(Y, (arma::uword)k, (arma::uword)n_cols, false, true);
  const arma::Mat<T> G(XtX, (arma::uword)k_xtx, (arma::uword)k_xtx, false, true);
  arma::Col<T> base(gbb, (arma::uword)gbb_len, false, true);
  return (double)als_implicit<T>(Conf, Xm, Ym, G, lambda, n_threads, solver, cg_steps,
                                 with_biases != 0, is_x_bias_last_row != 0, global_bias, base,
                                 initialize_bias_base != 0);
}

// initialize_biases<T> (wrmf_utils.hpp:170-183) as src/wrmf_init.cpp:6-34 calls it: csc = user x item matrix by
// item column, csr = the same entries by user row (a CSC of the transpose); values may be modified in place.
template <class T>
static double run_init_biases(int n_users, int n_items, size_t nnz, int* csc_ri, int* csc_cp, double* csc_v,
                              int* csr_ri, int* csr_cp, double* csr_v, T* user_bias, T* item_bias, double lambda,
                              int dynamic_lambda, int non_negative, int calc_global, int is_explicit) {
  dMappedCSC csc((arma::uword)n_users, (arma::uword)n_items, nnz, (arma::uword*)csc_ri, (arma::uword*)csc_cp, csc_v);
  dMappedCSC csr((arma::uword)n_items, (arma::uword)n_users, nnz, (arma::uword*)csr_ri, (arma::uword*)csr_cp, csr_v);
  arma::Col<T> ub(user_bias, (arma::uword)n_users, false, true);
  arma::Col<T> ib(item_bias, (arma::uword)n_items, false, true);
  return initialize_biases<T>(csc, csr, ub, ib, (T)lambda, dynamic_lambda != 0, non_negative != 0, calc_global != 0,
                              is_explicit != 0);
}

#endif
#ifdef REF_EXPLICIT
template <class T>
static double run_explicit(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                           T* X, int k, int n_src, T* Y, T* cnt_X, int cnt_len, double lambda,
                           int n_threads, unsigned solver, unsigned cg_steps, int dynamic_lambda,
                           int with_biases, int is_x_bias_last_row) {
  const dMappedCSC Conf((arma::uword)n_rows, (arma::uword)n_cols, nnz, (arma::uword*)ri,
                        (arma::uword*)cp, vals);
  arma::Mat<T> Xm(X, (arma::uword)k, (arma::uword)n_src, false, true);
  arma::Mat<T> Ym(Y, (arma::uword)k, (arma::uword)n_cols, false, true);
  const arma::Col<T> cnt(cnt_X, (arma::uword)cnt_len, false, true);
  return (double)als_explicit<T>(Conf, Xm, Ym, lambda, n_threads, solver, cg_steps,
                                 dynamic_lambda != 0, cnt, with_biases != 0,
                                 is_x_bias_last_row != 0);
}

#endif

extern "C" {

#ifdef REF_IMPLICIT
double ref_als_implicit_f32(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                            float* X, int k, int n_src, float* Y, float* XtX, int k_xtx, double lambda,
                            int n_threads, unsigned solver, unsigned cg_steps, int with_biases,
                            int is_x_bias_last_row, double global_bias, float* gbb, int gbb_len,
                            int initialize_bias_base) {
  return run_implicit<float>(n_rows, n_cols, nnz, ri, cp, vals, X, k, n_src, Y, XtX, k_xtx, lambda,
                             n_threads, solver, cg_steps, with_biases, is_x_bias_last_row,
                             global_bias, gbb, gbb_len, initialize_bias_base);
}
double ref_als_implicit_f64(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                            double* X, int k, int n_src, double* Y, double* XtX, int k_xtx, double lambda,
                            int n_threads, unsigned solver, unsigned cg_steps, int with_biases,
                            int is_x_bias_last_row, double global_bias, double* gbb, int gbb_len,
                            int initialize_bias_base) {
  return run_implicit<double>(n_rows, n_cols, nnz, ri, cp, vals, X, k, n_src, Y, XtX, k_xtx, lambda,
                              n_threads, solver, cg_steps, with_biases, is_x_bias_last_row,
                              global_bias, gbb, gbb_len, initialize_bias_base);
}
double ref_initialize_biases_f32(int n_users, int n_items, size_t nnz, int* csc_ri, int* csc_cp, double* csc_v,
                                 int* csr_ri, int* csr_cp, double* csr_v, float* user_bias, float* item_bias,
                                 double lambda, int dynamic_lambda, int non_negative, int calc_global, int is_explicit) {
  return run_init_biases<float>(n_users, n_items, nnz, csc_ri, csc_cp, csc_v, csr_ri, csr_cp, csr_v, user_bias,
                                item_bias, lambda, dynamic_lambda, non_negative, calc_global, is_explicit);
}
double ref_initialize_biases_f64(int n_users, int n_items, size_t nnz, int* csc_ri, int* csc_cp, double* csc_v,
                                 int* csr_ri, int* csr_cp, double* csr_v, double* user_bias, double* item_bias,
                                 double lambda, int dynamic_lambda, int non_negative, int calc_global, int is_explicit) {
  return run_init_biases<double>(n_users, n_items, nnz, csc_ri, csc_cp, csc_v, csr_ri, csr_cp, csr_v, user_bias,
                                 item_bias, lambda, dynamic_lambda, non_negative, calc_global, is_explicit);
}
#endif
#ifdef REF_EXPLICIT
double ref_als_explicit_f32(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                            float* X, int k, int n_src, float* Y, float* cnt_X, int cnt_len,
                            double lambda, int n_threads, unsigned solver, unsigned cg_steps,
                            int dynamic_lambda, int with_biases, int is_x_bias_last_row) {
  return run_explicit<float>(n_rows, n_cols, nnz, ri, cp, vals, X, k, n_src, Y, cnt_X, cnt_len,
                             lambda, n_threads, solver, cg_steps, dynamic_lambda, with_biases,
                             is_x_bias_last_row);
}
double ref_als_explicit_f64(int n_rows, int n_cols, size_t nnz, int* ri, int* cp, double* vals,
                            double* X, int k, int n_src, double* Y, double* cnt_X, int cnt_len,
                            double lambda, int n_threads, unsigned solver, unsigned cg_steps,
                            int dynamic_lambda, int with_biases, int is_x_bias_last_row) {
  return run_explicit<double>(n_rows, n_cols, nnz, ri, cp, vals, X, k, n_src, Y, cnt_X, cnt_len,
                              lambda, n_threads, solver, cg_steps, dynamic_lambda, with_biases,
                              is_x_bias_last_row);
}

#endif

}  // extern "C"
