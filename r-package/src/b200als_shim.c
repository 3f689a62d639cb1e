/* b200als_shim.c -- the `.Call` layer an rsparse maintainer drops into src/ in place of the four
 * Rcpp-generated ALS wrappers (src/RcppExports.cpp:329-416 -> src/wrmf_implicit.cpp:5-31,
 * src/wrmf_explicit.cpp:5-27).  Plain C against Rinternals.h only: no Rcpp, no Armadillo.
 * It unpacks the same SEXPs the reference unpacks (src/utils.cpp:69-78 for the dgCMatrix slots,
 * src/utils.cpp:115-128 for float32 S4 objects) into the C ABI of include/b200als.h and maps a
 * non-zero status to an R error, as END_RCPP does for C++ exceptions.
 *
 * The routines are registered under the reference's own symbol names (src/RcppExports.cpp:456-490),
 * so R/RcppExports.R:88-102 and R/model_WRMF.R:493-495,:514 keep working unchanged.
 *
 * NOT compiled in the authoring image (no R there); build: R CMD SHLIB b200als_shim.c -lb200als
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "b200als.h"

static void csc_from_s4(SEXP m, b200als_csc* out) { /* src/utils.cpp:69-78 */
  SEXP dim = R_do_slot(m, Rf_install("Dim"));
  SEXP x = R_do_slot(m, Rf_install("x"));
  SEXP i = R_do_slot(m, Rf_install("i"));
  SEXP p = R_do_slot(m, Rf_install("p"));
  out->n_rows = INTEGER(dim)[0];
  out->n_cols = INTEGER(dim)[1];
  out->nnz = (int64_t)XLENGTH(x);
  out->ptr = INTEGER(p);
  out->idx = INTEGER(i);
  out->val_f64 = REAL(x);
  out->val_f32 = NULL;
}

/* float32 S4 (package `float`): slot Data is an INTSXP matrix whose payload is IEEE float
 * (src/utils.cpp:115-128) */
static float* float32_matrix(SEXP s4, int* nrow, int* ncol) {
  SEXP d = R_do_slot(s4, Rf_install("Data"));
  SEXP dims = Rf_getAttrib(d, R_DimSymbol);
  if (nrow) *nrow = Rf_isNull(dims) ? (int)XLENGTH(d) : INTEGER(dims)[0];
  if (ncol) *ncol = Rf_isNull(dims) ? 1 : INTEGER(dims)[1];
  return (float*)INTEGER(d);
}
static float* float32_vector_or_null(SEXP s4) {
  SEXP d = R_do_slot(s4, Rf_install("Data"));
  return XLENGTH(d) ? (float*)INTEGER(d) : NULL;
}
/* global_bias_base: R sizes it rank-1 (R/model_WRMF.R:291-297) while als_implicit<T> reads and writes `rank` entries
 * (inst/include/wrmf_implicit.hpp:111-112, :155-157).  The engine takes the length the C++ code uses, so a shorter R
 * vector is staged through a buffer of the right length (zero padded) and copied back as far as it goes. */
static void* stage_bias_base(void* r_data, R_xlen_t r_len, int need, size_t elem) {
  if (r_len >= need) return r_data;
  char* buf = R_alloc((size_t)need, (int)elem);
  memset(buf, 0, (size_t)need * elem);
  if (r_len > 0) memcpy(buf, r_data, (size_t)r_len * elem);
  return buf;
}
static void check(int rc) {
  if (rc != B200ALS_OK) Rf_error("b200als: %s (status %d)", b200als_last_error(), rc);
}

SEXP _rsparse_als_implicit_float(SEXP m_csc_r, SEXP X_, SEXP Y_, SEXP XtX_, SEXP lambda, SEXP n_threads,
                                 SEXP solver, SEXP cg_steps, SEXP with_biases, SEXP is_x_bias_last_row,
                                 SEXP global_bias, SEXP global_bias_base_, SEXP initialize_bias_base) {
  b200als_csc A;
  csc_from_s4(m_csc_r, &A);
  int k = 0, kx = 0;
  float* X = float32_matrix(X_, &k, NULL);
  float* Y = float32_matrix(Y_, NULL, NULL); /* updated in place, like arma's aliasing fmat */
  float* XtX = float32_matrix(XtX_, &kx, NULL);
  double loss = 0.0;
  SEXP gd = R_do_slot(global_bias_base_, Rf_install("Data"));
  float* base_r = float32_vector_or_null(global_bias_base_);
  float* base = (float*)stage_bias_base(base_r, XLENGTH(gd), k, sizeof(float));
  check(b200als_als_implicit_float(&A, k, X, Y, XtX, Rf_asReal(lambda), Rf_asInteger(n_threads),
                                   (unsigned)Rf_asInteger(solver), (unsigned)Rf_asInteger(cg_steps),
                                   Rf_asLogical(with_biases), Rf_asLogical(is_x_bias_last_row),
                                   Rf_asReal(global_bias), base, Rf_asLogical(initialize_bias_base), &loss));
  if (base != base_r && XLENGTH(gd) > 0) memcpy(base_r, base, (size_t)XLENGTH(gd) * sizeof(float));
  return Rf_ScalarReal(loss);
}

SEXP _rsparse_als_implicit_double(SEXP m_csc_r, SEXP X, SEXP Y, SEXP XtX, SEXP lambda, SEXP n_threads,
                                  SEXP solver, SEXP cg_steps, SEXP with_biases, SEXP is_x_bias_last_row,
                                  SEXP global_bias, SEXP global_bias_base, SEXP initialize_bias_base) {
  b200als_csc A;
  csc_from_s4(m_csc_r, &A);
  double loss = 0.0;
  double* base_r = XLENGTH(global_bias_base) ? REAL(global_bias_base) : NULL;
  double* base = (double*)stage_bias_base(base_r, XLENGTH(global_bias_base), Rf_nrows(X), sizeof(double));
  check(b200als_als_implicit_double(&A, Rf_nrows(X), REAL(X), REAL(Y), REAL(XtX), Rf_asReal(lambda),
                                    Rf_asInteger(n_threads), (unsigned)Rf_asInteger(solver),
                                    (unsigned)Rf_asInteger(cg_steps), Rf_asLogical(with_biases),
                                    Rf_asLogical(is_x_bias_last_row), Rf_asReal(global_bias), base,
                                    Rf_asLogical(initialize_bias_base), &loss));
  if (base != base_r && XLENGTH(global_bias_base) > 0)
    memcpy(base_r, base, (size_t)XLENGTH(global_bias_base) * sizeof(double));
  return Rf_ScalarReal(loss);
}

SEXP _rsparse_als_explicit_float(SEXP m_csc_r, SEXP X_, SEXP Y_, SEXP cnt_X_, SEXP lambda, SEXP n_threads,
                                 SEXP solver, SEXP cg_steps, SEXP dynamic_lambda, SEXP with_biases,
                                 SEXP is_x_bias_last_row) {
  b200als_csc A;
  csc_from_s4(m_csc_r, &A);
  int k = 0;
  float* X = float32_matrix(X_, &k, NULL);
  float* Y = float32_matrix(Y_, NULL, NULL);
  double loss = 0.0;
  check(b200als_als_explicit_float(&A, k, X, Y, float32_vector_or_null(cnt_X_), Rf_asReal(lambda),
                                   (unsigned)Rf_asInteger(n_threads), (unsigned)Rf_asInteger(solver),
                                   (unsigned)Rf_asInteger(cg_steps), Rf_asLogical(dynamic_lambda),
                                   Rf_asLogical(with_biases), Rf_asLogical(is_x_bias_last_row), &loss));
  return Rf_ScalarReal(loss);
}

SEXP _rsparse_als_explicit_double(SEXP m_csc_r, SEXP X, SEXP Y, SEXP cnt_X, SEXP lambda, SEXP n_threads,
                                  SEXP solver, SEXP cg_steps, SEXP dynamic_lambda, SEXP with_biases,
                                  SEXP is_x_bias_last_row) {
  b200als_csc A;
  csc_from_s4(m_csc_r, &A);
  double loss = 0.0;
  check(b200als_als_explicit_double(&A, Rf_nrows(X), REAL(X), REAL(Y), XLENGTH(cnt_X) ? REAL(cnt_X) : NULL,
                                    Rf_asReal(lambda), (unsigned)Rf_asInteger(n_threads),
                                    (unsigned)Rf_asInteger(solver), (unsigned)Rf_asInteger(cg_steps),
                                    Rf_asLogical(dynamic_lambda), Rf_asLogical(with_biases),
                                    Rf_asLogical(is_x_bias_last_row), &loss));
  return Rf_ScalarReal(loss);
}

/* ---- device-resident session: what WRMF$fit_transform(precision = "float") binds to -------------- */
static void session_finalizer(SEXP ptr) {
  b200als_session* s = (b200als_session*)R_ExternalPtrAddr(ptr);
  if (s) b200als_destroy(s);
  R_ClearExternalPtr(ptr);
}
/* c_ui, c_iu: dgCMatrix (users x items as CSC; items x users as CSC == R/model_WRMF.R:184-189) */
SEXP b200als_R_create(SEXP c_ui, SEXP c_iu, SEXP rank, SEXP feedback, SEXP solver, SEXP cg_steps,
                      SEXP dynamic_lambda, SEXP lambda) {
  b200als_csc A, B;
  csc_from_s4(c_ui, &A);
  csc_from_s4(c_iu, &B);
  b200als_options o;
  b200als_default_options(&o);
  o.feedback = Rf_asInteger(feedback);
  o.solver = Rf_asInteger(solver);
  o.cg_steps = Rf_asInteger(cg_steps);
  o.dynamic_lambda = Rf_asLogical(dynamic_lambda);
  o.lambda = Rf_asReal(lambda);
  b200als_session* s = NULL;
  check(b200als_create(&s, &A, &B, A.n_rows, A.n_cols, Rf_asInteger(rank), &o));
  SEXP ptr = PROTECT(R_MakeExternalPtr(s, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, session_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
SEXP b200als_R_set_factors(SEXP ptr, SEXP which, SEXP fl) {
  check(b200als_set_factors((b200als_session*)R_ExternalPtrAddr(ptr), Rf_asInteger(which), float32_matrix(fl, NULL, NULL)));
  return R_NilValue;
}
SEXP b200als_R_get_factors(SEXP ptr, SEXP which, SEXP fl_out) {
  check(b200als_get_factors((b200als_session*)R_ExternalPtrAddr(ptr), Rf_asInteger(which), float32_matrix(fl_out, NULL, NULL)));
  return R_NilValue;
}
/* bias terms inside the session (R/model_WRMF.R:260-297): call after initialize_biases, before the fit */
SEXP b200als_R_set_bias(SEXP ptr, SEXP with_user_item_bias, SEXP global_bias) {
  check(b200als_set_bias((b200als_session*)R_ExternalPtrAddr(ptr), Rf_asLogical(with_user_item_bias), Rf_asReal(global_bias)));
  return R_NilValue;
}
SEXP b200als_R_fit(SEXP ptr, SEXP n_iter, SEXP tol) {
  const int n = Rf_asInteger(n_iter);
  SEXP trace = PROTECT(Rf_allocVector(REALSXP, 2 * (n > 0 ? n : 1)));
  int done = 0;
  check(b200als_fit((b200als_session*)R_ExternalPtrAddr(ptr), n, Rf_asReal(tol), REAL(trace), &done));
  SEXP out = PROTECT(Rf_lengthgets(trace, 2 * done));
  UNPROTECT(2);
  return out;
}
SEXP b200als_R_transform(SEXP ptr, SEXP fl_out) {
  double loss = 0.0;
  check(b200als_transform((b200als_session*)R_ExternalPtrAddr(ptr), float32_matrix(fl_out, NULL, NULL), &loss));
  return Rf_ScalarReal(loss);
}

/* initialize_biases_{double,float} (src/wrmf_init.cpp:6-34; R stubs R/RcppExports.R, caller R/model_WRMF.R:150-155):
 * m_csc_r = c_ui (user x item, dgCMatrix), m_csr_r = c_iu = t_shallow(csr(c_ui)), i.e. the same entries by user.
 * The @x slots are modified in place for explicit feedback with calculate_global_bias, as in the reference. */
SEXP _rsparse_initialize_biases_double(SEXP m_csc_r, SEXP m_csr_r, SEXP user_bias, SEXP item_bias, SEXP lambda,
                                       SEXP dynamic_lambda, SEXP non_negative, SEXP calculate_global_bias,
                                       SEXP is_explicit_feedback) {
  b200als_csc C, R_;
  csc_from_s4(m_csc_r, &C);
  csc_from_s4(m_csr_r, &R_);
  double g = 0.0;
  check(b200als_initialize_biases_double(R_.n_cols, C.n_cols, C.nnz, C.ptr, C.idx, (double*)C.val_f64, R_.ptr, R_.idx,
                                         (double*)R_.val_f64, REAL(user_bias), REAL(item_bias), Rf_asReal(lambda),
                                         Rf_asLogical(dynamic_lambda), Rf_asLogical(non_negative),
                                         Rf_asLogical(calculate_global_bias), Rf_asLogical(is_explicit_feedback), &g));
  return Rf_ScalarReal(g);
}
SEXP _rsparse_initialize_biases_float(SEXP m_csc_r, SEXP m_csr_r, SEXP user_bias, SEXP item_bias, SEXP lambda,
                                      SEXP dynamic_lambda, SEXP non_negative, SEXP calculate_global_bias,
                                      SEXP is_explicit_feedback) {
  b200als_csc C, R_;
  csc_from_s4(m_csc_r, &C);
  csc_from_s4(m_csr_r, &R_);
  double g = 0.0;
  check(b200als_initialize_biases_float(R_.n_cols, C.n_cols, C.nnz, C.ptr, C.idx, (double*)C.val_f64, R_.ptr, R_.idx,
                                        (double*)R_.val_f64, float32_matrix(user_bias, NULL, NULL),
                                        float32_matrix(item_bias, NULL, NULL), Rf_asReal(lambda),
                                        Rf_asLogical(dynamic_lambda), Rf_asLogical(non_negative),
                                        Rf_asLogical(calculate_global_bias), Rf_asLogical(is_explicit_feedback), &g));
  return Rf_ScalarReal(g);
}

/* top_product (src/matrix_top_product.cpp:20-102; Rcpp export src/RcppExports.cpp:188-203, 7 arguments; R caller
 * find_top_product, R/utils.R:31-59).  R hands over doubles (float models are widened at R/utils.R:35-36, so narrowing
 * them back is exact; a precision = "double" model is rounded to float here -- the engine's scores are still accumulated
 * in double like the reference's, but the ranking of near-ties may then differ).  x is n_user x rank, y is rank x n_item;
 * the result is the reference's: an n_user x k IntegerMatrix of 1-based item ids (NA padded) with attribute "scores". */
SEXP _rsparse_top_product(SEXP x, SEXP y, SEXP k, SEXP n_threads, SEXP not_recommend_r, SEXP exclude, SEXP glob_mean) {
  const int n_user = Rf_nrows(x), rank = Rf_ncols(x), n_item = Rf_ncols(y), top_k = Rf_asInteger(k);
  if (Rf_nrows(y) != rank) Rf_error("b200als: ncol(x) == nrow(y) is not TRUE");
  (void)n_threads;
  float* xe = (float*)R_alloc((size_t)rank * (size_t)(n_user > 0 ? n_user : 1), sizeof(float));   /* rank x n_user */
  float* ye = (float*)R_alloc((size_t)rank * (size_t)(n_item > 0 ? n_item : 1), sizeof(float));   /* rank x n_item */
  const double* xd = REAL(x);
  const double* yd = REAL(y);
  for (int u = 0; u < n_user; u++)
    for (int f = 0; f < rank; f++) xe[(size_t)u * rank + f] = (float)xd[(size_t)f * n_user + u];
  for (size_t e = 0; e < (size_t)rank * (size_t)n_item; e++) ye[e] = (float)yd[e];
  SEXP nj = R_do_slot(not_recommend_r, Rf_install("j"));   /* dgRMatrix: @p row pointers, @j column indices */
  SEXP np = R_do_slot(not_recommend_r, Rf_install("p"));
  const int have_filter = XLENGTH(nj) > 0;
  SEXP res = PROTECT(Rf_allocMatrix(INTSXP, n_user, top_k));
  SEXP scores = PROTECT(Rf_allocMatrix(REALSXP, n_user, top_k));
  check(b200als_top_product(xe, n_user, ye, n_item, rank, top_k, have_filter ? INTEGER(np) : NULL,
                            have_filter ? INTEGER(nj) : NULL, INTEGER(exclude), (int)XLENGTH(exclude), Rf_asReal(glob_mean),
                            INTEGER(res), REAL(scores)));
  Rf_setAttrib(res, Rf_install("scores"), scores);
  UNPROTECT(2);
  return res;
}

static const R_CallMethodDef CallEntries[] = {
    {"_rsparse_als_implicit_float", (DL_FUNC)&_rsparse_als_implicit_float, 13},
    {"_rsparse_als_implicit_double", (DL_FUNC)&_rsparse_als_implicit_double, 13},
    {"_rsparse_als_explicit_float", (DL_FUNC)&_rsparse_als_explicit_float, 11},
    {"_rsparse_als_explicit_double", (DL_FUNC)&_rsparse_als_explicit_double, 11},
    {"_rsparse_initialize_biases_double", (DL_FUNC)&_rsparse_initialize_biases_double, 9},
    {"_rsparse_initialize_biases_float", (DL_FUNC)&_rsparse_initialize_biases_float, 9},
    {"_rsparse_top_product", (DL_FUNC)&_rsparse_top_product, 7},
    {"b200als_R_create", (DL_FUNC)&b200als_R_create, 8},
    {"b200als_R_set_factors", (DL_FUNC)&b200als_R_set_factors, 3},
    {"b200als_R_get_factors", (DL_FUNC)&b200als_R_get_factors, 3},
    {"b200als_R_set_bias", (DL_FUNC)&b200als_R_set_bias, 3},
    {"b200als_R_fit", (DL_FUNC)&b200als_R_fit, 3},
    {"b200als_R_transform", (DL_FUNC)&b200als_R_transform, 2},
    {NULL, NULL, 0}};

void R_init_rsparse(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
