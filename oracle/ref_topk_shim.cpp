// TEST INFRASTRUCTURE ONLY.  extern "C" entry point around the reference's own, unmodified `top_product`
// (/root/reference/src/matrix_top_product.cpp:20-102), which build_ref.sh compiles IN PLACE against
// oracle/mini_rcpp + oracle/mini_arma into oracle/_ref/libref_topk.so.  This pins oracle/topk.py (tests/test_oracle.py)
// and generates tests/golden/topk.npz (tests/golden/make_golden_topk.py).
#include "rsparse.h"   // the reference's src/rsparse.h (via -I /root/reference/src); <RcppArmadillo.h> resolves to mini_rcpp

Rcpp::IntegerMatrix top_product(const arma::mat& x, const arma::mat& y, unsigned k, unsigned n_threads,
                                const Rcpp::S4& not_recommend_r, const Rcpp::IntegerVector& exclude, const double glob_mean);

// src/utils.cpp:58-67 of the reference on our slot-pointer handle
dMappedCSR extract_mapped_csr(Rcpp::S4 input) {
  return dMappedCSR((arma::uword)input.dim[0], (arma::uword)input.dim[1], input.nnz, (arma::uword*)input.j,
                    (arma::uword*)input.p, (double*)input.x);
}

// x: n_user x rank column-major, y: rank x n_item column-major (as R hands them over, R/utils.R:31-59);
// not_recommend: dgRMatrix slots (p, j), may be empty (nnz = 0); exclude: 1-based item ids.
// out_idx / out_scores: n_user x k column-major (NA_integer_ / NA_real_ where fewer than k items qualify).
extern "C" void ref_top_product(const double* x, int n_user, int rank, const double* y, int n_item, unsigned k,
                                unsigned n_threads, const int* nr_p, const int* nr_j, const double* nr_x, size_t nr_nnz,
                                const int* exclude, int n_exclude, double glob_mean, int* out_idx, double* out_scores) {
  const arma::mat xm(const_cast<double*>(x), (arma::uword)n_user, (arma::uword)rank, false, true);
  const arma::mat ym(const_cast<double*>(y), (arma::uword)rank, (arma::uword)n_item, false, true);
  Rcpp::S4 nr;
  nr.p = nr_p; nr.j = nr_j; nr.x = nr_x; nr.nnz = nr_nnz; nr.dim[0] = n_user; nr.dim[1] = n_item;
  const Rcpp::IntegerVector ex(exclude, (size_t)n_exclude);
  Rcpp::IntegerMatrix res = top_product(xm, ym, k, n_threads, nr, ex, glob_mean);
  std::memcpy(out_idx, res.begin(), sizeof(int) * (size_t)n_user * k);
  std::memcpy(out_scores, res.scores_attr.begin(), sizeof(double) * (size_t)n_user * k);
}
