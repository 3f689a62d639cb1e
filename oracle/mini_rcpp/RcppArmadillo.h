// mini_rcpp -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// The few Rcpp types /root/reference/src/matrix_top_product.cpp touches, restated so that file compiles UNMODIFIED,
// in place, against oracle/mini_arma (oracle/build_ref.sh -> oracle/_ref/libref_topk.so): dense R vectors / matrices
// as owning std::vector wrappers (column-major like R), NA constants with R's values (NA_integer_ = INT_MIN,
// NA_real_ = the NaN with payload 1954), an `attr("scores") = ...` slot on IntegerMatrix, and an S4 handle that just
// carries the dgRMatrix slot pointers the shim's extract_mapped_csr() reads (src/utils.cpp:58-67 in the reference).
#ifndef MINI_RCPP_HPP
#define MINI_RCPP_HPP
#include <armadillo>
#include <climits>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

namespace Rcpp {

template <class T>
class VecBase {
 public:
  std::vector<T> d;
  typedef const T* const_iterator;
  typedef T* iterator;
  VecBase() {}
  explicit VecBase(size_t n) : d(n, T(0)) {}
  VecBase(const T* p, size_t n) : d(p, p + n) {}
  T* begin() { return d.data(); }
  T* end() { return d.data() + d.size(); }
  const T* begin() const { return d.data(); }
  const T* end() const { return d.data() + d.size(); }
  long length() const { return (long)d.size(); }
  T& operator[](size_t i) { return d[i]; }
  const T& operator[](size_t i) const { return d[i]; }
};

class NumericVector : public VecBase<double> {
 public:
  using VecBase<double>::VecBase;
  static double get_na() {   // R's NA_real_: quiet NaN whose low word is 1954
    const unsigned long long bits = 0x7FF00000000007A2ull;
    double v;
    std::memcpy(&v, &bits, sizeof(v));
    return v;
  }
};
class IntegerVector : public VecBase<int> {
 public:
  using VecBase<int>::VecBase;
  static int get_na() { return INT_MIN; }
};
class NumericMatrix : public VecBase<double> {
 public:
  int nrow_ = 0, ncol_ = 0;
  NumericMatrix() {}
  NumericMatrix(int n, int m) : VecBase<double>((size_t)n * m), nrow_(n), ncol_(m) {}
};
class IntegerMatrix : public VecBase<int> {
 public:
  int nrow_ = 0, ncol_ = 0;
  NumericMatrix scores_attr;
  IntegerMatrix(int n, int m) : VecBase<int>((size_t)n * m), nrow_(n), ncol_(m) {}
  struct AttrProxy {
    IntegerMatrix& m;
    void operator=(const NumericMatrix& v) { m.scores_attr = v; }
  };
  AttrProxy attr(const char*) { return AttrProxy{*this}; }
};
// handle to a dgRMatrix: slot pointers only
class S4 {
 public:
  const int* j = nullptr;
  const int* p = nullptr;
  const double* x = nullptr;
  int dim[2] = {0, 0};
  size_t nnz = 0;
};

}  // namespace Rcpp
#endif
