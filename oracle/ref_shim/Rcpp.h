// Rcpp.h -- test-infrastructure stand-in for the Rcpp header, just large enough to compile the reference's own
// /root/reference/src/code.cpp UNCHANGED (it only needs Rcpp::NumericMatrix: construction from dimensions, nrow(), ncol()
// and operator()(row, col) on a column-major buffer, like an R numeric matrix).  Nothing here is product code: the object
// built with it (oracle/_ref/libcode_ref.so, recipe oracle/Makefile) is the checker the oracle and the CUDA builders are
// pinned against in tests/test_code_ref.py.
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

namespace Rcpp {
class NumericMatrix {
 public:
  NumericMatrix(int nrow, int ncol) : nr_(nrow), nc_(ncol), own_(new std::vector<double>((std::size_t)nrow * ncol, 0.0)), p_(own_->data()) {}
  // view on caller-owned column-major memory (what an SEXP of type REALSXP with a dim attribute is)
  NumericMatrix(double* data, int nrow, int ncol) : nr_(nrow), nc_(ncol), p_(data) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double& operator()(int r, int c) { return p_[(std::size_t)c * nr_ + r]; }
  const double& operator()(int r, int c) const { return p_[(std::size_t)c * nr_ + r]; }
  const double* data() const { return p_; }

 private:
  int nr_, nc_;
  std::shared_ptr<std::vector<double>> own_;
  double* p_;
};
}  // namespace Rcpp
