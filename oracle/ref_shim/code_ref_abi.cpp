// C ABI around the reference's five builders (/root/reference/src/code.cpp:6-131, registered as .Call routines in
// /root/reference/src/init.c:17-24).  The reference source is #included from where it lies -- it is never copied into this
// repository -- and compiled against the stand-in Rcpp.h of this directory.  Test infrastructure only.
#include "Rcpp.h"
#include REFERENCE_CODE_CPP

#include <cstring>

static int copy_out(const Rcpp::NumericMatrix& m, double* out) {
  std::memcpy(out, m.data(), sizeof(double) * (std::size_t)m.nrow() * m.ncol());
  return 0;
}
#define WRAP(call) try { return copy_out(call, out); } catch (const std::exception&) { return 1; }

extern "C" {
int ref_cpp_dist(double* m1, int n1, double* m2, int n2, int d1, int d2, double* out) {
  WRAP(cpp_dist(Rcpp::NumericMatrix(m1, n1, d1), Rcpp::NumericMatrix(m2, n2, d2)))
}
int ref_cpp_prod(double* m1, int n1, double* m2, int n2, int d1, int d2, double* out) {
  WRAP(cpp_prod(Rcpp::NumericMatrix(m1, n1, d1), Rcpp::NumericMatrix(m2, n2, d2)))
}
int ref_cpp_perio(double* m1, int n1, double* m2, int n2, int d1, int d2, double period, double* out) {
  WRAP(cpp_perio(Rcpp::NumericMatrix(m1, n1, d1), Rcpp::NumericMatrix(m2, n2, d2), period))
}
int ref_cpp_perio_deriv(double* m1, int n1, double* m2, int n2, int d1, int d2, double period, double* out) {
  WRAP(cpp_perio_deriv(Rcpp::NumericMatrix(m1, n1, d1), Rcpp::NumericMatrix(m2, n2, d2), period))
}
int ref_cpp_noise(double* m1, int n1, double* m2, int n2, int d1, int d2, double noise, double* out) {
  WRAP(cpp_noise(Rcpp::NumericMatrix(m1, n1, d1), Rcpp::NumericMatrix(m2, n2, d2), noise))
}
}
