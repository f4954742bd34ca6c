// linalg.cuh -- block-cooperative FP64 dense routines on one ROW-major matrix with leading
// dimension ld, callable on shared memory (per-task matrices, N <= 96) or on global/L2
// memory (one CTA per union-grid matrix).  Every thread of the CTA must call them with the
// same arguments; they start and end with the data visible to the whole CTA.
//
// They stand for the LAPACK calls under the reference's base-R functions:
//   chol()      -> dpotrf('U')   kernel-to-matrices.R:675   (reads the upper triangle only;
//                                fails on a pivot <= 0 or NaN like dpotf2)
//   chol2inv()  -> dpotri('U') + symmetrise                 kernel-to-matrices.R:675
//   determinant -> dgetrf (partial pivoting) log|det|       likelihoods.R:37-41
#pragma once
#include "common.cuh"

// In-place upper Cholesky A = R'R (right-looking).  Returns 0, or j+1 when pivot j fails.
// Only the upper triangle is read and written.
__device__ inline int blk_chol_upper(double* A, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < n; ++j) {
    double ajj = A[j * ld + j];
    if (!(ajj > 0.0)) return j + 1;  // uniform: every thread reads the same value
    double rjj = sqrt(ajj);
    __syncthreads();
    if (tid == 0) A[j * ld + j] = rjj;
    for (int k = j + 1 + tid; k < n; k += nt) A[j * ld + k] = A[j * ld + k] / rjj;
    __syncthreads();
    const int m = n - j - 1;
    for (int e = tid; e < m * m; e += nt) {
      int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k >= i) A[i * ld + k] -= A[j * ld + i] * A[j * ld + k];
    }
    __syncthreads();
  }
  return 0;
}

// X = R^{-1} for upper-triangular R (upper triangle of Rm), X upper triangular, lower part of
// X zeroed.  Row-sweep back substitution on all right-hand sides at once.
__device__ inline void blk_tri_inv_upper(const double* Rm, double* X, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n * n; e += nt) {
    int i = e / n, j = e % n;
    X[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = n - 1; k >= 0; --k) {
    double rkk = Rm[k * ld + k];
    for (int j = k + tid; j < n; j += nt) X[k * ld + j] = X[k * ld + j] / rkk;
    __syncthreads();
    const int w = n - k;
    for (int e = tid; e < k * w; e += nt) {
      int i = e / w, j = k + e % w;
      X[i * ld + j] -= Rm[i * ld + k] * X[k * ld + j];
    }
    __syncthreads();
  }
}

// A = X X' for upper-triangular X (dlauum), both triangles written (chol2inv symmetrises).
__device__ inline void blk_lauum_upper(const double* X, double* A, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n * n; e += nt) {
    int i = e / n, j = e % n;
    if (j >= i) {
      double s = 0.0;
      for (int k = j; k < n; ++k) s += X[i * ld + k] * X[j * ld + k];
      A[i * ld + j] = s;
      A[j * ld + i] = s;
    }
  }
  __syncthreads();
}

// log|det B| by LU with partial pivoting (implicit row permutation; B is destroyed).
// `done` is an n-int scratch array visible to the CTA.  Every thread returns the same value.
__device__ inline double blk_lu_logabsdet(double* B, int n, int ld, int* done) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  for (int i = tid; i < n; i += nt) done[i] = 0;
  __syncthreads();
  double logdet = 0.0;
  for (int j = 0; j < n; ++j) {
    // every warp finds the pivot row redundantly: max |B[i,j]| over rows not yet used,
    // lowest row index on ties
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      if (!done[i]) {
        double v = fabs(B[i * ld + j]);
        if (v != v) v = INFINITY;  // NaN poisons the determinant like idamax would pick it
        if (v > best) { best = v; bi = i; }
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const int p = bi;
    const double piv = B[p * ld + j];
    logdet += log(fabs(piv));
    __syncthreads();  // all warps have read done[] and column j before anything changes
    if (tid == 0) done[p] = 1;
    if (piv != 0.0) {
      const double rcp = 1.0 / piv;
      const int w = n - j - 1;
      for (int e = tid; e < n * w; e += nt) {
        int i = e / w, k = j + 1 + e % w;
        if (i != p && !done[i]) B[i * ld + k] -= (B[i * ld + j] * rcp) * B[p * ld + k];
      }
    }
    __syncthreads();
  }
  return logdet;
}

// C = A * B (all n x n, row-major, same ld); C must not alias A or B.
__device__ inline void blk_gemm_nn(double* C, const double* A, const double* B, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nb = (n + 3) / 4;
  for (int e = tid; e < nb * nb; e += nt) {
    const int i0 = (e / nb) * 4, j0 = (e % nb) * 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int k = 0; k < n; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = (i0 + a < n) ? A[(i0 + a) * ld + k] : 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = (j0 + b < n) ? B[k * ld + j0 + b] : 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bv[b];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (i0 + a < n && j0 + b < n) C[(i0 + a) * ld + j0 + b] = acc[a][b];
  }
  __syncthreads();
}

// y = A x (n x n); y must not alias x.
__device__ inline void blk_gemv(double* y, const double* A, const double* x, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += A[i * ld + k] * x[k];
    y[i] = s;
  }
  __syncthreads();
}

// Deterministic CTA-wide sum of NV values per thread; result valid in every thread.
// red: scratch of NV * 32 doubles.
template <int NV>
__device__ inline void blk_reduce_sum(double* v, double* red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = v[q];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) red[q * 32 + warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[q * 32 + w];
    v[q] = s;
  }
  __syncthreads();
}
