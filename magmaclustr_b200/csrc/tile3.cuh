// tile3.cuh -- swizzled 8x8 FP64 tiles in shared memory and the warp-level pieces of the blocked
// Cholesky built on them (shared by the per-task kernels, tasks3.cu, and the union-grid kernels, dense3.cu).
#pragma once
#include "tiles.cuh"

// swizzled position of element (r, c) inside an 8x8 tile
__device__ __forceinline__ int sw(int r, int c) { return (r << 3) + (c ^ ((r & 2) << 1)); }

struct Ln {  // per-lane constants
  int warp, lane, g, t, o_gt, o_tg, o_c;
  __device__ __forceinline__ Ln() {
    warp = threadIdx.x >> 5; lane = threadIdx.x & 31; g = lane >> 2; t = lane & 3;
    o_gt = sw(g, t); o_tg = sw(t, g); o_c = sw(g, 2 * t);
  }
};

// row fragment (P[g][t], P[g][t+4]): A operand of P, or B operand of P'
__device__ __forceinline__ void rf(const double* P, const Ln& L, double& x0, double& x1) { x0 = P[L.o_gt]; x1 = P[L.o_gt ^ 4]; }
// column fragment (P[t][g], P[t+4][g]): B operand of P, or A operand of P'
__device__ __forceinline__ void cf(const double* P, const Ln& L, double& x0, double& x1) { x0 = P[L.o_tg]; x1 = P[L.o_tg + 32]; }

// One warp: eliminate the row panel [A_kk | A_k,k+1 | A_k,k+2 | A_k,k+3], lane c owning column c & 7 of
// tile k + (c >> 3).  dpotf2 rule: fail on a pivot <= 0 or NaN.
// Every lane first factors the 8 x 8 diagonal tile FOR ITSELF in registers (36 broadcast loads, 84 FMAs with
// plenty of instruction-level parallelism, no shuffles), then runs the forward substitution R_kk^{-T} a on
// its own column -- for the columns of the diagonal tile this reproduces R_kk, for the panel tiles it is
// R_kj = R_kk^{-T} A_kj.  The earlier version passed every multiplier through a 64-bit warp shuffle: 8 x 7
// shuffle pairs on the serial pivot chain of the factorisation; the arithmetic (and hence every bit of the
// result) is the same.
// Returns 0 or the failing pivot (1-based), uniform across the warp.
// Pkk: tile (k, k); the next tiles of row k follow contiguously; navail = min(4, n8 - k); rinv8 receives
// the reciprocals of the eight pivots.
__device__ __forceinline__ int warp_factor_panel(double* Pkk, int navail, double* rinv8, int lane) {
  const int cb = lane >> 3, cc = lane & 7;
  const bool act = cb < navail;
  double R[8][8], ri[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i; j < 8; ++j) R[i][j] = Pkk[sw(i, j)];
  int fail = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double ajj = R[j][j];
    if (!(ajj > 0.0) && fail == 0) fail = j + 1;
    const double rj = rsqrt(ajj);
    ri[j] = rj;
#pragma unroll
    for (int c = j + 1; c < 8; ++c) R[j][c] *= rj;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int c = i; c < 8; ++c) R[i][c] -= R[j][i] * R[j][c];
  }
  if (fail) return fail;
  double* P = Pkk + (size_t)(act ? cb : 0) * 64;
  double a[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) a[r] = P[sw(r, cc)];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] *= ri[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i] -= R[j][i] * a[j];
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) P[sw(r, cc)] = (cb == 0 && r > cc) ? 0.0 : a[r];
  }
  if (lane < 8) {
    double v = ri[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) v = (lane == q) ? ri[q] : v;
    rinv8[lane] = v;
  }
  __syncwarp();
  return 0;
}

// X = R^{-1} of one upper-triangular 8x8 tile (lower part zero); any warp, lanes c & 7 own column c
__device__ __forceinline__ void warp_diag_inverse(const double* Rt, const double* rinv8, double* Xt, int lane) {
  const int c = lane & 7;
  double x[8];
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    double s = 0.0;
#pragma unroll
    for (int m = i + 1; m < 8; ++m) s += Rt[sw(i, m)] * x[m];
    const double ri = rinv8[i];
    x[i] = (i > c) ? 0.0 : ((i == c) ? ri : -s * ri);
  }
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) Xt[sw(r, c)] = x[r];
  }
  __syncwarp();
}


// Z = R_kk^{-T} A for up to four contiguous tiles (32 columns) by forward substitution, lane c owning column
// c & 7 of tile c >> 3 -- the panel step of dpotrf as a true triangular solve (what dtrsm does), not a product
// with an explicit inverse: on the ill-conditioned union-grid matrices (cond ~ 1e12) the latter loses
// cond(R_kk) more digits.  Rkk: factored diagonal tile, rinv8: reciprocals of its diagonal.
__device__ __forceinline__ void warp_panel_solve(const double* Rkk, const double* rinv8, double* P, int ntiles, int lane) {
  const int cb = lane >> 3, cc = lane & 7;
  if (cb < ntiles) {
    double* T = P + (size_t)cb * 64;
    double a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = T[sw(r, cc)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] *= rinv8[j];
#pragma unroll
      for (int i = j + 1; i < 8; ++i) a[i] -= Rkk[sw(j, i)] * a[j];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) T[sw(r, cc)] = a[r];
  }
  __syncwarp();
}

// X = R_ii^{-1} W for up to four contiguous tiles by back substitution (same lane mapping)
__device__ __forceinline__ void warp_row_backsolve(const double* Rii, const double* rinv8, double* P, int ntiles, int lane) {
  const int cb = lane >> 3, cc = lane & 7;
  if (cb < ntiles) {
    double* T = P + (size_t)cb * 64;
    double a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = T[sw(r, cc)];
#pragma unroll
    for (int r = 7; r >= 0; --r) {
      a[r] *= rinv8[r];
#pragma unroll
      for (int i = 0; i < r; ++i) a[i] -= Rii[sw(i, r)] * a[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) T[sw(r, cc)] = a[r];
  }
  __syncwarp();
}
