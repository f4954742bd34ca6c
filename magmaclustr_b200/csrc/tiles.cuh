// tiles.cuh -- 8x8-tile building blocks on the FP64 tensor cores (mma.sync.m8n8k4.f64 -> DMMA)
// and the blocked Cholesky / triangular inverse / X X' built from them.  Used on shared memory
// by the per-task kernels (tasks2.cu, 8 warps) and on global/L2 memory by the union-grid kernels
// (dense.cu, 32 warps).  All matrices are row-major, padded to a multiple of 8 with an identity
// block; NW = warps per CTA (blockDim.x == 32 * NW).
#pragma once
#include "common.cuh"

#define FULLMASK 0xffffffffu
#define XD_LD 12
#define XD_STRIDE (8 * XD_LD)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// c += op(At) * op(Bt) for 8x8 tiles at Am / Bm (row-major, leading dims lda / ldb).
// A operand element (row g, k) = TA ? Am[k][g] : Am[g][k]; B operand (k, col g) = TB ? Bm[g][k] : Bm[k][g].
template <bool TA, bool TB, bool NEG>
__device__ __forceinline__ void tile_mma(double& c0, double& c1, const double* Am, int lda,
                                         const double* Bm, int ldb, int g, int t) {
  double a0 = TA ? Am[t * lda + g] : Am[g * lda + t];
  double a1 = TA ? Am[(t + 4) * lda + g] : Am[g * lda + t + 4];
  double b0 = TB ? Bm[g * ldb + t] : Bm[t * ldb + g];
  double b1 = TB ? Bm[g * ldb + t + 4] : Bm[(t + 4) * ldb + g];
  if (NEG) { a0 = -a0; a1 = -a1; }
  dmma884(c0, c1, a0, b0);
  dmma884(c0, c1, a1, b1);
}


// idx-th tile (row-major) of the upper triangle of an m x m tile matrix -> (i, j), i <= j
__device__ __forceinline__ void tri_ij(int idx, int m, int& i, int& j) {
  const float fm = (float)(2 * m + 1);
  int ii = (int)((fm - sqrtf(fm * fm - 8.0f * (float)idx)) * 0.5f);
  int s = (ii * (2 * m - ii + 1)) >> 1;
  if (s > idx) { --ii; s = (ii * (2 * m - ii + 1)) >> 1; }
  else if (idx - s >= m - ii) { s += m - ii; ++ii; }
  i = ii;
  j = ii + (idx - s);
}

// Factor the 8x8 diagonal tile at Tp in place (upper Cholesky, LAPACK dpotf2 rule: fail on a
// pivot <= 0 or NaN; rows are scaled by the reciprocal like dpotf2's DSCAL) and write
// X = R^{-1} (upper, lower part zero) to Xk (8 x XD_LD).  One warp; lane c&7 owns column c.
// Returns 0 or the failing pivot (1-based), uniform across the warp.
__device__ __forceinline__ int warp_diag_factor(double* Tp, int ld, double* Xk, int lane) {
  const int c = lane & 7;
  double a[8], rinv[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) a[r] = Tp[r * ld + c];
  int fail = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double ajj = __shfl_sync(FULLMASK, a[j], j);
    if (!(ajj > 0.0) && fail == 0) fail = j + 1;
    double rjj = sqrt(ajj);
    double ri = 1.0 / rjj;
    rinv[j] = ri;
    a[j] = (c == j) ? rjj : a[j] * ri;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double rji = __shfl_sync(FULLMASK, a[j], i);
      if (c >= i) a[i] -= rji * a[j];
    }
  }
  if (fail) return fail;
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) Tp[r * ld + c] = (r <= c) ? a[r] : 0.0;
  }
  // every lane gathers the strict upper part of R, then solves its own column of X
  double x[8];
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    double s = 0.0;
#pragma unroll
    for (int m = i + 1; m < 8; ++m) {
      double rim = __shfl_sync(FULLMASK, a[i], m);
      if (m <= c) s += rim * x[m];
    }
    x[i] = (i > c) ? 0.0 : ((i == c) ? rinv[i] : -s * rinv[i]);
  }
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) Xk[r * XD_LD + c] = x[r];
  }
  return 0;
}

// Blocked upper Cholesky of A (np x np) in place; Xd receives the inverses of the diagonal
// tiles.  Returns 0 or a non-zero code (uniform) when a pivot fails.
template <int NW>
__device__ inline int chol_tiles(double* A, double* Xd, int n8, int ld, int* flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) *flag = 0;
  __syncthreads();
  if (warp == 0) {
    int f = warp_diag_factor(A, ld, Xd, lane);
    if (f && lane == 0) *flag = f;
  }
  __syncthreads();
  for (int k = 0; k < n8; ++k) {
    if (*flag) return *flag;
    // panel: R_kj = R_kk^{-T} A_kj = Xkk' A_kj
    for (int j = k + 1 + warp; j < n8; j += NW) {
      double c0 = 0.0, c1 = 0.0;
      double* Tkj = A + (8 * k) * ld + 8 * j;
      tile_mma<true, false, false>(c0, c1, Xd + k * XD_STRIDE, XD_LD, Tkj, ld, g, t);
      __syncwarp();
      *reinterpret_cast<double2*>(Tkj + g * ld + 2 * t) = make_double2(c0, c1);
    }
    __syncthreads();
    // trailing update A_ij -= R_ki' R_kj (k < i <= j); the warp that owns the next diagonal
    // tile factors it right away (look-ahead)
    // warp 0 takes the next diagonal tile (and factors it); the other warps share the rest
    const int m = n8 - k - 1, ntr = (m * (m + 1)) >> 1;
    for (int idx = (warp == 0) ? 0 : warp; idx < ntr; idx += (warp == 0) ? ntr : (NW - 1)) {
      int i, j;
      tri_ij(idx, m, i, j);
      i += k + 1;
      j += k + 1;
      {
        double* Tij = A + (8 * i) * ld + 8 * j;
        double2 cv = *reinterpret_cast<double2*>(Tij + g * ld + 2 * t);
        tile_mma<true, false, true>(cv.x, cv.y, A + (8 * k) * ld + 8 * i, ld, A + (8 * k) * ld + 8 * j, ld, g, t);
        *reinterpret_cast<double2*>(Tij + g * ld + 2 * t) = cv;
        if (idx == 0) {
          __syncwarp();
          int f = warp_diag_factor(Tij, ld, Xd + (k + 1) * XD_STRIDE, lane);
          if (f && lane == 0) *flag = 8 * (k + 1) + f;
        }
      }
    }
    __syncthreads();
  }
  return *flag;
}

// X = R^{-1} into Bm (upper tiles; tiles below the diagonal are not touched and never read).
// Column block j is independent of the others: X_jj = Xd_j, X_ij = -X_ii sum_{i<m<=j} R_im X_mj.
template <int NW>
__device__ inline void trtri_tiles(const double* A, const double* Xd, double* Bm, int n8, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // heavier column blocks first
  for (int jj = warp; jj < n8; jj += NW) {
    const int j = n8 - 1 - jj;
    double* Xjj = Bm + (8 * j) * ld + 8 * j;
    for (int e = lane; e < 64; e += 32) Xjj[(e >> 3) * ld + (e & 7)] = Xd[j * XD_STRIDE + (e >> 3) * XD_LD + (e & 7)];
    __syncwarp();
    for (int i = j - 1; i >= 0; --i) {
      double s0 = 0.0, s1 = 0.0;
      for (int m = i + 1; m <= j; ++m)
        tile_mma<false, false, false>(s0, s1, A + (8 * i) * ld + 8 * m, ld, Bm + (8 * m) * ld + 8 * j, ld, g, t);
      double* Xij = Bm + (8 * i) * ld + 8 * j;
      *reinterpret_cast<double2*>(Xij + g * ld + 2 * t) = make_double2(s0, s1);
      __syncwarp();
      double c0 = 0.0, c1 = 0.0;
      tile_mma<false, false, true>(c0, c1, Xd + i * XD_STRIDE, XD_LD, Xij, ld, g, t);
      __syncwarp();
      *reinterpret_cast<double2*>(Xij + g * ld + 2 * t) = make_double2(c0, c1);
      __syncwarp();
    }
  }
  __syncthreads();
}

// A = X X' (X upper triangular in Bm), both triangles written.
template <int NW>
__device__ inline void lauum_tiles(const double* Bm, double* A, int n8, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (n8 * (n8 + 1)) >> 1;
  for (int idx = warp; idx < ntiles; idx += NW) {
    int i, j;
    tri_ij(idx, n8, i, j);
    {
      double c0 = 0.0, c1 = 0.0;
      for (int m = j; m < n8; ++m)
        tile_mma<false, true, false>(c0, c1, Bm + (8 * i) * ld + 8 * m, ld, Bm + (8 * j) * ld + 8 * m, ld, g, t);
      *reinterpret_cast<double2*>(A + (8 * i + g) * ld + 8 * j + 2 * t) = make_double2(c0, c1);
      if (i != j) {
        A[(8 * j + 2 * t) * ld + 8 * i + g] = c0;
        A[(8 * j + 2 * t + 1) * ld + 8 * i + g] = c1;
      }
    }
  }
  __syncthreads();
  // diagonal tiles: mirror the upper part so the matrix is exactly symmetric (chol2inv)
  for (int e = threadIdx.x; e < n8 * 64; e += blockDim.x) {
    int k = e >> 6, r = (e >> 3) & 7, c = e & 7;
    if (r > c) A[(8 * k + r) * ld + 8 * k + c] = A[(8 * k + c) * ld + 8 * k + r];
  }
  __syncthreads();
}

