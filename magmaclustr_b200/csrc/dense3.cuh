// dense3.cuh -- device building blocks of the shared-memory-resident blocked Cholesky / triangular inverse /
// X X' on packed swizzled 8x8 tiles with a run-time tile count (dense3.cu: whole union-grid matrices up to
// 224 x 224; dense_big.cu: the 64 x 64 diagonal blocks of the multi-CTA path).
#pragma once
#include "linalg.cuh"
#include "tile3.cuh"

#define NWS 16                  /* warps of the shared-memory-resident kernel */
#define DSM_THREADS (32 * NWS)
#define DSM_MAX_N8 28

__host__ __device__ inline int d3_tri(int i, int n8) { return i * n8 - ((i * (i - 1)) >> 1); }

struct SmD {
  double *T1, *Xd, *rinv, *red;
  unsigned short* tab;
  int* flag;
};

__device__ inline SmD carve_d(double* base, int n8) {
  const int ntri = n8 * (n8 + 1) / 2;
  SmD s;
  s.T1 = base;
  s.Xd = s.T1 + (size_t)ntri * 64;
  s.rinv = s.Xd + (size_t)n8 * 64;
  s.red = s.rinv + (size_t)n8 * 8;
  s.tab = (unsigned short*)(s.red + 2 * NWS);
  s.flag = (int*)(s.tab + (((size_t)ntri + 3) & ~(size_t)3));
  return s;
}

__device__ __forceinline__ void d3_update_tile(double* T1, int idx, int e, int k, int dk, const Ln& L) {
  const int i = e >> 8, j = e & 255;
  double* Tij = T1 + (size_t)idx * 64;
  double2 cv = *reinterpret_cast<double2*>(Tij + L.o_c);
  double a0, a1, b0, b1;
  cf(T1 + (size_t)(dk + i - k) * 64, L, a0, a1);
  cf(T1 + (size_t)(dk + j - k) * 64, L, b0, b1);
  dmma884(cv.x, cv.y, -a0, b0);
  dmma884(cv.x, cv.y, -a1, b1);
  *reinterpret_cast<double2*>(Tij + L.o_c) = cv;
}

// blocked upper Cholesky in place; every panel tile is a substitution solve with the factored diagonal
// tile (no explicit inverse anywhere in the factorisation).  0 or failing pivot code.
template <int NW>
__device__ inline int d3_chol(const SmD& s, int n8, const Ln& L, int rot = 0) {
  // `rot` rotates the roles of the warps: the panel warp (virtual warp 0) of co-resident CTAs then sits on
  // different SM sub-partitions instead of all serial pivot chains sharing sub-partition 0
  static_assert((NW & (NW - 1)) == 0, "NW must be a power of two");
  const int vw = (L.warp + rot) & (NW - 1);
  const int ntri = n8 * (n8 + 1) / 2;
  if (threadIdx.x == 0) *s.flag = 0;
  __syncthreads();
  if (vw == 0) {
    const int f = warp_factor_panel(s.T1, n8 < 4 ? n8 : 4, s.rinv, L.lane);
    if (f && L.lane == 0) *s.flag = f;
  }
  __syncthreads();
  if (*s.flag) return *s.flag;
  for (int k = 0; k + 1 < n8; ++k) {
    const int dk = d3_tri(k, n8);
    if (k + 4 < n8) {
      // remaining panel tiles R_kj = R_kk^{-T} A_kj, four tiles per warp
      for (int j = k + 4 + 4 * vw; j < n8; j += 4 * NW)
        warp_panel_solve(s.T1 + (size_t)dk * 64, s.rinv + 8 * k, s.T1 + (size_t)(dk + j - k) * 64,
                         (n8 - j) < 4 ? (n8 - j) : 4, L.lane);
      __syncthreads();
    }
    const int m = n8 - k - 1, nla = m < 4 ? m : 4;
    const int j0 = d3_tri(k + 1, n8);
    if (vw == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < nla) d3_update_tile(s.T1, j0 + q, s.tab[j0 + q], k, dk, L);
      __syncwarp();
      const int f = warp_factor_panel(s.T1 + (size_t)j0 * 64, nla, s.rinv + 8 * (k + 1), L.lane);
      if (f && L.lane == 0) *s.flag = 8 * (k + 1) + f;
    } else {
      for (int idx = j0 + nla + vw - 1; idx < ntri; idx += NW - 1) d3_update_tile(s.T1, idx, s.tab[idx], k, dk, L);
    }
    __syncthreads();
    if (*s.flag) return *s.flag;
  }
  return 0;
}

// in place X = R^{-1}: diagonal tiles by substitution (Xd), then block rows bottom-up:
//   W_ij = -sum_{i<m<=j} R_im X_mj  (DMMA, held in registers until every read of row i is done),
//   X_ij = R_ii^{-1} W_ij           (back substitution with the diagonal tile, like dtrsm);
// finally the diagonal tiles X_ii are copied in.
template <int NW, int HT>
__device__ inline void d3_trtri(const SmD& s, int n8, const Ln& L) {
  for (int k = L.warp; k < n8; k += NW)
    warp_diag_inverse(s.T1 + (size_t)d3_tri(k, n8) * 64, s.rinv + 8 * k, s.Xd + (size_t)k * 64, L.lane);
  __syncthreads();
  for (int i = n8 - 2; i >= 0; --i) {
    double* Ti = s.T1 + (size_t)(d3_tri(i, n8) - i) * 64;   // tile (i, m) at Ti + m * 64
    double res[HT][2];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const int j = i + 1 + L.warp + h * NW;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      if (j < n8) {
        int m = i + 1;
        for (; m + 1 <= j; m += 2) {
          double a0, a1, b0, b1, p0, p1, q0, q1;
          rf(Ti + (size_t)m * 64, L, a0, a1);
          rf(Ti + (size_t)(m + 1) * 64, L, p0, p1);
          cf(s.T1 + (size_t)(d3_tri(m, n8) + j - m) * 64, L, b0, b1);
          if (m + 1 < j) cf(s.T1 + (size_t)(d3_tri(m + 1, n8) + j - m - 1) * 64, L, q0, q1);
          else cf(s.Xd + (size_t)j * 64, L, q0, q1);
          dmma884(c0, c1, -a0, b0);
          dmma884(e0, e1, -p0, q0);
          dmma884(c0, c1, -a1, b1);
          dmma884(e0, e1, -p1, q1);
        }
        if (m <= j) {   // m == j
          double a0, a1, b0, b1;
          rf(Ti + (size_t)m * 64, L, a0, a1);
          cf(s.Xd + (size_t)j * 64, L, b0, b1);
          dmma884(c0, c1, -a0, b0);
          dmma884(c0, c1, -a1, b1);
        }
      }
      res[h][0] = c0 + e0;
      res[h][1] = c1 + e1;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const int j = i + 1 + L.warp + h * NW;
      if (j < n8) *reinterpret_cast<double2*>(Ti + (size_t)j * 64 + L.o_c) = make_double2(res[h][0], res[h][1]);
    }
    __syncthreads();
    for (int j = i + 1 + 4 * L.warp; j < n8; j += 4 * NW)
      warp_row_backsolve(Ti + (size_t)i * 64, s.rinv + 8 * i, Ti + (size_t)j * 64, (n8 - j) < 4 ? (n8 - j) : 4, L.lane);
    __syncthreads();
  }
  for (int e = threadIdx.x; e < n8 * 64; e += (32 * NW)) s.T1[(size_t)d3_tri(e >> 6, n8) * 64 + (e & 63)] = s.Xd[e];
  __syncthreads();
}

// in place A = X X' (upper tiles), rows top-down; diagonal tiles mirrored afterwards
template <int NW, int HT>
__device__ inline void d3_lauum(const SmD& s, int n8, const Ln& L) {
  for (int i = 0; i < n8; ++i) {
    const double* Ti = s.T1 + (size_t)(d3_tri(i, n8) - i) * 64;
    double res[HT][2];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const int j = i + L.warp + h * NW;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      if (j < n8) {
        const double* Tj = s.T1 + (size_t)(d3_tri(j, n8) - j) * 64;
        int m = j;
        for (; m + 1 < n8; m += 2) {
          double a0, a1, b0, b1, p0, p1, q0, q1;
          rf(Ti + (size_t)m * 64, L, a0, a1);
          rf(Tj + (size_t)m * 64, L, b0, b1);
          rf(Ti + (size_t)(m + 1) * 64, L, p0, p1);
          rf(Tj + (size_t)(m + 1) * 64, L, q0, q1);
          dmma884(c0, c1, a0, b0);
          dmma884(e0, e1, p0, q0);
          dmma884(c0, c1, a1, b1);
          dmma884(e0, e1, p1, q1);
        }
        if (m < n8) {
          double a0, a1, b0, b1;
          rf(Ti + (size_t)m * 64, L, a0, a1);
          rf(Tj + (size_t)m * 64, L, b0, b1);
          dmma884(c0, c1, a0, b0);
          dmma884(c0, c1, a1, b1);
        }
      }
      res[h][0] = c0 + e0;
      res[h][1] = c1 + e1;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const int j = i + L.warp + h * NW;
      if (j < n8) *reinterpret_cast<double2*>(const_cast<double*>(Ti) + (size_t)j * 64 + L.o_c) = make_double2(res[h][0], res[h][1]);
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < n8 * 64; e += (32 * NW)) {
    const int k = e >> 6, r = (e >> 3) & 7, c = e & 7;
    double* Tk = s.T1 + (size_t)d3_tri(k, n8) * 64;
    if (r > c) Tk[sw(r, c)] = Tk[sw(c, r)];
  }
  __syncthreads();
}

