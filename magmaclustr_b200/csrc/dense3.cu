// dense3.cu -- the union-grid (mean / cluster process) problem, one matrix of U x U, version 3:
//   * k_dense_cholinv_sm: chol_inv_jitter of one matrix held ENTIRELY in the shared memory of one CTA as
//     packed, swizzled 8x8 tiles (U <= 224: 406 tiles = 203 KB), 16 warps, every tile product a DMMA;
//     in-place blocked Cholesky (warp 0 one step ahead), in-place triangular inverse (rows bottom-up)
//     and in-place X X' (rows top-down).  Replaces the L2-resident single-CTA kernel of dense.cu, whose
//     every dependent tile access paid an L2 round trip.
//   * k_gemm_dmma: C = alpha op(A) op(B) + beta C on the FP64 tensor cores, 32x32 output blocks.
//   * k_dense_obj_final3: parallel, fixed-order final reduction of logL_GP_mod / gr_GP_mod.
//
// Reference path: kern_to_inv(all_inputs, kern_0, hp_0) and chol_inv_jitter(post_inv) of e_step
// (/root/reference/R/em-magma.R:33,48-57), the hp_0 objective of m_step (R/em-magma.R:157-166),
// chol(post_cov) of logL_monitoring (R/likelihoods.R:303-307).
#include "dense3.cuh"

#include <stdlib.h>

#include <utility>

#define DSM_SPLIT_N 128   /* smallest matrix for which the X X' product leaves the single CTA */

size_t dense_sm_bytes(int n) {
  const int n8 = (n + 7) / 8, ntri = n8 * (n8 + 1) / 2;
  return ((size_t)ntri * 64 + (size_t)n8 * 64 + (size_t)n8 * 8 + 2 * NWS) * sizeof(double) +
         (((size_t)ntri + 3) & ~(size_t)3) * sizeof(unsigned short) + 16;
}
bool dense_sm_fits(int n) { return (n + 7) / 8 <= DSM_MAX_N8; }

extern __shared__ double4 dense3_smem_raw[];

// mode 0: chol_inv_jitter(src, pen) -> inv (n x n, full, exactly symmetric); info[0] = retries (-1: gave
//         up), out[0] = total jitter, out[1] = sum log diag(R)               kernel-to-matrices.R:670-680
// mode 1: sum log diag chol(src), NO jitter; info[0] = 0 or failing pivot    likelihoods.R:303-307
__global__ void __launch_bounds__(DSM_THREADS, 1) k_dense_cholinv_sm(const double* __restrict__ src, int ld_src,
                                                                     double* __restrict__ inv, int n, double pen,
                                                                     int max_retry, int mode, int* info, double* out,
                                                                     double* __restrict__ xt_out) {
  const int n8 = (n + 7) >> 3, ntri = n8 * (n8 + 1) / 2;
  const SmD s = carve_d((double*)dense3_smem_raw, n8);
  const Ln L;
  for (int i = threadIdx.x; i < n8; i += DSM_THREADS)
    for (int j = i; j < n8; ++j) s.tab[d3_tri(i, n8) + j - i] = (unsigned short)((i << 8) | j);
  __syncthreads();
  int retries = 0;
  double total = 0.0;
  long long tk0 = clock64(), tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;   // phase marks (thread 0 reports them in out[4..9])
  for (;;) {
    total = 0.0;
    if (mode == 0) { double p = pen; for (int r = 0; r <= retries; ++r) { total += p; p = 10 * p; } }
    // upper tiles only, one 8 x 8 tile per 64 consecutive threads (rows of 8 doubles = two 32-byte sectors), eight
    // independent loads in flight per thread: the matrix comes from L2 and a dependent load costs ~700 cycles
    for (int e0 = threadIdx.x; e0 < ntri * 64; e0 += 8 * DSM_THREADS) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + q * DSM_THREADS;
        v[q] = 0.0;
        if (e < ntri * 64) {
          const int t = s.tab[e >> 6], i = 8 * (t >> 8) + ((e >> 3) & 7), j = 8 * (t & 255) + (e & 7);
          if (i < n && j < n) { if (j >= i) v[q] = src[(size_t)i * ld_src + j]; }
          else if (i == j) v[q] = 1.0;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + q * DSM_THREADS;
        if (e < ntri * 64) {
          const int t = s.tab[e >> 6], i = 8 * (t >> 8) + ((e >> 3) & 7), j = 8 * (t & 255) + (e & 7);
          double x = v[q];
          if (i == j && i < n && mode == 0) { double p = pen; for (int r = 0; r <= retries; ++r) { x += p; p = 10 * p; } }
          s.T1[(size_t)(e >> 6) * 64 + sw(i & 7, j & 7)] = x;
        }
      }
    }
    __syncthreads();
    tk1 = clock64();
    const int f = d3_chol<NWS>(s, n8, L);
    tk2 = clock64();
    if (f == 0) break;
    __syncthreads();
    if (mode == 1) { if (threadIdx.x == 0) { info[0] = f; out[0] = NAN; } return; }
    if (++retries > max_retry) {
      if (threadIdx.x == 0) { info[0] = -1; out[0] = total; out[1] = NAN; }
      return;
    }
  }
  {
    double v = 0.0;
    for (int j = threadIdx.x; j < n; j += DSM_THREADS) v += log(s.T1[(size_t)d3_tri(j >> 3, n8) * 64 + sw(j & 7, j & 7)]);
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULLMASK, v, off);
    if (L.lane == 0) s.red[L.warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < NWS; ++w) t += s.red[w];
      if (mode == 1) { info[0] = 0; out[0] = t; }
      else { info[0] = retries; out[0] = total; out[1] = t; }
    }
  }
  if (mode == 1) return;
  tk3 = clock64();
  d3_trtri<NWS, 2>(s, n8, L);
  tk4 = clock64();
  if (xt_out) {
    // split path: X = R^-1 (packed upper tiles) goes to global memory and k_mid_lauum forms A = X X' on many SMs -- that
    // product is ~9 k DMMAs, 36 k cycles of ONE SM's tensor pipe here, a few microseconds when it is spread out
    for (int e = threadIdx.x; e < ntri * 64; e += DSM_THREADS) xt_out[e] = s.T1[e];
    if (threadIdx.x == 0) {
      out[4] = (double)(tk1 - tk0); out[5] = (double)(tk2 - tk1); out[6] = (double)(tk3 - tk2); out[7] = (double)(tk4 - tk3);
      out[8] = 0.0; out[9] = (double)(clock64() - tk4);
    }
    return;
  }
  d3_lauum<NWS, 2>(s, n8, L);
  tk5 = clock64();
  // full symmetric matrix out: a warp writes 32 consecutive doubles of one row, rows are dealt to the warps
  for (int r = L.warp; r < n; r += NWS) {
    const int i = r >> 3;
    for (int c = L.lane; c < n; c += 32) {
      const int j = c >> 3;
      inv[(size_t)r * n + c] = (i <= j) ? s.T1[(size_t)(d3_tri(i, n8) + j - i) * 64 + sw(r & 7, c & 7)]
                                        : s.T1[(size_t)(d3_tri(j, n8) + i - j) * 64 + sw(c & 7, r & 7)];
    }
  }
  (void)ntri;
  if (threadIdx.x == 0) {
    out[4] = (double)(tk1 - tk0); out[5] = (double)(tk2 - tk1); out[6] = (double)(tk3 - tk2); out[7] = (double)(tk4 - tk3);
    out[8] = (double)(tk5 - tk4); out[9] = (double)(clock64() - tk5);
  }
}

// A = X X' from the packed upper tiles of X = R^-1 left in global memory by k_dense_cholinv_sm (split path): one warp per
// upper tile (i, j), the same two accumulator chains over even / odd m and the same final c + e as d3_lauum, so every
// entry is bit-identical to the single-CTA kernel's; the operand fragments of four m-steps are fetched from L2 before the
// first product.  Writes the full symmetric n x n matrix (row-major), lower part mirrored from the upper one.
#define ML_WARPS 4
__global__ void __launch_bounds__(32 * ML_WARPS) k_mid_lauum(const double* __restrict__ xt, int n, double* __restrict__ inv,
                                                             const int* __restrict__ info) {
  if (info[0] < 0) return;   // the factorisation gave up: nothing to form
  const int n8 = (n + 7) >> 3, ntri = n8 * (n8 + 1) / 2;
  const Ln L;
  const int idx = blockIdx.x * ML_WARPS + L.warp;
  if (idx >= ntri) return;
  int i = 0, rem = idx;
  while (rem >= n8 - i) { rem -= n8 - i; ++i; }
  const int j = i + rem;
  const double* Ti = xt + (size_t)(d3_tri(i, n8) - i) * 64;   // tile (i, m) at Ti + m * 64
  const double* Tj = xt + (size_t)(d3_tri(j, n8) - j) * 64;
  double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
  int m = j;
  for (; m + 3 < n8; m += 4) {
    double a0, a1, b0, b1, p0, p1, q0, q1, r0, r1, s0, s1, t0, t1, u0, u1;
    rf(Ti + (size_t)m * 64, L, a0, a1);
    rf(Tj + (size_t)m * 64, L, b0, b1);
    rf(Ti + (size_t)(m + 1) * 64, L, p0, p1);
    rf(Tj + (size_t)(m + 1) * 64, L, q0, q1);
    rf(Ti + (size_t)(m + 2) * 64, L, r0, r1);
    rf(Tj + (size_t)(m + 2) * 64, L, s0, s1);
    rf(Ti + (size_t)(m + 3) * 64, L, t0, t1);
    rf(Tj + (size_t)(m + 3) * 64, L, u0, u1);
    dmma884(c0, c1, a0, b0);
    dmma884(e0, e1, p0, q0);
    dmma884(c0, c1, a1, b1);
    dmma884(e0, e1, p1, q1);
    dmma884(c0, c1, r0, s0);
    dmma884(e0, e1, t0, u0);
    dmma884(c0, c1, r1, s1);
    dmma884(e0, e1, t1, u1);
  }
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
  const double v[2] = {c0 + e0, c1 + e1};
  const int r = 8 * i + L.g;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c = 8 * j + 2 * L.t + q;
    if (r < n && c < n && (i != j || c >= r)) {
      inv[(size_t)r * n + c] = v[q];
      if (c != r) inv[(size_t)c * n + r] = v[q];
    }
  }
}

// ---- C = alpha A B + beta Cin on the FP64 tensor cores ------------------------------------------------
// A(i,k) = A[i*ars + k*acs], B(k,j) = B[k*brs + j*bcs], C row-major (ldc).  CTA: 4 warps, 32 x 32 block,
// K slabs of 32 staged in shared memory (leading dimension 36: fragment loads conflict-free); the next slab is fetched
// into registers while the current one is multiplied -- at the 201-point union grid the kernel is a chain of L2 round
// trips (one per slab), so the slab is as deep as the register budget allows.
#define GD_BM 32
#define GD_BK 32
__global__ void __launch_bounds__(128) k_gemm_dmma(int m, int n, int kk, const double* __restrict__ A, int64_t ars,
                                                   int64_t acs, const double* __restrict__ B, int64_t brs, int64_t bcs,
                                                   double* C, int64_t ldc, double alpha, const double* Cin, double beta) {
  __shared__ double As[GD_BM][GD_BK + 4];
  __shared__ double Bs[GD_BK][GD_BM + 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wy = warp >> 1, wx = warp & 1;
  const int i0 = blockIdx.y * GD_BM, j0 = blockIdx.x * GD_BM;
  double acc[2][2][2] = {};
  double pa[8], pb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 128 * q;
      const int ar = e >> 5, ac = e & 31;
      const int gi = i0 + ar, gk = k0 + ac;
      pa[q] = (gi < m && gk < kk) ? A[gi * ars + gk * acs] : 0.0;
      const int br = e >> 5, bc = e & 31;
      const int gk2 = k0 + br, gj = j0 + bc;
      pb[q] = (gk2 < kk && gj < n) ? B[gk2 * brs + gj * bcs] : 0.0;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < kk; k0 += GD_BK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 128 * q;
      As[e >> 5][e & 31] = pa[q];
      Bs[e >> 5][e & 31] = pb[q];
    }
    __syncthreads();
    if (k0 + GD_BK < kk) fetch(k0 + GD_BK);
#pragma unroll
    for (int ks = 0; ks < GD_BK; ks += 4) {
      double a[2], b[2];
#pragma unroll
      for (int ti = 0; ti < 2; ++ti) a[ti] = As[16 * wy + 8 * ti + g][ks + t];
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) b[tj] = Bs[ks + t][16 * wx + 8 * tj + g];
#pragma unroll
      for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) dmma884(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int ti = 0; ti < 2; ++ti)
#pragma unroll
    for (int tj = 0; tj < 2; ++tj)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int gi = i0 + 16 * wy + 8 * ti + g, gj = j0 + 16 * wx + 8 * tj + 2 * t + q;
        if (gi < m && gj < n) {
          double v = alpha * acc[ti][tj][q];
          if (Cin) v += beta * Cin[gi * ldc + gj];
          C[gi * ldc + gj] = v;
        }
      }
}

// Final scalars of logL_GP_mod / gr_GP_mod: f = 0.5 (n log 2pi - logdetA + r'al) + 0.5 tr ;
// g_p = -0.5 sum_b part[b][1+p].  One CTA, fixed-order tree (deterministic).
__global__ void __launch_bounds__(256) k_dense_obj_final3(int n, int P, int nparts, const double* part, const double* r,
                                                          const double* al, const double* logdet, int from_chol,
                                                          int add_trace, double* res) {
  __shared__ double sh[256];
  const int tid = threadIdx.x;
  for (int q = 0; q < 2 + P; ++q) {
    double v = 0.0;
    if (q == 0) { for (int i = tid; i < n; i += 256) v += r[i] * al[i]; }
    else { for (int b = tid; b < nparts; b += 256) v += part[(size_t)b * (1 + MB_MAX_HP) + (q - 1)]; }
    sh[tid] = v;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) sh[tid] += sh[tid + off];
      __syncthreads();
    }
    if (tid == 0) {
      if (q == 0) res[16 + 8] = sh[0];           // quad (scratch slot)
      else if (q == 1) res[16 + 9] = sh[0];      // trace
      else res[1 + (q - 2)] = -0.5 * sh[0];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double ldA = from_chol ? -2.0 * logdet[0] : logdet[0];
    double f = 0.5 * (n * 1.8378770664093453 - ldA + res[16 + 8]);
    if (add_trace) f = f + 0.5 * res[16 + 9];
    res[0] = f;
  }
}

// r = y - mean ; al = A r   (one warp per row, fixed summation tree)
__global__ void k_resid_gemv(int n, const double* A, const double* y, const double* mean, double* r, double* al) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s += A[(size_t)row * n + k] * (y[k] - mean[k]);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) { al[row] = s; r[row] = y[row] - mean[row]; }
}

// ---- launchers -----------------------------------------------------------------------------------------
// xt: caller-owned scratch of at least npad^2 doubles (npad = n rounded up to 8), private to the stream `st` (the
// contexts pass the workspace of that stream's DenseWs); NULL keeps the whole inverse in the one CTA.
// The kernel's shared-memory opt-in is per device: a process drives ONE device (mb_create enforces it).
cudaError_t launch_dense_cholinv_sm(const double* src, int ld_src, double* inv, int n, double pen, int mode,
                                    int* info, double* out, double* xt, cudaStream_t st) {
  const size_t smem = dense_sm_bytes(n);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_dense_cholinv_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  // mode 0 at n >= DSM_SPLIT_N: the kernel stops after X = R^-1 and a multi-CTA kernel forms X X' (k_mid_lauum) from the
  // packed X in xt
  if (!(mode == 0 && n >= DSM_SPLIT_N)) xt = nullptr;
  k_dense_cholinv_sm<<<1, DSM_THREADS, smem, st>>>(src, ld_src, inv, n, pen, 40, mode, info, out, xt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !xt) return e;
  const int n8 = (n + 7) / 8, ntri = n8 * (n8 + 1) / 2;
  k_mid_lauum<<<(ntri + ML_WARPS - 1) / ML_WARPS, 32 * ML_WARPS, 0, st>>>(xt, n, inv, info);
  return cudaGetLastError();
}

cudaError_t launch_gemm_dmma(int m, int n, int k, const double* A, int64_t ars, int64_t acs, const double* B,
                             int64_t brs, int64_t bcs, double* C, int64_t ldc, double alpha, const double* Cin,
                             double beta, cudaStream_t st) {
  dim3 grid((n + GD_BM - 1) / GD_BM, (m + GD_BM - 1) / GD_BM);
  k_gemm_dmma<<<grid, 128, 0, st>>>(m, n, k, A, ars, acs, B, brs, bcs, C, ldc, alpha, Cin, beta);
  return cudaGetLastError();
}

cudaError_t launch_dense_obj_final3(int n, int P, int nparts, const double* part, const double* r, const double* al,
                                    const double* logdet, int from_chol, int add_trace, double* res, cudaStream_t st) {
  k_dense_obj_final3<<<1, 256, 0, st>>>(n, P, nparts, part, r, al, logdet, from_chol, add_trace, res);
  return cudaGetLastError();
}

cudaError_t launch_resid_gemv(int n, const double* A, const double* y, const double* mean, double* r, double* al,
                              cudaStream_t st) {
  k_resid_gemv<<<(n + 7) / 8, 256, 0, st>>>(n, A, y, mean, r, al);
  return cudaGetLastError();
}
