// dense.cu -- kernels for ONE matrix-sized problem on the union grid (mean process, cluster
// processes, prediction): covariance build, chol_inv_jitter, log-determinants, products and
// the logL_GP_mod / gr_GP_mod reductions.  Matrices are row-major n x n in global memory
// (L2-resident at U = 201: 323 KB); symmetric ones are stored full.
//
// Reference path replaced: kern_to_inv(all_inputs, kern_0, hp_0) and chol_inv_jitter(post_inv)
// in e_step (/root/reference/R/em-magma.R:33,48-57), logL_GP_mod/gr_GP_mod on the mean
// process (R/em-magma.R:157-166; R/likelihoods.R:155-174; R/gradients-likelihoods.R:71-97),
// chol(post_cov) of logL_monitoring (R/likelihoods.R:303-307) and the matrix algebra of
// pred_magma (R/prediction.R:1146-1189).
#include "linalg.cuh"
#include "tiles.cuh"

#define NWD 16 /* warps of the single-CTA union-grid kernels */

struct HpVal { double v[MB_MAX_HP]; };

// out[i*rs + j*cs] = cov (deriv < 0) or d cov / d hp[deriv], x1: n1 x D, x2: n2 x D row-major.
__global__ void k_dense_cov(DevKern kd, HpVal hp, const double* x1, int n1, const double* x2,
                            int n2, int D, int deriv, int with_noise, double* out, int64_t rs,
                            int64_t cs) {
  const int64_t tot = (int64_t)n1 * n2;
  double hpe[MB_MAX_HP];
  prep_hp(kd, hp.v, hpe);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot;
       e += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(e / n2), j = (int)(e % n2);
    double dv[MB_MAX_HP];
    double v;
    if (deriv < 0) v = cov_eval<false>(kd, hp.v, hpe, x1 + (size_t)i * D, x2 + (size_t)j * D, D,
                                       with_noise != 0, nullptr);
    else { cov_eval<true>(kd, hp.v, hpe, x1 + (size_t)i * D, x2 + (size_t)j * D, D, with_noise != 0, dv);
           v = dv[deriv]; }
    out[i * rs + j * cs] = v;
  }
}

// The five native builders of /root/reference/src/code.cpp; inputs and output column-major.
__global__ void k_legacy_builder(int which, const double* m1, int n1, const double* m2, int n2,
                                 int d, double par, double* out) {
  const int64_t tot = (int64_t)n1 * n2;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot;
       e += (int64_t)gridDim.x * blockDim.x) {
    int r1 = (int)(e % n1), r2 = (int)(e / n1);
    double total = 0.0;
    for (int c = 0; c < d; ++c) {
      double a = m1[r1 + (size_t)c * n1], b = m2[r2 + (size_t)c * n2];
      if (which == 0) { double df = a - b; total += df * df; }                 // cpp_dist
      else if (which == 1) { double sn = sin(MB_PI / exp(par) * fabs(a - b)); total += sn * sn; }
      else if (which == 2) { double ang = MB_PI / exp(par) * fabs(a - b);      // cpp_perio_deriv
                             total += 2 * sin(ang) * cos(ang) * ang; }
      else if (which == 3) total += a * b;                                     // cpp_prod
      else total += fabs(a - b);                                               // cpp_noise
    }
    if (which == 4) total = (total < 0.000001) ? exp(par) : 0.0;
    out[e] = total;
  }
}

// log|det src| through LU with partial pivoting (determinant(), likelihoods.R:37-41).
__global__ void __launch_bounds__(1024) k_dense_lu_logdet(const double* src, double* W, int n,
                                                          int* done, double* out) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) W[e] = src[e];
  __syncthreads();
  double ldet = blk_lu_logabsdet(W, n, n, done);
  if (threadIdx.x == 0) out[0] = ldet;
}

// y[i] = sum_k A[i*ld + k] x[k] (+ y0[i]); one warp per row, fixed summation tree.
__global__ void k_gemv(int m, int n, const double* A, int64_t ld, const double* x,
                       const double* y0, double* y) {
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= m) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s += A[row * ld + k] * x[k];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) y[row] = y0 ? y0[row] + s : s;
}

// Partial sums for logL_GP_mod / gr_GP_mod of one n x n problem:
//   part[b][0]      = sum A o S                       (likelihoods.R:171)
//   part[b][1 + p]  = sum W o dK_p,  W = al al' + T - A   (T = A S A; T NULL for the marginal form)
// over the elements handled by CTA b (both triangles, like the reference's full trace).
__global__ void __launch_bounds__(256) k_dense_obj_partials(
    DevKern kd, HpVal hp, const double* x, int n, int D, const double* A, const double* S,
    const double* T, const double* al, int want_grad, double* part) {
  __shared__ double red[4 * 32];
  const int P = kd.n_hp;
  double tr = 0.0;
  double gacc[MB_MAX_HP], hpe[MB_MAX_HP];
  prep_hp(kd, hp.v, hpe);
#pragma unroll
  for (int p = 0; p < MB_MAX_HP; ++p) gacc[p] = 0.0;
  const int64_t tot = (int64_t)n * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot;
       e += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(e / n), j = (int)(e % n);
    double a = A[e];
    if (S) tr += a * S[e];
    if (want_grad) {
      double w = al[i] * al[j] - a;
      if (T) w += T[e];
      double dv[MB_MAX_HP];
      cov_eval<true>(kd, hp.v, hpe, x + (size_t)i * D, x + (size_t)j * D, D, true, dv);
      for (int p = 0; p < P; ++p) gacc[p] += w * dv[p];
    }
  }
  double* out = part + (size_t)blockIdx.x * (1 + MB_MAX_HP);
  {
    double v[4] = {tr, 0, 0, 0};
    blk_reduce_sum<4>(v, red);
    if (threadIdx.x == 0) out[0] = v[0];
  }
  for (int p0 = 0; p0 < P; p0 += 4) {
    double v[4];
    for (int q = 0; q < 4; ++q) v[q] = (p0 + q < P) ? gacc[p0 + q] : 0.0;
    blk_reduce_sum<4>(v, red);
    if (threadIdx.x == 0)
      for (int q = 0; q < 4 && p0 + q < P; ++q) out[1 + p0 + q] = v[q];
  }
}

// Final scalars: f = 0.5 (n log 2pi - logdetA + r'al) + 0.5 tr ;  g_p = -0.5 sum_b part[b][1+p].
// res[0] = f, res[1..P] = g.  logdet[0] = log|det A| (LU) or, when from_chol, sum log diag R.
__global__ void k_dense_obj_final(int n, int P, int nparts, const double* part, const double* r,
                                  const double* al, const double* logdet, int from_chol,
                                  int add_trace, double* res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double quad = 0.0;
  for (int i = 0; i < n; ++i) quad += r[i] * al[i];
  double tr = 0.0;
  for (int b = 0; b < nparts; ++b) tr += part[(size_t)b * (1 + MB_MAX_HP)];
  double ldA = from_chol ? -2.0 * logdet[0] : logdet[0];
  double f = 0.5 * (n * 1.8378770664093453 - ldA + quad);
  if (add_trace) f = f + 0.5 * tr;
  res[0] = f;
  for (int p = 0; p < P; ++p) {
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += part[(size_t)b * (1 + MB_MAX_HP) + 1 + p];
    res[1 + p] = -0.5 * s;
  }
}

// Sum of n partial vectors (part + g * part_stride, g < n) in a FIXED BINARY TREE over g: node (level, i) = node(level - 1,
// 2 i) + node(level - 1, 2 i + 1); a missing right child passes the left one through.  For n a power of two the sum over an
// aligned block of 2^k partials is a subtree, so reducing blocks separately (one per GPU) and then the block sums gives the
// same bits as one reduction over all partials: this is what makes the E step independent of the number of GPUs.
// One thread per element, unit-stride loads across the warp; out = base + tree (base may be NULL).
template <int LOG>
__device__ __forceinline__ bool tree_sum(const double* __restrict__ p, int64_t stride, int g0, int n, double& out) {
  if (g0 >= n) return false;
  if (LOG == 0) { out = __ldcg(p + (size_t)g0 * stride); return true; }
  double l, r;
  tree_sum<(LOG > 0 ? LOG - 1 : 0)>(p, stride, g0, n, l);
  const bool hr = tree_sum<(LOG > 0 ? LOG - 1 : 0)>(p, stride, g0 + (1 << (LOG > 0 ? LOG - 1 : 0)), n, r);
  out = hr ? l + r : l;
  return true;
}
#define RT_MAX_LOG 8   /* up to 256 partials */
__global__ void __launch_bounds__(256) k_reduce_tree(int64_t count, int n, const double* __restrict__ base,
                                                     const double* __restrict__ part, int64_t part_stride,
                                                     double* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < count; e += (int64_t)gridDim.x * 256) {
    double v = 0.0;
    if (n <= 2) tree_sum<1>(part + e, part_stride, 0, n, v);
    else if (n <= 8) tree_sum<3>(part + e, part_stride, 0, n, v);
    else if (n <= 32) tree_sum<5>(part + e, part_stride, 0, n, v);
    else tree_sum<RT_MAX_LOG>(part + e, part_stride, 0, n, v);
    out[e] = base ? base[e] + v : v;
  }
}

// The same tree with eight threads per element: warp w of a 256-thread block sums the aligned subtree of 2^SLOG partials
// number w (a node of the tree above) for 32 consecutive elements -- unit-stride 256-byte loads -- and the first warp adds the
// eight nodes in tree order.  Same bits as k_reduce_tree; eight times the loads in flight (the leaf accumulators sit in L2,
// one thread per element left the kernel latency bound: 38 -> 8.7 us for 128 x 323 KB = 4.8 TB/s).
template <int SLOG>
__global__ void __launch_bounds__(256) k_reduce_tree8(int64_t count, int n, const double* __restrict__ base,
                                                      const double* __restrict__ part, int64_t part_stride,
                                                      double* __restrict__ out) {
  __shared__ double sv[8][32];
  __shared__ int sh[8][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  double v = 0.0;
  bool has = false;
  if (e < count) has = tree_sum<SLOG>(part + e, part_stride, w << SLOG, n, v);
  sv[w][lane] = v;
  sh[w][lane] = has ? 1 : 0;
  __syncthreads();
  if (w == 0 && e < count) {
    double a[8];
    bool h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { a[q] = sv[q][lane]; h[q] = sh[q][lane] != 0; }
#pragma unroll
    for (int step = 1; step < 8; step <<= 1)
#pragma unroll
      for (int q = 0; q < 8; q += 2 * step)
        if (h[q + step]) { a[q] = a[q] + a[q + step]; }   // a missing right child passes the left one through
    const double t = h[0] ? a[0] : 0.0;
    out[e] = base ? base[e] + t : t;
  }
}

// lower triangle := mirror image of the upper one, for `nmat` row-major U x U matrices `stride` doubles apart
__global__ void __launch_bounds__(256) k_mirror_upper(double* m, int U, int nmat, int64_t stride) {
  const int64_t tot = (int64_t)nmat * U * U;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < tot; e += (int64_t)gridDim.x * 256) {
    const int k = (int)(e / ((int64_t)U * U));
    const int64_t r = e - (int64_t)k * U * U;
    const int i = (int)(r / U), j = (int)(r - (int64_t)i * U);
    if (i > j) m[(int64_t)k * stride + r] = m[(int64_t)k * stride + (int64_t)j * U + i];
  }
}
cudaError_t launch_mirror_upper(double* m, int U, int nmat, int64_t stride, cudaStream_t st) {
  int64_t blocks = ((int64_t)nmat * U * U + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_mirror_upper<<<(unsigned)blocks, 256, 0, st>>>(m, U, nmat, stride);
  return cudaGetLastError();
}

// out[b] = sum_i v[b + i*stride] for block b: the reference's sapply(...) %>% sum() / rowSums().
// R accumulates in long double; here every thread keeps a Neumaier-compensated partial over a strided
// subset and thread 0 combines the 256 (sum, compensation) pairs in index order, which gives the
// correctly rounded sum up to one ulp, independent of the launch geometry (deterministic).
__device__ __forceinline__ void neumaier_add(double& s, double& c, double x) {
  const double t = s + x;
  c += (fabs(s) >= fabs(x)) ? (s - t) + x : (x - t) + s;
  s = t;
}
__global__ void __launch_bounds__(256) k_sum_compensated(const double* v, int64_t n, int64_t stride, double* out) {
  __shared__ double sh[256], sc[256];
  v += blockIdx.x;
  double s = 0.0, c = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) neumaier_add(s, c, v[i * stride]);
  sh[threadIdx.x] = s;
  sc[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, C = 0.0, plain = 0.0;
    for (int w = 0; w < 256; ++w) { neumaier_add(S, C, sh[w]); C += sc[w]; plain += sh[w]; }
    out[blockIdx.x] = isfinite(plain) ? S + C : plain;   // Inf / NaN propagate like a plain sum
  }
}

// per-leaf version: block (p, l) sums column p of the row-major table v (P columns) over the tasks of leaf l
__global__ void __launch_bounds__(256) k_leaf_sum_compensated(const double* v, int P, const int64_t* leaf_offs, double* out) {
  __shared__ double sh[256], sc[256];
  const int p = blockIdx.x, l = blockIdx.y;
  const int64_t t0 = leaf_offs[l], t1 = leaf_offs[l + 1];
  double s = 0.0, c = 0.0;
  for (int64_t t = t0 + threadIdx.x; t < t1; t += 256) neumaier_add(s, c, v[t * P + p]);
  sh[threadIdx.x] = s;
  sc[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, C = 0.0, plain = 0.0;
    for (int i = 0; i < 256; ++i) { neumaier_add(S, C, sh[i]); C += sc[i]; plain += sh[i]; }
    out[(size_t)l * P + p] = isfinite(plain) ? S + C : plain;   // Inf / NaN propagate like a plain sum
  }
}

__global__ void k_axpby(int64_t n, double a, const double* x, double b, const double* y,
                        double* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    out[e] = a * x[e] + (y ? b * y[e] : 0.0);
}

__global__ void k_fill(int64_t n, double v, double* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    out[e] = v;
}

// diag extraction (+ add): out[i] = M[i*ld + i] + add
__global__ void k_diag(int n, const double* M, int64_t ld, double add, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = M[i * ld + i] + add;
}

// gather: out[i*no + j] = M[ri[i]*ld + ci[j]]
__global__ void k_gather2(const double* M, int64_t ld, const int32_t* ri, int nr,
                          const int32_t* ci, int nc, double* out) {
  const int64_t tot = (int64_t)nr * nc;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot;
       e += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(e / nc), j = (int)(e % nc);
    out[e] = M[(size_t)ri[i] * ld + ci[j]];
  }
}

// Var_j = k(x_j, x_j) + S_jj - sum_i C_ij (Gamma^{-1} C)_ij + exp(noise): the diagonal of
// prediction.R:1189 plus :1194, without forming the G x G matrix.  C column-major no x np,
// GC row-major no x np.
__global__ void k_pred_var(DevKern kd, HpVal hp, const double* xp, int np, int D,
                           const double* sdiag, const double* C, const double* GC, int no,
                           double noise, double* out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  double hpe[MB_MAX_HP];
  prep_hp(kd, hp.v, hpe);
  double kjj = cov_eval<false>(kd, hp.v, hpe, xp + (size_t)j * D, xp + (size_t)j * D, D, false, nullptr);
  double s = 0.0;
  for (int i = 0; i < no; ++i) s += C[i + (size_t)j * no] * GC[(size_t)i * np + j];
  out[j] = ((kjj + sdiag[j]) - s) + noise;
}

// ---- launchers ---------------------------------------------------------------------------
static inline int grid_for(int64_t n, int threads, int cap = 148 * 8) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

cudaError_t launch_dense_cov(const DevKern& kd, const double* hp_host, const double* x1, int n1,
                             const double* x2, int n2, int D, int deriv, int with_noise,
                             double* out, int64_t rs, int64_t cs, cudaStream_t st) {
  HpVal h;
  for (int p = 0; p < MB_MAX_HP; ++p) h.v[p] = p < kd.n_hp ? hp_host[p] : 0.0;
  k_dense_cov<<<grid_for((int64_t)n1 * n2, 256), 256, 0, st>>>(kd, h, x1, n1, x2, n2, D, deriv,
                                                               with_noise, out, rs, cs);
  return cudaGetLastError();
}

cudaError_t launch_legacy_builder(int which, const double* m1, int n1, const double* m2, int n2,
                                  int d, double par, double* out, cudaStream_t st) {
  k_legacy_builder<<<grid_for((int64_t)n1 * n2, 256), 256, 0, st>>>(which, m1, n1, m2, n2, d, par,
                                                                    out);
  return cudaGetLastError();
}


cudaError_t launch_dense_lu_logdet(const double* src, double* W, int n, int* done, double* out,
                                   cudaStream_t st) {
  k_dense_lu_logdet<<<1, 1024, 0, st>>>(src, W, n, done, out);
  return cudaGetLastError();
}

cudaError_t launch_gemv(int m, int n, const double* A, int64_t ld, const double* x,
                        const double* y0, double* y, cudaStream_t st) {
  k_gemv<<<(m + 7) / 8, 256, 0, st>>>(m, n, A, ld, x, y0, y);
  return cudaGetLastError();
}

int dense_obj_nparts(int n) { return grid_for((int64_t)n * n, 256, 148 * 2); }

cudaError_t launch_dense_obj_final3(int n, int P, int nparts, const double* part, const double* r, const double* al,
                                    const double* logdet, int from_chol, int add_trace, double* res, cudaStream_t st);

cudaError_t launch_dense_obj(const DevKern& kd, const double* hp_host, const double* x, int n,
                             int D, const double* A, const double* S, const double* T,
                             const double* al, const double* r, const double* logdet,
                             int from_chol, int want_grad, double* part, double* res,
                             cudaStream_t st) {
  HpVal h;
  for (int p = 0; p < MB_MAX_HP; ++p) h.v[p] = p < kd.n_hp ? hp_host[p] : 0.0;
  int nparts = dense_obj_nparts(n);
  k_dense_obj_partials<<<nparts, 256, 0, st>>>(kd, h, x, n, D, A, S, T, al, want_grad, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_dense_obj_final3(n, want_grad ? kd.n_hp : 0, nparts, part, r, al, logdet, from_chol, S != nullptr, res, st);
}

// out2[0] = max(retries), out2[1] = 1 when some task never factored (retries < 0); out2 is reset here
__global__ void k_retry_stats(const int32_t* retries, int64_t n, int* out2) {
  int mx = 0, bad = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int v = retries[t];
    if (v < 0) bad = 1;
    if (v > mx) mx = v;
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  bad = __reduce_max_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) { if (mx > 0) atomicMax(out2, mx); if (bad) atomicMax(out2 + 1, 1); }
}
cudaError_t launch_retry_stats(const int32_t* retries, int64_t n, int* out2, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(int), st);
  if (e != cudaSuccess) return e;
  k_retry_stats<<<grid_for(n, 256), 256, 0, st>>>(retries, n, out2);
  return cudaGetLastError();
}

cudaError_t launch_reduce_tree(int64_t count, int nparts, const double* base, const double* part, int64_t part_stride,
                               double* out, cudaStream_t st) {
  if (nparts < 1 || nparts > (1 << RT_MAX_LOG)) return cudaErrorInvalidValue;
  if (nparts >= 8 && nparts <= 256) {
    int slog = 0;
    while ((8 << slog) < nparts) ++slog;          // eight subtrees of 2^slog partials cover all of them
    const unsigned nb = (unsigned)((count + 31) / 32);
    switch (slog) {
      case 0: k_reduce_tree8<0><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
      case 1: k_reduce_tree8<1><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
      case 2: k_reduce_tree8<2><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
      case 3: k_reduce_tree8<3><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
      case 4: k_reduce_tree8<4><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
      default: k_reduce_tree8<5><<<nb, 256, 0, st>>>(count, nparts, base, part, part_stride, out); break;
    }
    return cudaGetLastError();
  }
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_reduce_tree<<<(unsigned)blocks, 256, 0, st>>>(count, nparts, base, part, part_stride, out);
  return cudaGetLastError();
}

// out[l * P + p] = compensated sum over the tasks t of leaf l of v[t * P + p]
cudaError_t launch_leaf_sums(const double* v, int P, const int64_t* leaf_offs, int n_leaves, double* out, cudaStream_t st) {
  k_leaf_sum_compensated<<<dim3((unsigned)P, (unsigned)n_leaves), 256, 0, st>>>(v, P, leaf_offs, out);
  return cudaGetLastError();
}

cudaError_t launch_sum_sequential(const double* v, int64_t n, int64_t stride, double* out,
                                  cudaStream_t st) {
  k_sum_compensated<<<1, 256, 0, st>>>(v, n, stride, out);
  return cudaGetLastError();
}

cudaError_t launch_colsum_sequential(const double* v, int64_t n, int P, double* out,
                                     cudaStream_t st) {
  k_sum_compensated<<<P, 256, 0, st>>>(v, n, P, out);
  return cudaGetLastError();
}

cudaError_t launch_axpby(int64_t n, double a, const double* x, double b, const double* y,
                         double* out, cudaStream_t st) {
  k_axpby<<<grid_for(n, 256), 256, 0, st>>>(n, a, x, b, y, out);
  return cudaGetLastError();
}

cudaError_t launch_fill(int64_t n, double v, double* out, cudaStream_t st) {
  k_fill<<<grid_for(n, 256), 256, 0, st>>>(n, v, out);
  return cudaGetLastError();
}

cudaError_t launch_diag(int n, const double* M, int64_t ld, double add, double* out,
                        cudaStream_t st) {
  k_diag<<<(n + 255) / 256, 256, 0, st>>>(n, M, ld, add, out);
  return cudaGetLastError();
}

cudaError_t launch_gather2(const double* M, int64_t ld, const int32_t* ri, int nr,
                           const int32_t* ci, int nc, double* out, cudaStream_t st) {
  k_gather2<<<grid_for((int64_t)nr * nc, 256), 256, 0, st>>>(M, ld, ri, nr, ci, nc, out);
  return cudaGetLastError();
}

cudaError_t launch_pred_var(const DevKern& kd, const double* hp_host, const double* xp, int np, int D,
                            const double* sdiag, const double* C, const double* GC, int no,
                            double noise, double* out, cudaStream_t st) {
  HpVal h;
  for (int p = 0; p < MB_MAX_HP; ++p) h.v[p] = p < kd.n_hp ? hp_host[p] : 0.0;
  k_pred_var<<<(np + 127) / 128, 128, 0, st>>>(kd, h, xp, np, D, sdiag, C, GC, no, noise, out);
  return cudaGetLastError();
}
