// dense_big.cu -- the large-matrix path (n from 225 to tens of thousands): chol_inv_jitter, log-determinant
// and the N x N x N products of logL_GP_mod / gr_GP_mod on matrices that live in HBM, spread over all SMs.
//
//   * k_big_gemm<AK,BK,V16>: C = alpha op(A) op(B) + beta C0 on the FP64 tensor cores (mma.sync.m8n8k4.f64,
//     SASS DMMA), 128 x 128 output tile per CTA, 16 warps of 32 x 32, k-tile 16, four-stage cp.async
//     pipeline; tile masks / k-ranges for triangular operands, optional mirrored store (exactly symmetric
//     output) and an epilogue that contracts the product with the kernel derivatives instead of storing it
//     (gr_GP_mod, /root/reference/R/gradients-likelihoods.R:84-94).
//   * blocked upper Cholesky A = R'R (chol() -> dpotrf('U'), R/kernel-to-matrices.R:675): two-level
//     right-looking, outer block 256, inner 64; 64 x 64 diagonal blocks factored in one CTA's shared memory
//     (dense3.cuh), row panels by TRUE forward substitution (k_big_trsm_rt; like dtrsm -- a product with an
//     explicit inverse loses cond(R_kk) digits on the union-grid matrices, cond ~ 1e12), all trailing
//     updates on DMMA.
//   * inverse (chol2inv() -> dpotri): Y = R^{-T} by the same blocked forward substitution applied to the
//     identity, then A^{-1} = Y'Y as one masked GEMM launch.  chol + inverse = n^3 flops.
//
// Every kernel returns at once when the device fail flag is set, so a failed factorisation (jitter retry,
// R/kernel-to-matrices.R:670-680) costs one flag read per remaining launch; the host loops over the jitter.
#include "dense3.cuh"
#include "dense_big.cuh"

size_t dense_sm_bytes(int n);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// one operand tile of one stage: 1024 pairs of doubles, two per thread
template <bool KMAJ, bool V16>
__device__ __forceinline__ void bg_load_tile(double* dst, const double* src, int64_t ld, int k0, int kend, int m0,
                                             int mend, int tid) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int c = tid + BG_THREADS * r;
    int kk, mm;
    double* d;
    if (KMAJ) { kk = k0 + (c >> 6); mm = m0 + 2 * (c & 63); d = dst + (c >> 6) * BG_LDK + 2 * (c & 63); }
    else { mm = m0 + (c >> 3); kk = k0 + 2 * (c & 7); d = dst + (c >> 3) * BG_LDM + 2 * (c & 7); }
    // the contiguous pair runs along m (KMAJ) or along k
    const int run = KMAJ ? mm : kk, runend = KMAJ ? mend : kend;
    const bool other_ok = KMAJ ? (kk < kend) : (mm < mend);
    int nvalid = other_ok ? (runend - run) : 0;
    nvalid = nvalid < 0 ? 0 : (nvalid > 2 ? 2 : nvalid);
    const double* g = nvalid ? (KMAJ ? src + (int64_t)kk * ld + mm : src + (int64_t)mm * ld + kk) : src;
    if (V16) {
      cp_async16(d, g, 8 * nvalid);
    } else {
      cp_async8(d, g, nvalid >= 1 ? 8 : 0);
      cp_async8(d + 1, nvalid >= 2 ? g + 1 : src, nvalid >= 2 ? 8 : 0);
    }
  }
}

extern __shared__ double4 big_smem_raw[];

// ---- bulk asynchronous copies (TMA engine, SASS UBLKCP) completing on shared-memory mbarriers ----------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(b), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// BULK (k-major operands with 16-byte aligned rows only): the operand tiles of a stage arrive as 2 x 16 bulk copies -- one
// 1 KB row of the k-major tile each, issued by the 32 lanes of warp 0 into the padded rows of the shared-memory tile --
// that complete on the stage's mbarrier; the other 15 warps issue nothing but fragment loads and DMMAs.  Otherwise: the
// cp.async pipeline (every thread copies two 16-byte pieces per operand and stage).
template <bool AK, bool BK, bool V16, bool GRAD, bool BULK>
__global__ void __launch_bounds__(BG_THREADS, 1) k_big_gemm(const __grid_constant__ BigGemmArgs a) {
  static_assert(!BULK || (AK && BK && V16), "bulk copies need k-major operands with aligned rows");
  if (a.fail && *a.fail) return;
  const int m0 = blockIdx.y * BG_BM, n0 = blockIdx.x * BG_BN;
  const int gr0 = a.row0 + m0, gc0 = a.col0 + n0;
  if ((a.flags & BG_UPPER) && gc0 + BG_BN - 1 < gr0) return;
  int klo = 0;
  if (a.flags & BG_KLO_COL) {
    klo = gc0 - a.kbase;
    klo = klo < 0 ? 0 : (klo / BG_BK) * BG_BK;
  }
  const int nk = (a.K - klo + BG_BK - 1) / BG_BK;
  double* sm = (double*)big_smem_raw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if constexpr (BULK) {
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)BG_STAGES * 2 * BG_TILE_D);
    if (tid == 0) {
      for (int s = 0; s < BG_STAGES; ++s) mbar_init(full + s, 32);
      mbar_fence_init();
    }
    __syncthreads();
    // warp 0, lane = 16 * operand + row of the k-tile
    const int op = lane >> 4, row = lane & 15;
    const double* src = op ? a.B : a.A;
    const int64_t ld = op ? a.ldb : a.lda;
    const int x0 = op ? n0 : m0;
    int cnt = (op ? a.N : a.M) - x0;
    cnt = cnt > BG_BM ? BG_BM : cnt;
    auto issue = [&](int kt) {
      const int stage = kt % BG_STAGES;
      double* dst = sm + (size_t)stage * 2 * BG_TILE_D + (size_t)op * BG_TILE_D + (size_t)row * BG_LDK;
      const int kk = klo + kt * BG_BK + row;
      unsigned bytes = 0;
      if (kk < a.K) {
        bytes = 8u * (unsigned)(cnt & ~1);
        if (cnt & 1) { dst[cnt - 1] = src[(int64_t)kk * ld + x0 + cnt - 1]; fence_proxy_async_smem(); }
      } else {   // rows beyond K multiply every output: zeros (only the last k-tile of a ragged K)
        for (int c = 0; c < BG_BM; ++c) dst[c] = 0.0;
        fence_proxy_async_smem();
      }
      mbar_arrive_expect_tx(full + stage, bytes);
      if (bytes) bulk_g2s(dst, src + (int64_t)kk * ld + x0, bytes, full + stage);
    };
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < BG_STAGES - 1; ++s)
        if (s < nk) issue(s);
    }
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(full + (kt % BG_STAGES), (unsigned)((kt / BG_STAGES) & 1));
      __syncthreads();   // every warp has left the stage that is refilled next
      if (warp == 0 && kt + BG_STAGES - 1 < nk) issue(kt + BG_STAGES - 1);
      const double* As = sm + (size_t)(kt % BG_STAGES) * 2 * BG_TILE_D;
      const double* Bs = As + BG_TILE_D;
#pragma unroll
    for (int ks = 0; ks < BG_BK; ks += 4) {
      double fa[4], fb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        fa[i] = AK ? As[(ks + t) * BG_LDK + wm + 8 * i + g] : As[(wm + 8 * i + g) * BG_LDM + ks + t];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        fb[j] = BK ? Bs[(ks + t) * BG_LDK + wn + 8 * j + g] : Bs[(wn + 8 * j + g) * BG_LDM + ks + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
    }
  } else {
    auto load = [&](int kt) {
      double* As = sm + (size_t)(kt % BG_STAGES) * 2 * BG_TILE_D;
      double* Bs = As + BG_TILE_D;
      const int k0 = klo + kt * BG_BK;
      bg_load_tile<AK, V16>(As, a.A, a.lda, k0, a.K, m0, a.M, tid);
      bg_load_tile<BK, V16>(Bs, a.B, a.ldb, k0, a.K, n0, a.N, tid);
    };
  #pragma unroll
    for (int s = 0; s < BG_STAGES - 1; ++s) {
      if (s < nk) load(s);
      cp_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
      cp_wait<BG_STAGES - 2>();
      __syncthreads();
      if (kt + BG_STAGES - 1 < nk) load(kt + BG_STAGES - 1);
      cp_commit();
      const double* As = sm + (size_t)(kt % BG_STAGES) * 2 * BG_TILE_D;
      const double* Bs = As + BG_TILE_D;
  #pragma unroll
      for (int ks = 0; ks < BG_BK; ks += 4) {
        double fa[4], fb[4];
  #pragma unroll
        for (int i = 0; i < 4; ++i)
          fa[i] = AK ? As[(ks + t) * BG_LDK + wm + 8 * i + g] : As[(wm + 8 * i + g) * BG_LDM + ks + t];
  #pragma unroll
        for (int j = 0; j < 4; ++j)
          fb[j] = BK ? Bs[(ks + t) * BG_LDK + wn + 8 * j + g] : Bs[(wn + 8 * j + g) * BG_LDM + ks + t];
  #pragma unroll
        for (int i = 0; i < 4; ++i)
  #pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
      }
    }
    cp_wait<0>();
  }

  if (GRAD) {
    // contraction with the kernel derivatives (upper triangle, off-diagonal weight 2)
    const int P = a.kd.n_hp;
    double hpe[MB_MAX_HP], gacc[MB_MAX_HP], dv[MB_MAX_HP];
    prep_hp(a.kd, a.hp, hpe);
    for (int p = 0; p < MB_MAX_HP; ++p) gacc[p] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = m0 + wm + 8 * i + g;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = n0 + wn + 8 * j + 2 * t + q;
          if (r < a.M && c < a.N && c >= r) {
            const double tij = acc[i][j][q] - a.Ainv[(int64_t)r * a.ldainv + c];
            double w;
            if (r == c) w = a.lr[r] * a.al[r] + tij;
            else w = (a.lr[r] * a.al[c] + a.lr[c] * a.al[r]) + 2.0 * tij;
            cov_eval<true>(a.kd, a.hp, hpe, a.X + (int64_t)r * a.D, a.X + (int64_t)c * a.D, a.D, true, dv);
            for (int p = 0; p < P; ++p) gacc[p] += w * dv[p];
          }
        }
    }
    __syncthreads();   // every warp is done with the operand tiles: reuse them for the reduction
    double* red = sm;
    for (int p = 0; p < P; ++p) {
      double v = gacc[p];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULLMASK, v, off);
      if (lane == 0) red[p * 16 + warp] = v;
    }
    __syncthreads();
    if (tid < P) {
      double v = 0.0;
      for (int w = 0; w < 16; ++w) v += red[tid * 16 + w];
      a.gpart[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MB_MAX_HP + tid] = v;
    }
    return;
  }

  const bool mirror = (a.flags & BG_MIRROR) != 0;
  const bool vec_ok = ((a.ldc & 1) == 0) && ((((uintptr_t)a.C) & 15) == 0) &&
                      (!a.C0 || (((a.ldc0 & 1) == 0) && ((((uintptr_t)a.C0) & 15) == 0)));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + 8 * i + g;
    if (r >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn + 8 * j + 2 * t;
      if (c >= a.N) continue;
      double v0 = a.alpha * acc[i][j][0], v1 = a.alpha * acc[i][j][1];
      const bool two = c + 1 < a.N;
      if (mirror) {
        // upper elements only, each also written to its transposed place
        const int gr = a.row0 + r, gc = a.col0 + c;
        if (gc >= gr) { a.C[(int64_t)r * a.ldc + c] = v0; a.C[(int64_t)c * a.ldc + r] = v0; }
        if (two && gc + 1 >= gr) { a.C[(int64_t)r * a.ldc + c + 1] = v1; a.C[(int64_t)(c + 1) * a.ldc + r] = v1; }
        continue;
      }
      if (vec_ok && two) {
        if (a.C0) {
          const double2 o = *reinterpret_cast<const double2*>(a.C0 + (int64_t)r * a.ldc0 + c);
          v0 += a.beta * o.x; v1 += a.beta * o.y;
        }
        *reinterpret_cast<double2*>(a.C + (int64_t)r * a.ldc + c) = make_double2(v0, v1);
      } else {
        if (a.C0) {
          v0 += a.beta * a.C0[(int64_t)r * a.ldc0 + c];
          if (two) v1 += a.beta * a.C0[(int64_t)r * a.ldc0 + c + 1];
        }
        a.C[(int64_t)r * a.ldc + c] = v0;
        if (two) a.C[(int64_t)r * a.ldc + c + 1] = v1;
      }
    }
  }
}

static const size_t BG_SMEM = (size_t)BG_STAGES * 2 * BG_TILE_D * sizeof(double) + BG_STAGES * sizeof(uint64_t);

template <bool AK, bool BK, bool V16, bool GRAD, bool BULK = false>
static cudaError_t bg_launch_t(const BigGemmArgs& a, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_big_gemm<AK, BK, V16, GRAD, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BG_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_big_gemm<AK, BK, V16, GRAD, BULK><<<grid, BG_THREADS, BG_SMEM, st>>>(a);
  return cudaGetLastError();
}

// MB_BIG_GEMM_LOADER=cpasync keeps the cp.async pipeline for k-major operands too (A/B measurements)
static bool bg_use_bulk() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MB_BIG_GEMM_LOADER"); v = (e && !strcmp(e, "cpasync")) ? 0 : 1; }
  return v == 1;
}

void big_gemm_grid(int M, int N, int* gx, int* gy) { *gx = (N + BG_BN - 1) / BG_BN; *gy = (M + BG_BM - 1) / BG_BM; }

cudaError_t launch_big_gemm(const BigGemmArgs& a, int a_kmajor, int b_kmajor, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return cudaSuccess;
  dim3 grid((a.N + BG_BN - 1) / BG_BN, (a.M + BG_BM - 1) / BG_BM);
  const bool v16 = ((a.lda & 1) == 0) && ((a.ldb & 1) == 0) && ((((uintptr_t)a.A) & 15) == 0) &&
                   ((((uintptr_t)a.B) & 15) == 0);
  if (a.flags & BG_EPI_GRAD) {
    if (!a_kmajor || !b_kmajor) return cudaErrorInvalidValue;
    if (v16 && bg_use_bulk()) return bg_launch_t<true, true, true, true, true>(a, grid, st);
    return v16 ? bg_launch_t<true, true, true, true>(a, grid, st) : bg_launch_t<true, true, false, true>(a, grid, st);
  }
  if (a_kmajor && b_kmajor && v16 && bg_use_bulk()) return bg_launch_t<true, true, true, false, true>(a, grid, st);
#define BG_CASE(ak, bk)                                                                     \
  if ((a_kmajor != 0) == ak && (b_kmajor != 0) == bk)                                       \
    return v16 ? bg_launch_t<ak, bk, true, false>(a, grid, st) : bg_launch_t<ak, bk, false, false>(a, grid, st);
  BG_CASE(true, true)
  BG_CASE(true, false)
  BG_CASE(false, true)
  BG_CASE(false, false)
#undef BG_CASE
  return cudaErrorInvalidValue;
}

// plain product on strided operands (element (i, k) of op(A) at A[i * ars + k * acs], one stride must be 1)
cudaError_t launch_big_gemm_plain(int m, int n, int k, const double* A, int64_t ars, int64_t acs, const double* B,
                                  int64_t brs, int64_t bcs, double* C, int64_t ldc, double alpha, const double* Cin,
                                  double beta, cudaStream_t st) {
  BigGemmArgs a;
  memset(&a, 0, sizeof(a));
  if ((ars != 1 && acs != 1) || (brs != 1 && bcs != 1)) return cudaErrorInvalidValue;
  // op(A)(i, kk): k-major when the m index is the contiguous one
  const int ak = (ars == 1 && acs != 1) ? 1 : 0;
  const int bk = (bcs == 1) ? 1 : 0;
  a.A = A; a.lda = ak ? acs : ars;
  a.B = B; a.ldb = bk ? brs : bcs;
  a.C = C; a.ldc = ldc; a.C0 = Cin; a.ldc0 = ldc; a.M = m; a.N = n; a.K = k; a.alpha = alpha; a.beta = beta;
  return launch_big_gemm(a, ak, bk, st);
}

// ------------------------------------------------------------------------------------------------------
// 64 x 64 diagonal block: upper Cholesky in one CTA's shared memory
__global__ void __launch_bounds__(DSM_THREADS, 1) k_big_potrf_diag(double* W, int64_t ld, int n, int k, int* fail) {
  if (*fail) return;
  const int kb = (n - k) < 64 ? (n - k) : 64;
  const int n8 = (kb + 7) >> 3, np = 8 * n8;
  const SmD s = carve_d((double*)big_smem_raw, n8);
  const Ln L;
  for (int i = threadIdx.x; i < n8; i += DSM_THREADS)
    for (int j = i; j < n8; ++j) s.tab[d3_tri(i, n8) + j - i] = (unsigned short)((i << 8) | j);
  double* Wk = W + (int64_t)k * ld + k;
  for (int e = threadIdx.x; e < np * np; e += DSM_THREADS) {
    const int i = e / np, j = e - i * np;
    const int ti = i >> 3, tj = j >> 3;
    if (ti > tj) continue;
    double v = 0.0;
    if (i < kb && j < kb) { if (j >= i) v = Wk[(int64_t)i * ld + j]; }
    else if (i == j) v = 1.0;
    s.T1[(size_t)(d3_tri(ti, n8) + tj - ti) * 64 + sw(i & 7, j & 7)] = v;
  }
  __syncthreads();
  const int f = d3_chol<NWS>(s, n8, L);
  if (f) { if (threadIdx.x == 0) *fail = k + f; return; }
  for (int e = threadIdx.x; e < np * np; e += DSM_THREADS) {
    const int i = e / np, j = e - i * np;
    if (j < i || i >= kb || j >= kb) continue;
    Wk[(int64_t)i * ld + j] = s.T1[(size_t)(d3_tri(i >> 3, n8) + (j >> 3) - (i >> 3)) * 64 + sw(i & 7, j & 7)];
  }
}

// Z = R_kk^{-T} B in place for a row panel of up to 64 rows: one thread per column, forward substitution
// in registers against the factored diagonal block held in shared memory (broadcast reads).  256 columns per
// CTA so that the 64 x 64 block is fetched with 16 independent loads per thread (with 64 threads the fetch
// alone was 64 dependent round trips to L2, ~20 us).
#define TRS_THREADS 256
__global__ void __launch_bounds__(TRS_THREADS) k_big_trsm_rt(const double* __restrict__ Rkk, int64_t ldr, double* Bm,
                                                             int64_t ldb, int kb, int ncols, const int* fail) {
  if (*fail) return;
  __shared__ double Rs[64][64];
  __shared__ double rinv[64];
  const int tid = threadIdx.x;
  {
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + TRS_THREADS * q, i = e >> 6, j = e & 63;
      v[q] = (i < kb && j < kb && j >= i) ? Rkk[(int64_t)i * ldr + j] : ((i == j && i >= kb) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int e = tid + TRS_THREADS * q; Rs[e >> 6][e & 63] = v[q]; }
  }
  __syncthreads();
  if (tid < 64) rinv[tid] = 1.0 / Rs[tid][tid];
  __syncthreads();
  const int col = blockIdx.x * TRS_THREADS + tid;
  if (col >= ncols) return;
  double x[64];
#pragma unroll
  for (int r = 0; r < 64; ++r) x[r] = (r < kb) ? Bm[(int64_t)r * ldb + col] : 0.0;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    x[j] *= rinv[j];
#pragma unroll
    for (int i = j + 1; i < 64; ++i) x[i] -= Rs[j][i] * x[j];
  }
#pragma unroll
  for (int r = 0; r < 64; ++r)
    if (r < kb) Bm[(int64_t)r * ldb + col] = x[r];
}

// W = src (+ jitter on the diagonal), rows of the upper triangle only matter
// (retries < 0: no jitter; else pen, 10 pen, ... added one after the other like the reference's recursion)
__global__ void k_big_load(const double* __restrict__ src, int64_t ld_src, double* W, int64_t ldw, int n, double pen,
                           int retries) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * n) return;
  const int i = (int)(e / n), j = (int)(e - (int64_t)i * n);
  double v = src[(int64_t)i * ld_src + j];
  if (i == j) { double q = pen; for (int r = 0; r <= retries; ++r) { v += q; q = 10 * q; } }
  W[(int64_t)i * ldw + j] = v;
}
__global__ void k_big_identity(double* Y, int64_t ld, int n) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * ld) return;
  const int i = (int)(e / ld), j = (int)(e - (int64_t)i * ld);
  Y[e] = (i == j) ? 1.0 : 0.0;
}
// info / out in the convention of k_dense_cholinv_sm; fixed-order reduction of sum log diag(R)
__global__ void __launch_bounds__(1024) k_big_finish(const double* W, int64_t ld, int n, const int* fail, int retries,
                                                     double total, int mode, int* info, double* out) {
  __shared__ double red[1024];
  const int f = *fail;
  if (f) {
    if (threadIdx.x == 0) {
      if (mode == 1) { info[0] = f; out[0] = NAN; }
      else { info[0] = -1; out[0] = total; out[1] = NAN; }
    }
    return;
  }
  double v = 0.0;
  for (int j = threadIdx.x; j < n; j += 1024) v += log(W[(int64_t)j * ld + j]);
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (mode == 1) { info[0] = 0; out[0] = red[0]; }
    else { info[0] = retries; out[0] = total; out[1] = red[0]; }
  }
}
__global__ void k_big_clear_flag(int* fail) { *fail = 0; }

#define BIG_OUTER 256
#define BIG_INNER 64

static cudaError_t big_configure() {
  static bool done = false;
  if (done) return cudaSuccess;
  const size_t bytes = dense_sm_bytes(64);
  cudaError_t e = cudaFuncSetAttribute(k_big_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

static cudaError_t big_update(const double* Ak, int64_t lda, const double* Bk, int64_t ldb, double* C, int64_t ldc, int M,
                              int N, int K, int row0, int col0, int kbase, int flags, const int* fail, cudaStream_t st,
                              int64_t* launches) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  BigGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.A = Ak; a.lda = lda; a.B = Bk; a.ldb = ldb; a.C = C; a.ldc = ldc; a.C0 = C; a.ldc0 = ldc;
  a.M = M; a.N = N; a.K = K; a.alpha = -1.0; a.beta = 1.0; a.row0 = row0; a.col0 = col0; a.kbase = kbase;
  a.flags = flags; a.fail = fail;
  if (launches) ++*launches;
  return launch_big_gemm(a, 1, 1, st);
}

#define BIG_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

// in-place blocked upper Cholesky of W (n x n, ld even); *fail receives 0 or (1-based) failing pivot
cudaError_t launch_big_potrf(double* W, int64_t ld, int n, int* fail, cudaStream_t st, int64_t* launches) {
  BIG_OK(big_configure());
  const size_t dsm = dense_sm_bytes(64);
  for (int K0 = 0; K0 < n; K0 += BIG_OUTER) {
    const int Kend = (K0 + BIG_OUTER < n) ? K0 + BIG_OUTER : n;
    for (int k = K0; k < Kend; k += BIG_INNER) {
      const int kb = (n - k) < BIG_INNER ? (n - k) : BIG_INNER;
      k_big_potrf_diag<<<1, DSM_THREADS, dsm, st>>>(W, ld, n, k, fail);
      if (launches) ++*launches;
      const int c1 = k + kb;
      if (c1 >= n) break;
      k_big_trsm_rt<<<(n - c1 + TRS_THREADS - 1) / TRS_THREADS, TRS_THREADS, 0, st>>>(
          W + (int64_t)k * ld + k, ld, W + (int64_t)k * ld + c1, ld, kb, n - c1, fail);
      if (launches) ++*launches;
      // rows c1..Kend of the outer block row: W[c1:Kend, c1:n] -= R[k:c1, c1:Kend]' R[k:c1, c1:n]
      BIG_OK(big_update(W + (int64_t)k * ld + c1, ld, W + (int64_t)k * ld + c1, ld, W + (int64_t)c1 * ld + c1, ld,
                        Kend - c1, n - c1, kb, c1, c1, k, BG_UPPER, fail, st, launches));
    }
    if (Kend < n)
      BIG_OK(big_update(W + (int64_t)K0 * ld + Kend, ld, W + (int64_t)K0 * ld + Kend, ld, W + (int64_t)Kend * ld + Kend,
                        ld, n - Kend, n - Kend, Kend - K0, Kend, Kend, K0, BG_UPPER, fail, st, launches));
  }
  return cudaGetLastError();
}

// Y = R^{-T} (lower triangular, n x n, ldy even); Y is overwritten
cudaError_t launch_big_trtri_t(const double* W, int64_t ld, double* Y, int64_t ldy, int n, const int* fail,
                               cudaStream_t st, int64_t* launches) {
  {
    const int64_t tot = (int64_t)n * ldy;
    k_big_identity<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, ldy, n);
    if (launches) ++*launches;
  }
  for (int K0 = 0; K0 < n; K0 += BIG_OUTER) {
    const int Kend = (K0 + BIG_OUTER < n) ? K0 + BIG_OUTER : n;
    for (int k = K0; k < Kend; k += BIG_INNER) {
      const int kb = (n - k) < BIG_INNER ? (n - k) : BIG_INNER;
      const int c1 = k + kb;
      // rows k..c1 of Y, columns 0..c1
      k_big_trsm_rt<<<(c1 + TRS_THREADS - 1) / TRS_THREADS, TRS_THREADS, 0, st>>>(W + (int64_t)k * ld + k, ld,
                                                                                 Y + (int64_t)k * ldy, ldy, kb, c1, fail);
      if (launches) ++*launches;
      // Y[c1:Kend, 0:c1] -= R[k:c1, c1:Kend]' Y[k:c1, 0:c1]
      if (c1 < Kend)
        BIG_OK(big_update(W + (int64_t)k * ld + c1, ld, Y + (int64_t)k * ldy, ldy, Y + (int64_t)c1 * ldy, ldy, Kend - c1,
                          c1, kb, c1, 0, k, 0, fail, st, launches));
    }
    // Y[Kend:n, 0:Kend] -= R[K0:Kend, Kend:n]' Y[K0:Kend, 0:Kend]   (Y[m, j] = 0 for j > m: k starts at the column)
    if (Kend < n)
      BIG_OK(big_update(W + (int64_t)K0 * ld + Kend, ld, Y + (int64_t)K0 * ldy, ldy, Y + (int64_t)Kend * ldy, ldy,
                        n - Kend, Kend, Kend - K0, Kend, 0, K0, BG_KLO_COL, fail, st, launches));
  }
  return cudaGetLastError();
}

// inv = Y'Y, full and exactly symmetric (upper tiles computed, mirrored on store)
cudaError_t launch_big_lauum(const double* Y, int64_t ldy, double* inv, int64_t ldi, int n, const int* fail,
                             cudaStream_t st, int64_t* launches) {
  BigGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.A = Y; a.lda = ldy; a.B = Y; a.ldb = ldy; a.C = inv; a.ldc = ldi; a.M = n; a.N = n; a.K = n;
  a.alpha = 1.0; a.flags = BG_UPPER | BG_MIRROR | BG_KLO_COL; a.fail = fail;
  if (launches) ++*launches;
  return launch_big_gemm(a, 1, 1, st);
}

// chol (+ inverse) of the matrix already sitting in W (upper triangle read); total_jitter / retries are only
// reported.  mode 0: inv = W^{-1}; mode 1: factor + log-diagonal only.  The fail flag must have been cleared.
cudaError_t launch_big_cholinv_inplace(double* W, double* Y, int64_t ldw, double* inv, int64_t ldi, int n,
                                       double total_jitter, int retries, int mode, int* fail, int* info, double* out,
                                       cudaStream_t st, int64_t* launches) {
  BIG_OK(launch_big_potrf(W, ldw, n, fail, st, launches));
  k_big_finish<<<1, 1024, 0, st>>>(W, ldw, n, fail, retries, total_jitter, mode, info, out);
  if (launches) ++*launches;
  if (mode == 1) return cudaGetLastError();
  BIG_OK(launch_big_trtri_t(W, ldw, Y, ldw, n, fail, st, launches));
  BIG_OK(launch_big_lauum(Y, ldw, inv, ldi, n, fail, st, launches));
  return cudaGetLastError();
}
cudaError_t launch_big_clear_flag(int* fail, cudaStream_t st) {
  k_big_clear_flag<<<1, 1, 0, st>>>(fail);
  return cudaGetLastError();
}

// One attempt of chol_inv_jitter with `retries` escalations (mode 0) or of the plain Cholesky log-diagonal
// (mode 1).  W, Y: n x ldw workspaces (ldw even, >= n); inv: n x n (ld n).  info/out as k_dense_cholinv_sm;
// info[0] = -1 means "this jitter failed" -- the host escalates and calls again.
cudaError_t launch_big_cholinv_attempt(const double* src, int64_t ld_src, double* W, double* Y, int64_t ldw, double* inv,
                                       int n, double pen, int retries, int mode, int* fail, int* info,
                                       double* out, cudaStream_t st, int64_t* launches) {
  k_big_clear_flag<<<1, 1, 0, st>>>(fail);
  const int64_t tot = (int64_t)n * n;
  double total_jitter = 0.0;
  if (mode == 0) { double p = pen; for (int r = 0; r <= retries; ++r) { total_jitter += p; p = 10 * p; } }
  k_big_load<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, ld_src, W, ldw, n, pen, mode == 0 ? retries : -1);
  if (launches) *launches += 2;
  return launch_big_cholinv_inplace(W, Y, ldw, inv, n, n, total_jitter, retries, mode, fail, info, out, st, launches);
}
