// covbuild.cu -- list_kern_to_cov (/root/reference/R/kernel-to-matrices.R:586-652): one covariance (or one
// derivative) matrix per task written to HBM, column-major like R.  This is the stand-alone, HBM-bound form of the
// covariance construction (the EM loop never writes per-task matrices).
//
// Layout of the work: persistent CTAs walk the tasks.  A task's rows of the ragged Input table are staged once in
// shared memory with coalesced loads; its matrix is produced in 64 x 64 blocks.  Every kernel of the grammar is an
// even function of (x_i - x_j) or symmetric in (x_i, x_j) term by term (code.cpp:6-131), so K_ij and K_ji are the same
// floating-point number: only blocks on or above the diagonal are evaluated (diagonal blocks: their upper triangle,
// enumerated without idle lanes by folding row p onto row nb-1-p), staged in shared memory and written twice -- once
// as they are, once transposed -- with unit-stride stores on both sides (row pitch 65 doubles: both read patterns of
// the staging tile are bank-conflict free).  Algorithmic bytes per task: 8 n^2 written + 8 n D + 8 P read.
#include "tasks.cuh"

#define CB_THREADS 256
#define CB_TILE 64
#define CB_PITCH 65

// SE (+ noise) is the kernel of every benchmark configuration: SE_FAST keeps its three exponentiated HPs in
// registers (hpe[0..2] = v, exp(-l), exp(noise)) instead of walking the descriptor through local memory.  Same
// expressions and rounding as base_kernel / cov_eval in common.cuh.
template <bool DERIV, bool SE_FAST>
__device__ __forceinline__ double cb_eval(const DevKern& kd, const double* hpl, const double* hpe, const double* xa,
                                          const double* xb, int D, int deriv) {
  if (SE_FAST) {
    double d2 = 0.0, l1 = 0.0;
    for (int c = 0; c < D; ++c) { const double df = xa[c] - xb[c]; d2 += df * df; l1 += fabs(df); }
    const double top = hpe[1] * 0.5 * d2;
    const double k = exp(hpl[0] - top);
    const double nz = (kd.has_noise && l1 < 0.000001) ? hpe[2] : 0.0;
    if (!DERIV) return kd.has_noise ? k + nz : k;
    return deriv == 0 ? k : (deriv == 1 ? k * top : nz);
  }
  if (!DERIV) return cov_eval<false>(kd, hpl, hpe, xa, xb, D, true, nullptr);
  double dv[MB_MAX_HP];
  cov_eval<true>(kd, hpl, hpe, xa, xb, D, true, dv);
  return dv[deriv];
}

// One 64-column block pair (bi <= bj) of a task: evaluate into the staging tile, write it (and its mirror image).
// FULL: a complete 64 x 64 block whose output column starts are 16-byte aligned -- index arithmetic by shifts and
// compile-time constants, double2 stores.
template <bool DERIV, bool SE_FAST, bool FULL>
__device__ __forceinline__ void cb_block(const DevKern& kd, const double* hpl, const double* hpe, const double* xs, int D,
                                         int deriv, double* tile, double* __restrict__ o, int n, int i0, int nbi_, int j0,
                                         int nbj_, bool diag) {
  const int tid = threadIdx.x;
  const int nbi = FULL ? CB_TILE : nbi_, nbj = FULL ? CB_TILE : nbj_;
  if (diag) {
    // upper triangle of an nb x nb block as an (h x (2h+1)) rectangle, h = ceil(nb / 2): row p carries
    // columns p..m-1 of row p followed by the whole of row m-1-p (m = 2h; entries beyond nb are skipped)
    const int h = (nbi + 1) >> 1, m = 2 * h, w = m + 1;
    const float rw = 1.0f / (float)w;
    for (int e = tid; e < h * w; e += CB_THREADS) {
      const int p = FULL ? e / (CB_TILE + 1) : (int)(((float)e + 0.5f) * rw);
      const int q = e - p * w;
      int i, j;
      if (q < m - p) { i = p; j = p + q; }
      else { i = m - 1 - p; j = i + (q - (m - p)); }
      if (FULL || j < nbi)
        tile[i * CB_PITCH + j] = cb_eval<DERIV, SE_FAST>(kd, hpl, hpe, xs + (i0 + i) * D, xs + (j0 + j) * D, D, deriv);
    }
  } else {
    const float rw = 1.0f / (float)nbi;
    for (int e = tid; e < nbi * nbj; e += CB_THREADS) {
      const int j = FULL ? e >> 6 : (int)(((float)e + 0.5f) * rw);
      const int i = e - j * nbi;
      tile[i * CB_PITCH + j] = cb_eval<DERIV, SE_FAST>(kd, hpl, hpe, xs + (i0 + i) * D, xs + (j0 + j) * D, D, deriv);
    }
  }
  __syncthreads();
  if (FULL) {
    // block (bi, bj): rows (i, i + 1) of column j as one 16-byte store; a warp covers one 512-byte column
    for (int e = tid; e < CB_TILE * CB_TILE / 2; e += CB_THREADS) {
      const int j = e >> 5, i = (e & 31) << 1;
      double2 v;
      if (diag) {
        v.x = tile[min(i, j) * CB_PITCH + max(i, j)];
        v.y = tile[min(i + 1, j) * CB_PITCH + max(i + 1, j)];
      } else {
        v.x = tile[i * CB_PITCH + j];
        v.y = tile[(i + 1) * CB_PITCH + j];
      }
      __stcs(reinterpret_cast<double2*>(o + (size_t)(j0 + j) * n + i0 + i), v);
    }
    if (!diag) {
      for (int e = tid; e < CB_TILE * CB_TILE / 2; e += CB_THREADS) {
        const int i = e >> 5, j = (e & 31) << 1;
        const double2 v = make_double2(tile[i * CB_PITCH + j], tile[i * CB_PITCH + j + 1]);
        __stcs(reinterpret_cast<double2*>(o + (size_t)(i0 + i) * n + j0 + j), v);
      }
    }
  } else {
    // block (bi, bj) as it is: element (i, j) -> o[(j0 + j) n + i0 + i], consecutive threads along i
    const float rw = 1.0f / (float)nbi;
    for (int e = tid; e < nbi * nbj; e += CB_THREADS) {
      const int j = (int)(((float)e + 0.5f) * rw);
      const int i = e - j * nbi;
      const double v = (diag && i > j) ? tile[j * CB_PITCH + i] : tile[i * CB_PITCH + j];
      __stcs(o + (size_t)(j0 + j) * n + i0 + i, v);
    }
    if (!diag) {   // mirrored block (bj, bi): element (j, i) -> o[(i0 + i) n + j0 + j], consecutive threads along j
      const float rw2 = 1.0f / (float)nbj;
      for (int e = tid; e < nbi * nbj; e += CB_THREADS) {
        const int i = (int)(((float)e + 0.5f) * rw2);
        const int j = e - i * nbj;
        __stcs(o + (size_t)(i0 + i) * n + j0 + j, tile[i * CB_PITCH + j]);
      }
    }
  }
  __syncthreads();
}

template <bool DERIV, bool SE_FAST, bool D1>
__global__ void __launch_bounds__(CB_THREADS, SE_FAST ? 6 : 4) k_task_cov_sym(TaskData db, DevKern kd, const double* __restrict__ hp,
                                                                 int64_t hp_stride, int deriv, double* __restrict__ out,
                                                                 const int64_t* __restrict__ out_offs) {
  extern __shared__ double cb_sh[];
  double* tile = cb_sh;                         // CB_TILE x CB_PITCH
  double* hps = cb_sh + CB_TILE * CB_PITCH;     // 2 x MB_MAX_HP: the task's HPs and their exponentials
  double* xs = hps + 2 * MB_MAX_HP;             // nmax x D
  const int tid = threadIdx.x;
  const int D = D1 ? 1 : db.D;
  // Software pipeline over the tasks of this CTA: the next task's table rows, HPs and output offset are fetched into
  // registers while the current task is evaluated, so the dependent chain offs -> X -> barrier is off the critical path
  // (rows beyond CB_THREADS values per task are fetched in place, unpipelined).
  int64_t nx_row0 = 0, nx_oo = 0;
  int nx_n = 0;
  double nx_x = 0.0, nx_hp = 0.0;
  auto prefetch = [&](int64_t t) {
    if (t >= db.n_tasks) return;
    nx_row0 = db.offs[t];
    nx_n = (int)(db.offs[t + 1] - nx_row0);
    nx_oo = out_offs[t];
    if (tid < nx_n * D) nx_x = db.X[nx_row0 * D + tid];
    if (tid < kd.n_hp) nx_hp = hp[t * hp_stride + tid];
  };
  prefetch(blockIdx.x);
  for (int64_t t = blockIdx.x; t < db.n_tasks; t += gridDim.x) {
    const int64_t row0 = nx_row0, oo = nx_oo;
    const int n = nx_n;
    __syncthreads();                            // previous task's tile, HPs and xs are no longer read
    if (tid < n * D) xs[tid] = nx_x;
    for (int e = tid + CB_THREADS; e < n * D; e += CB_THREADS) xs[e] = db.X[row0 * D + e];
    // the exponentials of the HPs once per task (not once per thread), broadcast through shared memory
    if (SE_FAST) {
      if (tid < 3) {
        const double h = (tid < 2 || kd.has_noise) ? nx_hp : 0.0;
        hps[tid] = h;
        hps[MB_MAX_HP + tid] = tid == 0 ? 0.0 : (tid == 1 ? exp(-h) : (kd.has_noise ? exp(h) : 0.0));
      }
    } else {
      if (tid < kd.n_hp) hps[tid] = nx_hp;
      __syncthreads();
      if (tid == 0) {
        double hl[MB_MAX_HP], he[MB_MAX_HP];
        for (int p = 0; p < kd.n_hp; ++p) hl[p] = hps[p];
        prep_hp(kd, hl, he);
        for (int p = 0; p < kd.n_hp; ++p) hps[MB_MAX_HP + p] = he[p];
      }
    }
    prefetch(t + gridDim.x);
    __syncthreads();
    double hpl[MB_MAX_HP], hpe[MB_MAX_HP];
    if (SE_FAST) { hpl[0] = hps[0]; hpe[1] = hps[MB_MAX_HP + 1]; hpe[2] = hps[MB_MAX_HP + 2]; }
    else { for (int p = 0; p < kd.n_hp; ++p) { hpl[p] = hps[p]; hpe[p] = hps[MB_MAX_HP + p]; } }
    double* o = out + oo;
    const bool aligned = ((n & 1) == 0) && ((oo & 1) == 0);
    const int nblk = (n + CB_TILE - 1) / CB_TILE;
    for (int bj = 0; bj < nblk; ++bj) {
      const int j0 = bj * CB_TILE, nbj = min(CB_TILE, n - j0);
      for (int bi = 0; bi <= bj; ++bi) {
        const int i0 = bi * CB_TILE, nbi = min(CB_TILE, n - i0);
        if (aligned && nbi == CB_TILE && nbj == CB_TILE)
          cb_block<DERIV, SE_FAST, true>(kd, hpl, hpe, xs, D, deriv, tile, o, n, i0, nbi, j0, nbj, bi == bj);
        else
          cb_block<DERIV, SE_FAST, false>(kd, hpl, hpe, xs, D, deriv, tile, o, n, i0, nbi, j0, nbj, bi == bj);
      }
    }
  }
}

size_t task_cov_sym_smem(int nmax, int D) { return ((size_t)CB_TILE * CB_PITCH + 2 * MB_MAX_HP + (size_t)nmax * D) * sizeof(double); }

// Returns cudaErrorInvalidValue when the task rows do not fit the staging buffer (the caller then uses k_task_cov).
cudaError_t launch_task_cov_sym(TaskData db, DevKern kd, const double* hp, int64_t hp_stride, int deriv, double* out,
                                const int64_t* out_offs, int sm_count, cudaStream_t st) {
  const size_t smem = task_cov_sym_smem(db.nmax, db.D);
  const bool se_grid = kd.n_terms == 1 && kd.types[0] == MB_KERN_SE && smem <= 36 * 1024;
  if (smem > 56 * 1024) return cudaErrorInvalidValue;
  int64_t g = (int64_t)sm_count * (se_grid ? 6 : 4);
  if (g > db.n_tasks) g = db.n_tasks;
  if (g < 1) return cudaSuccess;
  const bool se = kd.n_terms == 1 && kd.types[0] == MB_KERN_SE;
#define CB_LAUNCH(DV, SE, D1)                                                                                        \
  do {                                                                                                               \
    cudaFuncSetAttribute(k_task_cov_sym<DV, SE, D1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    k_task_cov_sym<DV, SE, D1><<<(unsigned)g, CB_THREADS, smem, st>>>(db, kd, hp, hp_stride, deriv, out, out_offs);  \
  } while (0)
  if (deriv < 0) {
    if (se && db.D == 1) CB_LAUNCH(false, true, true);
    else if (se) CB_LAUNCH(false, true, false);
    else CB_LAUNCH(false, false, false);
  } else {
    if (se) CB_LAUNCH(true, true, false);
    else CB_LAUNCH(true, false, false);
  }
#undef CB_LAUNCH
  return cudaGetLastError();
}
