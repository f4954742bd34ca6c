// lbfgsb_dev.cu -- device side of the lock-step batched L-BFGS-B: one thread per problem
// advances the reverse-communication state machine of lbfgsb.h; states stay in HBM between
// objective launches, so the M step never ships gradients to the host.
// Compiled with -fmad=false so the sums match the host build and the Python oracle bit for bit.
//
// Reference: the per-task stats::optim(..., method = "L-BFGS-B", control = list(factr = 1e13,
// maxit = 25)) calls of m_step / vm_step (/root/reference/R/em-magma.R:200-230,
// R/vem-magmaclust.R:249-261).
#include <cuda_runtime.h>

#include "lbfgsb.h"

__global__ void k_lbfgsb_init(LbfgsbState* st, int64_t m, int n, const double* x0,
                              int64_t x0_stride, double factr, double pgtol, int maxit) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  lb_init(&st[t], n, x0 + t * x0_stride, factr, pgtol, maxit);
}

// Feed (f[t], g[t,:]) to every still-active problem; n_active[0] += problems still active.
// One CTA of 8 warps per 32 problems.  The 2192-byte states of the ACTIVE problems are staged through shared
// memory by all 8 warps (each warp moves 4 states, nine independent coalesced loads in flight per lane -- a
// single warp moving 32 states one after the other spent ~60 us waiting on HBM), then warp 0 advances one
// problem per lane on its staged copy (padded stride of 275 doubles: bank-conflict free), then all warps
// write the states back.  Inactive problems cost one 4-byte read.
#define LB_STRIDE_D 275   /* doubles per staged state: 274 + 1 pad (odd => conflict-free half-warps) */
#define LB_ADV_THREADS 256
__global__ void __launch_bounds__(LB_ADV_THREADS) k_lbfgsb_advance(LbfgsbState* st, int64_t m, int n, const double* f,
                                                                   const double* g, int* n_active) {
  extern __shared__ double lb_smem[];
  __shared__ int act[32];
  static_assert(sizeof(LbfgsbState) == 274 * 8, "LbfgsbState layout changed: update LB_STRIDE_D");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  if (tid < 32) act[tid] = (t0 + tid < m) ? st[t0 + tid].active : 0;
  __syncthreads();
  unsigned any = 0;
  for (int q = 0; q < 32; ++q) any |= (unsigned)act[q];
  if (!any) return;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int q = warp + 8 * h;
    if (!act[q]) continue;
    const double* src = reinterpret_cast<const double*>(&st[t0 + q]);
    double* dst = lb_smem + (size_t)q * LB_STRIDE_D;
    double v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; v[i] = (e < 274) ? src[e] : 0.0; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; if (e < 274) dst[e] = v[i]; }
  }
  __syncthreads();
  if (warp == 0) {
    int still = 0;
    if (act[lane]) {
      const int64_t t = t0 + lane;
      LbfgsbState* s = reinterpret_cast<LbfgsbState*>(lb_smem + (size_t)lane * LB_STRIDE_D);
      double gl[LB_NP];
      for (int p = 0; p < n; ++p) gl[p] = g[t * n + p];
      lb_advance(s, f[t], gl);
      still = s->active;
    }
    const unsigned alive = __ballot_sync(0xffffffffu, still != 0);
    if (lane == 0 && alive) atomicAdd(n_active, __popc(alive));
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int q = warp + 8 * h;
    if (!act[q]) continue;
    double* dst = reinterpret_cast<double*>(&st[t0 + q]);
    const double* src = lb_smem + (size_t)q * LB_STRIDE_D;
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; if (e < 274) dst[e] = src[e]; }
  }
}

// x -> HP table; fail codes and evaluation counts out.
__global__ void k_lbfgsb_export(const LbfgsbState* st, int64_t m, int n, double* x,
                                int64_t x_stride, int* fail, int* nfgv) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  for (int p = 0; p < n; ++p) x[t * x_stride + p] = st[t].x[p];
  if (fail) fail[t] = st[t].fail;
  if (nfgv) nfgv[t] = st[t].nfgv;
}

cudaError_t launch_lbfgsb_init(LbfgsbState* st, int64_t m, int n, const double* x0,
                               int64_t x0_stride, double factr, double pgtol, int maxit,
                               cudaStream_t s) {
  k_lbfgsb_init<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(st, m, n, x0, x0_stride, factr, pgtol,
                                                            maxit);
  return cudaGetLastError();
}

cudaError_t launch_lbfgsb_advance(LbfgsbState* st, int64_t m, int n, const double* f,
                                  const double* g, int* n_active, cudaStream_t s) {
  const size_t smem = (size_t)32 * LB_STRIDE_D * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_lbfgsb_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_lbfgsb_advance<<<(unsigned)((m + 31) / 32), LB_ADV_THREADS, smem, s>>>(st, m, n, f, g, n_active);
  return cudaGetLastError();
}

cudaError_t launch_lbfgsb_export(const LbfgsbState* st, int64_t m, int n, double* x,
                                 int64_t x_stride, int* fail, int* nfgv, cudaStream_t s) {
  k_lbfgsb_export<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(st, m, n, x, x_stride, fail, nfgv);
  return cudaGetLastError();
}

// ---- host self test (mb_lbfgsb_selftest) ----------------------------------------------------
extern "C" int mb_lbfgsb_selftest(int which, double* x, int n, double factr, int maxit,
                                  double* fval, int32_t* n_eval, int32_t* fail) {
  if (which != 0 || n != 2) return -1;
  LbfgsbState s;
  lb_init(&s, n, x, factr, 0.0, maxit);
  while (s.active) {
    const double* v = s.x;
    double f = 100 * (v[1] - v[0] * v[0]) * (v[1] - v[0] * v[0]) + (1 - v[0]) * (1 - v[0]);
    double g[2] = {-400 * v[0] * (v[1] - v[0] * v[0]) - 2 * (1 - v[0]),
                   200 * (v[1] - v[0] * v[0])};
    lb_advance(&s, f, g);
  }
  for (int i = 0; i < n; ++i) x[i] = s.x[i];
  *fval = s.f;
  *n_eval = s.nfgv;
  *fail = s.fail;
  return 0;
}
