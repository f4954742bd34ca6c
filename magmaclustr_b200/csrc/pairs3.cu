// pairs3.cu -- pair tables of the resident tasks (tasks.cuh: TaskData::pair_d2 / pair_idx / pair_n).
//
// The reference evaluates a kernel element by element (src/code.cpp:6-131 under kern_to_cov, R/kernel-to-matrices.R:71-503):
// n (n + 1) / 2 exponentials per task and objective evaluation.  A stationary kernel sees a pair of points only through
// d2 = sum_q (x_rq - x_cq)^2, and tasks observed on a common grid have few DISTINCT values of d2 (BASELINE config 2: ~385
// for the 2080 pairs of a 64-point task).  This kernel finds them once per upload: one CTA per task inserts the bit pattern
// of every pair's d2 into a shared-memory hash set, numbers the occupied slots, and writes
//   pair_d2[task][id]        the distinct values,
//   pair_idx[task][element]  for every element of the packed upper 8 x 8 tiles, in the order of the DMMA accumulator
//                            fragments the matrix is built in: id | (sum_q |x_rq - x_cq| < 1e-6) << 15,
//   pair_n[task]             the count, or -1 when there are more than MB_PAIR_CAP (irregular inputs: nothing to share) or a
//                            NaN among the inputs -- those tasks keep the element-by-element evaluation.
// The per-task kernels (tasks3.cuh: build_cov3) then evaluate exp(v - hel * d2) once per distinct d2 and look the elements
// up: the same expression on the same operand, so every matrix entry has the bits of the direct evaluation
// (tests/test_gpu_magma.py::test_pair_tables_are_bit_identical).  Slot numbering may differ from run to run (the inserts
// race), the looked-up values cannot.
#include "tasks3.cuh"

#define PAIR_HCAP 2048                       /* hash slots (a power of two, < 2^15: slot numbers travel in 15 bits) */
#define PAIR_EMPTY 0xffffffffffffffffull     /* a NaN pattern: never the bits of a squared distance that is inserted */

extern __shared__ double4 pairs3_smem_raw[];

__device__ __forceinline__ unsigned pair_hash(unsigned long long k) {
  k ^= k >> 29;
  k *= 0x9e3779b97f4a7c15ull;
  return (unsigned)(k >> 53);   // 11 bits = PAIR_HCAP slots
}

#define PAIR_MAX_PROBE 128   /* a longer probe sequence means a table far beyond MB_PAIR_CAP entries: give up on the task */

template <int N8C, bool D1>
__global__ void __launch_bounds__(T3_THREADS) k_task_pairs3(TaskData db, double* pair_d2, unsigned short* pair_idx, int32_t* pair_n,
                                                            int pair_stride) {
  constexpr int TPW = (Lay<N8C>::NTRI + NW3 - 1) / NW3;   // tiles per warp
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(pairs3_smem_raw);
  unsigned short* ids = reinterpret_cast<unsigned short*>(keys + PAIR_HCAP);
  int* ctl = reinterpret_cast<int*>(ids + PAIR_HCAP);   // [1] give up, [2..5] warp counts
  unsigned short* tiles = reinterpret_cast<unsigned short*>(ctl + 8);   // upper tiles row by row: (i << 8) | j, NTRI entries
  double* xs = reinterpret_cast<double*>(tiles + ((Lay<N8C>::NTRI + 3) & ~3));
  const int64_t task = blockIdx.x;
  const int64_t row0 = db.offs[task];
  const int n = (int)(db.offs[task + 1] - row0);
  const int n8 = t3_n8(n), np = 8 * n8, ntri = t3_ntri(n8), D = db.D;
  const int tid = threadIdx.x;
  const Ln L;
  for (int e = tid; e < np * D; e += T3_THREADS) xs[e] = (e < n * D) ? db.X[row0 * D + e] : 0.0;
  for (int h = tid; h < PAIR_HCAP; h += T3_THREADS) keys[h] = PAIR_EMPTY;
  if (tid < 8) ctl[tid] = 0;
  if (tid < n8) {
    int o = tid * n8 - ((tid * (tid - 1)) >> 1);
    for (int j = tid; j < n8; ++j) tiles[o++] = (unsigned short)((tid << 8) | j);
  }
  __syncthreads();
  Cov<KS_SE, D1> cov;      // only sums(): the very expression the matrix builders evaluate
  cov.xs = xs; cov.D = D;
  // ---- pass 1: the set of distinct values; every element remembers the slot of its value (two 16-bit slots per tile)
  unsigned slots[TPW];
  bool giveup = false;
#pragma unroll
  for (int u = 0; u < TPW; ++u) {
    const int idx = L.warp + u * NW3;
    slots[u] = 0;
    if (idx < ntri) {
      const int i = tiles[idx] >> 8, j = tiles[idx] & 255;
      const int r = 8 * i + L.g;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 8 * j + 2 * L.t + q;
        if (r < n && c < n) {
          double d2, l1;
          cov.sums(r, c, d2, l1);
          if (!(d2 == d2)) { giveup = true; continue; }
          const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
          unsigned h = pair_hash(key);
          int probes = 0;
          for (;;) {
            unsigned long long old = keys[h];
            if (old == PAIR_EMPTY) old = atomicCAS(keys + h, PAIR_EMPTY, key);
            if (old == PAIR_EMPTY || old == key) break;
            h = (h + 1) & (PAIR_HCAP - 1);
            if (++probes > PAIR_MAX_PROBE) { giveup = true; break; }
          }
          slots[u] |= (h | (l1 < 0.000001 ? 0x8000u : 0u)) << (16 * q);
        }
      }
    }
  }
  if (giveup) ctl[1] = 1;
  __syncthreads();
  if (ctl[1]) {
    if (tid == 0) pair_n[task] = -1;
    return;
  }
  // ---- number the occupied slots in slot order (16 consecutive slots per thread, CTA-wide exclusive scan)
  constexpr int SPT = PAIR_HCAP / T3_THREADS;
  int mine = 0;
#pragma unroll
  for (int k = 0; k < SPT; ++k) mine += (keys[tid * SPT + k] != PAIR_EMPTY);
  int incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(FULLMASK, incl, off);
    if (L.lane >= off) incl += v;
  }
  if (L.lane == 31) ctl[2 + L.warp] = incl;
  __syncthreads();
  int base = incl - mine;
  for (int w = 0; w < L.warp; ++w) base += ctl[2 + w];
  const int total = ctl[2] + ctl[3] + ctl[4] + ctl[5];
  if (total > MB_PAIR_CAP) {   // uniform
    if (tid == 0) pair_n[task] = -1;
    return;
  }
  double* dv = pair_d2 + task * MB_PAIR_CAP;
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    const unsigned long long key = keys[tid * SPT + k];
    if (key != PAIR_EMPTY) {
      ids[tid * SPT + k] = (unsigned short)base;
      dv[base] = __longlong_as_double((long long)key);
      ++base;
    }
  }
  __syncthreads();
  // ---- pass 2: slot -> index (two 16-bit entries per lane and tile, written as one word)
  unsigned* out = reinterpret_cast<unsigned*>(pair_idx + task * pair_stride);
#pragma unroll
  for (int u = 0; u < TPW; ++u) {
    const int idx = L.warp + u * NW3;
    if (idx < ntri) {
      const int i = tiles[idx] >> 8, j = tiles[idx] & 255;
      const int r = 8 * i + L.g;
      unsigned word = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 8 * j + 2 * L.t + q;
        const unsigned sl = (slots[u] >> (16 * q)) & 0xffffu;
        if (r < n && c < n) word |= ((unsigned)ids[sl & 0x7fffu] | (sl & 0x8000u)) << (16 * q);
      }
      out[(Lay<N8C>::tri(i) + j - i) * 32 + L.lane] = word;
    }
  }
  if (tid == 0) pair_n[task] = total;
}

// pair_stride: 16-bit entries per task = 64 x the packed upper tiles of the layout the per-task kernels will use for this
// dataset (8 tiles per side for nmax <= 64, else 12).  Tasks of more than TK_MAXN points have no tables.
cudaError_t launch_task_pairs3(TaskData db, double* pair_d2, unsigned short* pair_idx, int32_t* pair_n, int pair_stride,
                               cudaStream_t st) {
  if (db.n_tasks <= 0) return cudaSuccess;
  if (db.nmax > TK_MAXN) return cudaErrorInvalidValue;
  const bool small = db.nmax <= 64, d1 = db.D == 1;
  const int np = small ? 64 : 96;
  const size_t smem = (size_t)PAIR_HCAP * (8 + 2) + 32 + (size_t)(small ? 36 : 80) * 2 + (size_t)np * db.D * 8;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaSuccess;
#define PAIRS_GO(N8C_, D1_)                                                                                                     \
  do {                                                                                                                          \
    e = cudaFuncSetAttribute(k_task_pairs3<N8C_, D1_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
    if (e == cudaSuccess) k_task_pairs3<N8C_, D1_><<<(unsigned)db.n_tasks, T3_THREADS, smem, st>>>(db, pair_d2, pair_idx, pair_n, pair_stride); \
  } while (0)
  if (small) { if (d1) PAIRS_GO(8, true); else PAIRS_GO(8, false); }
  else { if (d1) PAIRS_GO(12, true); else PAIRS_GO(12, false); }
#undef PAIRS_GO
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}
