// dense_big.cuh -- interface of the large-matrix path (dense_big.cu): DMMA GEMM with triangular masks,
// blocked Cholesky / inverse on matrices resident in HBM.
#pragma once
#include "common.cuh"

#define BG_BM 128
#define BG_BN 128
#define BG_BK 16
#define BG_STAGES 4
#define BG_THREADS 512
#define BG_LDK (BG_BM + 4)   /* k-major tile: [16][132] */
#define BG_LDM (BG_BK + 4)   /* m-major tile: [128][20] */
#define BG_TILE_D 2560       /* doubles reserved per operand per stage (max of 16*132, 128*20) */

enum { BG_UPPER = 1, BG_MIRROR = 2, BG_KLO_COL = 4, BG_EPI_GRAD = 8 };

struct BigGemmArgs {
  const double* A; int64_t lda;   // AK: element (k, m) at A[k * lda + m]; else A[m * lda + k]
  const double* B; int64_t ldb;   // BK: element (k, n) at B[k * ldb + n]; else B[n * ldb + k]
  double* C; int64_t ldc;         // row-major M x N
  const double* C0; int64_t ldc0; // optional addend (may alias C)
  int M, N, K;
  double alpha, beta;
  int row0, col0, kbase;          // global indices of C(0,0) and of k = 0 (triangular logic)
  int flags;
  const int* fail;
  // BG_EPI_GRAD: sum over the upper triangle of  w(i,j) * dK_p(i,j),
  //   w = (i == j ? 1 : 2) * (acc - Ainv(i,j)) + sym(lr al')(i,j);  partial sums per CTA in gpart
  DevKern kd; const double* hp; const double* X; int D;
  const double* Ainv; int64_t ldainv; const double* al; const double* lr;
  double* gpart;                  // gridDim.y * gridDim.x * MB_MAX_HP
};


cudaError_t launch_big_gemm(const BigGemmArgs& a, int a_kmajor, int b_kmajor, cudaStream_t st);
cudaError_t launch_big_gemm_plain(int m, int n, int k, const double* A, int64_t ars, int64_t acs, const double* B,
                                  int64_t brs, int64_t bcs, double* C, int64_t ldc, double alpha, const double* Cin,
                                  double beta, cudaStream_t st);
void big_gemm_grid(int M, int N, int* gx, int* gy);
cudaError_t launch_big_potrf(double* W, int64_t ld, int n, int* fail, cudaStream_t st, int64_t* launches);
cudaError_t launch_big_trtri_t(const double* W, int64_t ld, double* Y, int64_t ldy, int n, const int* fail,
                               cudaStream_t st, int64_t* launches);
cudaError_t launch_big_lauum(const double* Y, int64_t ldy, double* inv, int64_t ldi, int n, const int* fail,
                             cudaStream_t st, int64_t* launches);
cudaError_t launch_big_cholinv_attempt(const double* src, int64_t ld_src, double* W, double* Y, int64_t ldw, double* inv,
                                       int n, double pen, int retries, int mode, int* fail, int* info, double* out,
                                       cudaStream_t st, int64_t* launches);
cudaError_t launch_big_cholinv_inplace(double* W, double* Y, int64_t ldw, double* inv, int64_t ldi, int n,
                                       double total_jitter, int retries, int mode, int* fail, int* info, double* out,
                                       cudaStream_t st, int64_t* launches);
cudaError_t launch_big_clear_flag(int* fail, cudaStream_t st);
