// Dependent-issue latencies on the SM (one warp, clock64 around a chain of N dependent operations).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 0.5;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
  // DMUL + DADD (no contraction) chain
  x = seed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __dadd_rn(__dmul_rn(x, y), z);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  out[32 + threadIdx.x] = x;
  // rsqrt chain
  x = seed + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  out[64 + threadIdx.x] = x;
  // exp chain
  x = seed * 0.1;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = exp(-x) * 0.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  out[96 + threadIdx.x] = x;
  // 64-bit shuffle chain
  x = seed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  out[128 + threadIdx.x] = x;
  // DMMA dependent chain
  double c0 = 0, c1 = 0;
  x = seed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma884(c0, c1, x, y);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  out[160 + threadIdx.x] = c0 + c1;
  // DMMA independent (8 accumulators) throughput, one warp
  double a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) dmma884(a0[q], a1[q], x, y);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  double sacc = 0;
  for (int q = 0; q < 8; ++q) sacc += a0[q] + a1[q];
  out[192 + threadIdx.x] = sacc;
  // shared-memory dependent load chain (pointer chasing)
  if (threadIdx.x < 64) sm[threadIdx.x] = (double)((threadIdx.x * 7 + 3) & 63);
  __syncwarp();
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx];
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[224 + threadIdx.x] = idx;
  // division chain
  x = seed + 3.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 + 2.0 / x;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[256 + threadIdx.x] = x;
  // independent DFMA throughput (8 chains)
  double f[8];
  for (int q = 0; q < 8; ++q) f[q] = seed + q;
  t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = fma(f[q], y, z);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = t1 - t0;
  sacc = 0;
  for (int q = 0; q < 8; ++q) sacc += f[q];
  out[288 + threadIdx.x] = sacc;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 16 * 8);
  k<<<1, 32>>>(out, cyc, 1.25);
  k<<<1, 32>>>(out, cyc, 1.25);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dependent", "DMUL+DADD dependent (pair)", "rsqrt(double) + add", "exp(double) * c", "shfl 64-bit",
                         "DMMA dependent", "DMMA 8 independent (per DMMA)", "LDS dependent (+cvt)", "1 + 2/x (div + add)",
                         "DFMA 8 independent (per DFMA)"};
  for (int i = 0; i < 10; ++i) printf("%-34s %7.1f cycles\n", names[i], (double)h[i] / N);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
