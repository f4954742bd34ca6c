__global__ void k16(const double* a, const double* b, double* c) {
  double A[8], B[4], C[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) A[i] = a[threadIdx.x * 8 + i];
  for (int i = 0; i < 4; ++i) B[i] = b[threadIdx.x * 4 + i];
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(C[0]), "+d"(C[1]), "+d"(C[2]), "+d"(C[3])
               : "d"(A[0]), "d"(A[1]), "d"(A[2]), "d"(A[3]), "d"(A[4]), "d"(A[5]), "d"(A[6]), "d"(A[7]), "d"(B[0]), "d"(B[1]), "d"(B[2]), "d"(B[3]));
  for (int i = 0; i < 4; ++i) c[threadIdx.x * 4 + i] = C[i];
}
__global__ void k8(const double* a, const double* b, double* c) {
  double A[2], B[1], C[4] = {0, 0, 0, 0};
  for (int i = 0; i < 2; ++i) A[i] = a[threadIdx.x * 2 + i];
  B[0] = b[threadIdx.x];
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(C[0]), "+d"(C[1]), "+d"(C[2]), "+d"(C[3])
               : "d"(A[0]), "d"(A[1]), "d"(B[0]));
  for (int i = 0; i < 4; ++i) c[threadIdx.x * 4 + i] = C[i];
}
