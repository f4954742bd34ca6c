// tiles.cuh -- the FP64 tensor-core primitive (mma.sync.m8n8k4.f64 -> SASS DMMA) shared by the tile kernels.
#pragma once
#include "common.cuh"

#define FULLMASK 0xffffffffu

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
