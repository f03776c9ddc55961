// exact.cuh -- the reference's fp32 distance arithmetic, restated for the exact ("re-check") kernels.
//
// The tensor-core kernel (assign_tc.cu) only *filters*: every assignment / neighbour decision that
// leaves this library is made by the functions below, which reproduce the instruction sequence of
// the reference bit for bit:
//
//   fma_rd(a,b,c)      = __fmaf_rd(a,b,c)                     reference fp_abstraction.h:88-90
//   Kahan "inverted c" = y=fma_rd(a,b,r); t=p+y; r=y-(t-p)    reference kmeans.cu:331-341
//   L2 ranking score   = fma_rd(-2, dot, 0+csqr)              reference metric_abstraction.h:55-57
//   cos ranking score  = clamp-acosf(dot)                     reference metric_abstraction.h:171-177
//   true distances     = sqrt_rn(Kahan sum (a-b)^2) / acosf   reference metric_abstraction.h:59-101,179-222
//
// No fast-math: none of these bodies contains a contractible mul+add pair besides the explicit
// intrinsics, so nvcc's default -fmad=true cannot change them.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

namespace kmb {

#ifndef M_PI_F
#define M_PI_F 3.14159265358979323846f
#endif

struct Kahan {
  float sum, corr;
  __device__ __forceinline__ Kahan() : sum(0.f), corr(0.f) {}
  // one step of sum += a*b
  __device__ __forceinline__ void mac(float a, float b) {
    float y = __fmaf_rd(a, b, corr);
    float t = sum + y;
    corr = y - (t - sum);
    sum = t;
  }
  // one step of sum += (a-b)^2
  __device__ __forceinline__ void sqdiff(float a, float b) {
    float d = a - b;
    mac(d, d);
  }
};

__device__ __forceinline__ float acos_clamped(float p) {
  if (p >= 1.f) return 0.f;
  if (p <= -1.f) return M_PI_F;
  return acosf(p);
}

template <int METRIC>  // 0 = L2, 1 = cosine
__device__ __forceinline__ float lloyd_score(float dot, float csqr) {
  if (METRIC == 1) return acos_clamped(dot);
  return __fmaf_rd(-2.f, dot, 0.f + csqr);
}

template <int METRIC>
__device__ __forceinline__ float finalize_distance(float partial) {
  if (METRIC == 1) return acos_clamped(partial);
  return __fsqrt_rn(partial);
}

// ||c||^2 as the reference computes it (constant 1 for cosine): metric_abstraction.h:21-36,149-158
template <int METRIC>
__device__ __forceinline__ float csqr_exact(const float* __restrict__ c, int D) {
  if (METRIC == 1) return 1.f;
  Kahan k;
  for (int f = 0; f < D; f++) {
    float v = c[f];
    k.mac(v, v);
  }
  return k.sum;
}

// true distance between two row-major vectors in global/shared memory
template <int METRIC>
__device__ __forceinline__ float distance_exact(const float* __restrict__ a,
                                                const float* __restrict__ b, int D) {
  Kahan k;
  if (METRIC == 1) {
    for (int f = 0; f < D; f++) k.mac(a[f], b[f]);
  } else {
    for (int f = 0; f < D; f++) k.sqdiff(a[f], b[f]);
  }
  return finalize_distance<METRIC>(k.sum);
}

}  // namespace kmb
