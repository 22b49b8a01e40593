// Small device utilities shared by the BA kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ba_types.cuh"

namespace svin {

// ------------------------------------------------------------------------------------------ utilities
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum K per-thread values over the CTA and atomically add them to K targets.
template <int K, int THREADS>
__device__ __forceinline__ void block_atomic_add(double (&v)[K], double* const (&dst)[K]) {
  __shared__ double red[K][THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += red[threadIdx.x][i];
    if (s != 0.0) atomicAdd(dst[threadIdx.x], s);
  }
  __syncthreads();
}
// Recursive-halving reduce-scatter over a warp: on return v[0] of lane l holds the sum over all lanes of
// the original v[l].  31 shuffles instead of the 160 a butterfly all-reduce of 32 values needs.
__device__ __forceinline__ void warp_reduce_scatter32(double (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? v[i] : v[i + h];
      const double keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ceres::CauchyLoss / HuberLoss + Corrector
__device__ __forceinline__ void loss_eval(int type, double a, double s, double& rho0, double& rho1, double& rho2) {
  if (type == SVIN_LOSS_CAUCHY) {
    const double bb = a * a, c = 1.0 / bb;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho0 = bb * log(sum);
    rho1 = fmax(inv, 2.2250738585072014e-308);
    rho2 = -c * (inv * inv);
  } else if (type == SVIN_LOSS_HUBER) {
    const double bb = a * a;
    if (s > bb) {
      const double r = sqrt(s);
      rho0 = 2.0 * a * r - bb;
      rho1 = fmax(a / r, 2.2250738585072014e-308);
      rho2 = -rho1 / (2.0 * s);
    } else {
      rho0 = s;
      rho1 = 1.0;
      rho2 = 0.0;
    }
  } else {
    rho0 = s;
    rho1 = 1.0;
    rho2 = 0.0;
  }
}


}  // namespace svin
