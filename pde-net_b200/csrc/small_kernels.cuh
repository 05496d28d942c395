// Tiny finishing kernels used directly by the C-ABI layer.
#pragma once
#include "common.cuh"

namespace pdenet {

// ---- tiny deterministic finishing kernels ----------------------------------------------------
// out[n*KK + e] = sum over slots of partial[slot][n][e]   (fixed order)
template <typename T>
__global__ void reduce_dk_kernel(const double *partial, int n_slots, int N, int KK, int VLB, T *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * KK) return;
  const int n = i / KK, e = i % KK;
  double s = 0.0;
  for (int sl = 0; sl < n_slots; ++sl) s += partial[((size_t)sl * N + n) * VLB + e];
  out[i] = (T)s;
}
// out[i] = sum over slots of partial[slot][i]
template <typename T> __global__ void reduce_flat_kernel(const double *partial, int n_slots, int n, T *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int sl = 0; sl < n_slots; ++sl) s += partial[(size_t)sl * n + i];
  out[i] = (T)s;
}

} // namespace pdenet
