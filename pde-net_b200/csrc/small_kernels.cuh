// Tiny finishing kernels used directly by the C-ABI layer.
#pragma once
#include "common.cuh"

namespace pdenet {

// ---- deterministic finishing reductions --------------------------------------------------------
// One warp per output element: lane l adds slots l, l+32, ... in order, then a fixed shuffle tree.
// (The first version used one thread per output looping over all ~1800 slots: 0.2 ms of pure
// latency, half of a config-1 pass.)
__device__ __forceinline__ double warp_sum_fixed(double s) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

// out[n*KK + e] = sum over slots of partial[slot][n][e]
template <typename T>
__global__ void reduce_dk_kernel(const double *partial, int n_slots, int N, int KK, int VLB, T *out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N * KK) return;
  const int n = i / KK, e = i % KK;
  double s = 0.0;
  for (int sl = lane; sl < n_slots; sl += 32) s += partial[((size_t)sl * N + n) * VLB + e];
  s = warp_sum_fixed(s);
  if (lane == 0) out[i] = (T)s;
}
// out[i] = sum over slots of partial[slot][i]
template <typename T> __global__ void reduce_flat_kernel(const double *partial, int n_slots, int n, T *out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int sl = lane; sl < n_slots; sl += 32) s += partial[(size_t)sl * n + i];
  s = warp_sum_fixed(s);
  if (lane == 0) out[i] = (T)s;
}
// ---- fused MSE loss + gradient seed (row f-3) ----------------------------------------------------
constexpr int kMseBlocks = 512; // partial sums per launch (fixed => reproducible)
struct MseWeights {
  double w[64];
};
// grid (kMseBlocks, steps): block (bx, t) strides over step t; partial[t*kMseBlocks + bx] = its sum
template <typename T>
__global__ void __launch_bounds__(256) mse_seed_kernel(const T *states, const T *observed, T *grad, double *partial,
                                                       size_t n, MseWeights wts) {
  const int t = blockIdx.y;
  const T *s = states + (size_t)t * n, *o = observed + (size_t)t * n;
  T *g = grad + (size_t)t * n;
  const double scale = 2.0 * wts.w[t] / (double)n;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double d = (double)s[i] - (double)o[i];
    acc += d * d;
    g[i] = (T)(scale * d);
  }
  __shared__ double red[8];
  acc = warp_sum_fixed(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    partial[(size_t)t * gridDim.x + blockIdx.x] = tot * wts.w[t] / (double)n;
  }
}
__global__ void mse_finish_kernel(const double *partial, int count, double *loss_out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += 32) acc += partial[i];
  acc = warp_sum_fixed(acc);
  if (threadIdx.x == 0) *loss_out = acc;
}

// launch geometry for the two kernels above: 4 outputs (warps) per 128-thread block
inline unsigned reduce_blocks(int n_out) { return (unsigned)((n_out + 3) / 4); }

} // namespace pdenet
