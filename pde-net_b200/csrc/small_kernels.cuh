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
__global__ void reduce_dk_kernel(const double *partial, int n_slots, int N, int KK, int VLB, int n_out, T *out) {
  // N = filters per slot in the partial layout; n_out <= N*KK = elements of `out` (the launch is rounded up to whole
  // blocks, and a caller may reduce only the first filters of the layout: never write past n_out)
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n_out) return;
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
// both finishing reductions of a reverse rollout in ONE launch (a small pass is a chain of dependent launches of a
// few microseconds each): warps [0, n_dk) do what reduce_dk_kernel does, warps [n_dk, n_dk + n_sym) what
// reduce_flat_kernel does -- same addends in the same order, bit-identical to the two separate launches
template <typename T>
__global__ void reduce_grads_kernel(const double *dk_partial, int dk_slots, int N, int KK, int VLB, int n_dk, T *dk_out,
                                    const double *sym_partial, int sym_slots, int n_sym, T *sym_out) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  if (i < n_dk) {
    const int n = i / KK, e = i % KK;
    for (int sl = lane; sl < dk_slots; sl += 32) s += dk_partial[((size_t)sl * N + n) * VLB + e];
    s = warp_sum_fixed(s);
    if (lane == 0) dk_out[i] = (T)s;
    return;
  }
  i -= n_dk;
  if (i >= n_sym) return;
  for (int sl = lane; sl < sym_slots; sl += 32) s += sym_partial[(size_t)sl * n_sym + i];
  s = warp_sum_fixed(s);
  if (lane == 0) sym_out[i] = (T)s;
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


// ---- piecewise Lagrange interpolation on a uniform 1-D mesh (row f-2) -----------------------------
// The nonlinear source term f(u) of PDE-Net 1's SingleNonLinear2d (aTEAM LagrangeInterp with m = 1
// [RECALL]): `mesh` cells of width h on [lo, lo + mesh*h], DEG+1 equispaced nodes per cell (shared
// end nodes), coefficients coef[mesh*DEG + 1] = the values at the nodes.  Points outside the mesh
// use the first / last cell's polynomial (cell index clamped, local coordinate not).
template <typename T, int DEG> __device__ __forceinline__ void lagrange_basis(T t, T (&l)[DEG + 1], T (&dl)[DEG + 1]) {
#pragma unroll
  for (int j = 0; j <= DEG; ++j) {
    T v = T(1), dv = T(0);
#pragma unroll
    for (int m = 0; m <= DEG; ++m) {
      if (m == j) continue;
      const T inv = T(DEG) / T(j - m), fac = (t - T(m) / T(DEG)) * inv; // (t - s_m) / (s_j - s_m), s_m = m/DEG
      dv = dv * fac + v * inv;
      v *= fac;
    }
    l[j] = v;
    dl[j] = dv;
  }
}
template <typename T> __device__ __forceinline__ int lagrange_cell(T x, T lo, T inv_h, int mesh, T &t) {
  const T s = (x - lo) * inv_h;
  int c = (int)floor((double)s);
  c = c < 0 ? 0 : (c >= mesh ? mesh - 1 : c);
  t = s - (T)c;
  return c;
}
template <typename T, int DEG>
__global__ void __launch_bounds__(256) lagrange1d_fwd_kernel(const T *x, const T *coef, T *y, size_t n, T lo, T inv_h,
                                                             int mesh) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T t, l[DEG + 1], dl[DEG + 1];
    const int c = lagrange_cell(x[i], lo, inv_h, mesh, t);
    lagrange_basis<T, DEG>(t, l, dl);
    T v = T(0);
#pragma unroll
    for (int j = 0; j <= DEG; ++j) v = fma_(coef[c * DEG + j], l[j], v);
    y[i] = v;
  }
}
// gx[i] = gy[i] * f'(x[i]) (may be NULL); partial[block][q] = this block's share of dLoss/dcoef[q].
// Every thread owns private bins in shared memory ([Q][NT] doubles), the block sums them in a fixed
// order: deterministic, no atomics.
constexpr int kLagrangeThreads = 128;
template <typename T, int DEG>
__global__ void __launch_bounds__(kLagrangeThreads) lagrange1d_bwd_kernel(const T *x, const T *gy, const T *coef, T *gx,
                                                                          double *partial, size_t n, T lo, T inv_h,
                                                                          int mesh) {
  extern __shared__ double bins[]; // [Q][NT]
  constexpr int NT = kLagrangeThreads;
  const int Q = mesh * DEG + 1, tid = threadIdx.x;
  for (int q = 0; q < Q; ++q) bins[q * NT + tid] = 0.0;
  for (size_t i = (size_t)blockIdx.x * NT + tid; i < n; i += (size_t)gridDim.x * NT) {
    T t, l[DEG + 1], dl[DEG + 1];
    const int c = lagrange_cell(x[i], lo, inv_h, mesh, t);
    lagrange_basis<T, DEG>(t, l, dl);
    const T g = gy[i];
    T d = T(0);
#pragma unroll
    for (int j = 0; j <= DEG; ++j) {
      d = fma_(coef[c * DEG + j], dl[j], d);
      bins[(c * DEG + j) * NT + tid] += (double)g * (double)l[j];
    }
    if (gx) gx[i] = g * d * inv_h;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int q = warp; q < Q; q += NT / 32) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) s += bins[q * NT + k * 32 + lane];
    s = warp_sum_fixed(s);
    if (lane == 0) partial[(size_t)blockIdx.x * Q + q] = s;
  }
}

// launch geometry for the two reduce kernels: 4 outputs (warps) per 128-thread block
inline unsigned reduce_blocks(int n_out) { return (unsigned)((n_out + 3) / 4); }

} // namespace pdenet
