// Small-grid path of the PDE-Net 1 linear rollout (row a12, BASELINE config 2): ONE launch per direction,
// one trajectory per thread-block CLUSTER, the state (forward) / the adjoint (reverse) resident in shared
// memory across all delta-t blocks, halo rows read from the neighbouring CTAs' shared memory (DSMEM), one
// cluster barrier per step.  No inter-launch latency, no HBM round trip of the derivative adjoints.
//
// It also uses what makes the linear model special: its filters and coefficient fields do not change
// during a rollout, so one step is ONE variable-coefficient k x k stencil
//     u+(p) = sum_ab w_ab(p) u(p + ab - r),      w_ab(p) = K0[ab] + dt * sum_n c_n(p) K_n[ab],
// evaluated with k^2 instead of (N+1) k^2 multiply-adds per point.  The reverse pass needs, besides the
// transposed stencil  lam_t(q) = sum_ab w_ab(q - ab + r) lam_{t+1}(q - ab + r), only the tap-correlation
// fields  R_ab(p) = sum_t lam_{t+1}(p) u_t(p + ab - r)  (k^2 accumulators per point, kept in registers over
// the whole rollout), from which every parameter gradient follows once per rollout:
//     dK0[ab] = sum_p R_ab(p),   dK_n[ab] = dt sum_p c_n(p) R_ab(p),   dc_n(p) = dt sum_ab K_n[ab] R_ab(p).
// Same mathematics as the tiled path (results differ by re-association only); ~16x fewer flops.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace pdenet {
namespace cg = cooperative_groups;

constexpr int kLinClusterThreads = 256;

template <typename T> struct LinClusterParams {
  const T *u0;       // [B][H][W]
  const T *k0;       // [K][K]           learnable order-(0,0) filter D0
  const T *kernels;  // [N][K][K]
  const T *coe;      // [N][H][W]
  T *states;         // forward: out [steps][B][H][W]; reverse: in (may be NULL when steps == 1)
  const T *grad_states; // reverse: seeds, [B][H][W] (GRAD_LAST) or [steps][B][H][W] (GRAD_ALL)
  int grad_all;
  T *grad_u0;        // reverse: [B][H][W] or NULL
  T *R;              // reverse: tap correlations [B][K*K][H][W] (workspace)
  int B, H, W, N, steps, cs; // cs = CTAs per cluster (= per trajectory)
  int hb_max;        // max rows per band = ceil(H / cs)
  T dt;
  // PDE-Net 1 SingleNonLinear2d (row f-2): constant coefficients coe[N] instead of fields, and a pointwise source
  // term dt * f(u), f a piecewise Lagrange polynomial (fcoef[f_mesh * f_deg + 1] node values on [f_lo, f_lo + f_mesh/f_inv_h])
  int coe_const;
  const T *fcoef;    // NULL = no source term
  T f_lo, f_inv_h;
  int f_mesh, f_deg;
  T *lam_all;        // reverse, optional: dt * lam_{t+1} of every step, [steps][B][H][W] (feeds the node-gradient reduction)
};

// value and derivative (w.r.t. x) of the piecewise Lagrange interpolant, run-time degree 1..4
template <typename T>
__device__ __forceinline__ void lagrange_eval_rt(T x, const T *coef, T lo, T inv_h, int mesh, int deg, T &f, T &df) {
  const T sx = (x - lo) * inv_h;
  int c = (int)floor((double)sx);
  c = c < 0 ? 0 : (c >= mesh ? mesh - 1 : c);
  const T t = sx - (T)c;
  f = T(0);
  df = T(0);
  for (int j = 0; j <= deg; ++j) {
    T v = T(1), dv = T(0);
    for (int m = 0; m <= deg; ++m) {
      if (m == j) continue;
      const T inv = T(deg) / T(j - m), fac = (t - T(m) / T(deg)) * inv;
      dv = dv * fac + v * inv;
      v *= fac;
    }
    const T cj = coef[c * deg + j];
    f = fma_(cj, v, f);
    df = fma_(cj, dv, df);
  }
  df *= inv_h;
}

// rows [y0, y1) of cluster rank c
__device__ __forceinline__ int band_start(int c, int H, int cs) { return (int)(((long long)c * H) / cs); }

// shared-memory layout (elements of T): taps [(N+1)][KK] | w [(hb_max+2r)][KK][W] | buf0, buf1 [(hb_max+2r)][W]
// (state / adjoint, double-buffered) | reverse only: ubuf0, ubuf1 [(hb_max+2r)][W] (u_t, double-buffered: the copy of
// u_{t-1} runs behind the arithmetic of step t)
template <typename T, int K> struct LinClusterSmem {
  static constexpr int R = (K - 1) / 2, KK = K * K;
  static __host__ __device__ size_t elems(int N, int hb_max, int W) {
    return (size_t)(N + 1) * KK + (size_t)(hb_max + 2 * R) * KK * W + 4 * (size_t)(hb_max + 2 * R) * W;
  }
};

// w[row][ab][x] for band rows -R .. hb+R-1 (global rows wrapped): the combined stencil of one step
template <typename T, int K>
__device__ __forceinline__ void build_combined_stencil(const LinClusterParams<T> &p, const T *taps, T *w, int y0, int hb) {
  constexpr int R = (K - 1) / 2, KK = K * K;
  const int W = p.W, H = p.H, N = p.N;
  const int rows = hb + 2 * R;
  for (int idx = threadIdx.x; idx < rows * W; idx += blockDim.x) {
    const int row = idx / W, x = idx - row * W;
    const int g = wrap_index(y0 - R + row, H);
    T acc[KK];
#pragma unroll
    for (int e = 0; e < KK; ++e) acc[e] = T(0);
    // all coefficient loads of the point first (independent, in flight together), then the multiply-adds in filter
    // order: a run-time loop with the load at its top paid one global-memory round trip per filter and point -- ~40 of
    // the 80 us of a BASELINE config-2 forward launch were this prologue
    T cv[kMaxFilters];
#pragma unroll
    for (int n = 0; n < kMaxFilters; ++n)
      cv[n] = n < N ? (p.coe_const ? p.coe[n] : p.coe[((size_t)n * H + g) * W + x]) : T(0);
#pragma unroll
    for (int n = 0; n < kMaxFilters; ++n) {
      if (n < N) {
        const T *kn = taps + (n + 1) * KK;
#pragma unroll
        for (int e = 0; e < KK; ++e) acc[e] = fma_(cv[n], kn[e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < KK; ++e) w[((size_t)row * KK + e) * W + x] = fma_(p.dt, acc[e], taps[e]);
  }
}

// copy the R halo rows above / below the band from the neighbouring CTAs' buffers (DSMEM reads)
template <typename T, int K>
__device__ __forceinline__ void fetch_halo(cg::cluster_group &cluster, T *buf, int hb, int W, int H, int cs, int rank) {
  constexpr int R = (K - 1) / 2;
  const int up = (rank + cs - 1) % cs, dn = (rank + 1) % cs;
  const int hb_up = band_start(up + 1, H, cs) - band_start(up, H, cs);
  const T *rup = cluster.map_shared_rank(buf, up);
  const T *rdn = cluster.map_shared_rank(buf, dn);
  for (int idx = threadIdx.x; idx < R * W; idx += blockDim.x) {
    const int j = idx / W, x = idx - j * W;
    buf[j * W + x] = rup[(R + hb_up - R + j) * W + x];       // last R interior rows of `up`
    buf[(R + hb + j) * W + x] = rdn[(R + j) * W + x];         // first R interior rows of `down`
  }
}

// ---- forward: all steps of one trajectory inside one cluster ---------------------------------------------
template <typename T, int K, int PTS>
__global__ void __launch_bounds__(kLinClusterThreads, 1) linear_cluster_fwd_kernel(const LinClusterParams<T> p) {
  constexpr int R = (K - 1) / 2, KK = K * K;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), cs = p.cs, b = blockIdx.x / cs;
  const int W = p.W, H = p.H;
  const int y0 = band_start(rank, H, cs), hb = band_start(rank + 1, H, cs) - y0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *taps = reinterpret_cast<T *>(smem_raw);
  T *w = taps + (p.N + 1) * KK;
  T *buf0 = w + (size_t)(p.hb_max + 2 * R) * KK * W;
  T *buf1 = buf0 + (size_t)(p.hb_max + 2 * R) * W;
  const int tid = threadIdx.x;
  for (int i = tid; i < (p.N + 1) * KK; i += blockDim.x) taps[i] = i < KK ? p.k0[i] : p.kernels[i - KK];
  __syncthreads();
  build_combined_stencil<T, K>(p, taps, w, y0, hb);
  // initial state: interior rows, then the halo from the neighbours
  const T *u0 = p.u0 + (size_t)b * H * W;
  for (int idx = tid; idx < hb * W; idx += blockDim.x) buf0[R * W + idx] = u0[(size_t)y0 * W + idx];
  cluster.sync();
  fetch_halo<T, K>(cluster, buf0, hb, W, H, cs, rank);
  __syncthreads();

  // this thread's points (band-local linear index), wrapped column of every tap
  int py[PTS], px[PTS], xm[PTS][K];
  bool ok[PTS];
#pragma unroll
  for (int q = 0; q < PTS; ++q) {
    const int idx = tid + q * kLinClusterThreads;
    ok[q] = idx < hb * W;
    py[q] = ok[q] ? idx / W : 0;
    px[q] = ok[q] ? idx - py[q] * W : 0;
#pragma unroll
    for (int bb = 0; bb < K; ++bb) xm[q][bb] = wrap_index(px[q] + bb - R, W);
  }
  // Halo rows of the NEXT state are pushed: the thread that computes a point of the first / last R rows of the band
  // also stores it into the ring neighbour's halo row of the buffer being written (DSMEM store), so ONE cluster
  // barrier separates two steps (the first version pulled the halos after the barrier and needed a second, block-wide
  // barrier behind the pull: two barriers + a remote-load round trip per step on a path that is all latency).
  // Safe: the neighbour's `nxt` was its `cur` one step ago, and it finished reading that before the last barrier.
  const int up = (rank + cs - 1) % cs, dn = (rank + 1) % cs;
  const int hb_up = band_start(up + 1, H, cs) - band_start(up, H, cs);
  T *cur = buf0, *nxt = buf1;
  T *nxt_up = cluster.map_shared_rank(buf1, up), *cur_up = cluster.map_shared_rank(buf0, up);
  T *nxt_dn = cluster.map_shared_rank(buf1, dn), *cur_dn = cluster.map_shared_rank(buf0, dn);
  const size_t stride = (size_t)p.B * H * W;
  for (int t = 0; t < p.steps; ++t) {
    T *out = p.states + (size_t)t * stride + (size_t)b * H * W + (size_t)y0 * W;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
      if (!ok[q]) continue;
      const T *wp = w + ((size_t)(py[q] + R) * KK) * W + px[q];
      T acc = T(0);
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const T *ur = cur + (py[q] + a) * W;
#pragma unroll
        for (int bb = 0; bb < K; ++bb) acc = fma_(wp[(size_t)(a * K + bb) * W], ur[xm[q][bb]], acc);
      }
      if (p.fcoef) {
        T fv, dfv;
        lagrange_eval_rt(cur[(R + py[q]) * W + px[q]], p.fcoef, p.f_lo, p.f_inv_h, p.f_mesh, p.f_deg, fv, dfv);
        acc = fma_(p.dt, fv, acc);
      }
      nxt[(R + py[q]) * W + px[q]] = acc;
      if (py[q] < R) nxt_up[(R + hb_up + py[q]) * W + px[q]] = acc;          // bottom halo of the band above
      if (py[q] >= hb - R) nxt_dn[(py[q] - (hb - R)) * W + px[q]] = acc;     // top halo of the band below
      out[py[q] * W + px[q]] = acc;
    }
    // interiors and halos of every `nxt` are written (release / acquire over the cluster); nobody reads `cur` any more.
    // Also the last step: no CTA may exit while a neighbour can still store into its shared memory.
    cluster.sync();
    T *tmp = cur;
    cur = nxt;
    nxt = tmp;
    tmp = cur_up, cur_up = nxt_up, nxt_up = tmp;
    tmp = cur_dn, cur_dn = nxt_dn, nxt_dn = tmp;
  }
}

// ---- reverse: all steps of one trajectory inside one cluster ---------------------------------------------
template <typename T, int K, int PTS>
__global__ void __launch_bounds__(kLinClusterThreads, 1) linear_cluster_bwd_kernel(const LinClusterParams<T> p) {
  constexpr int R = (K - 1) / 2, KK = K * K;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), cs = p.cs, b = blockIdx.x / cs;
  const int W = p.W, H = p.H;
  const int y0 = band_start(rank, H, cs), hb = band_start(rank + 1, H, cs) - y0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *taps = reinterpret_cast<T *>(smem_raw);
  T *w = taps + (p.N + 1) * KK;
  T *lam0 = w + (size_t)(p.hb_max + 2 * R) * KK * W;
  T *lam1 = lam0 + (size_t)(p.hb_max + 2 * R) * W;
  T *ucur = lam1 + (size_t)(p.hb_max + 2 * R) * W;
  T *unxt = ucur + (size_t)(p.hb_max + 2 * R) * W;
  const int tid = threadIdx.x;
  const size_t img = (size_t)H * W, stride = (size_t)p.B * img;
  // u_t (band + halo rows, wrapped) from global memory, asynchronously (cp.async: the copy of u_{t-1} is issued before
  // the arithmetic of step t and awaited behind it -- the first version loaded u_t synchronously at the top of a step)
  auto stage_u = [&](int t, T *dst) {
    const T *ut = (t == 0 ? p.u0 : p.states + (size_t)(t - 1) * stride) + (size_t)b * img;
    for (int idx = tid; idx < (hb + 2 * R) * W; idx += blockDim.x) {
      const int row = idx / W, x = idx - row * W;
      cp_async_elem(dst + idx, ut + (size_t)wrap_index(y0 - R + row, H) * W + x);
    }
    cp_async_commit();
  };
  stage_u(p.steps - 1, ucur);
  for (int i = tid; i < (p.N + 1) * KK; i += blockDim.x) taps[i] = i < KK ? p.k0[i] : p.kernels[i - KK];
  __syncthreads();
  build_combined_stencil<T, K>(p, taps, w, y0, hb);
  // adjoint of the last state = its seed
  const T *seed_last = p.grad_states + (p.grad_all ? (size_t)(p.steps - 1) * stride : 0) + (size_t)b * img;
  for (int idx = tid; idx < hb * W; idx += blockDim.x) lam0[R * W + idx] = seed_last[(size_t)y0 * W + idx];
  cluster.sync();
  fetch_halo<T, K>(cluster, lam0, hb, W, H, cs, rank);
  cp_async_wait<0>();
  __syncthreads();

  int py[PTS], px[PTS], xm[PTS][K];
  bool ok[PTS];
  T Racc[PTS][KK];
#pragma unroll
  for (int q = 0; q < PTS; ++q) {
    const int idx = tid + q * kLinClusterThreads;
    ok[q] = idx < hb * W;
    py[q] = ok[q] ? idx / W : 0;
    px[q] = ok[q] ? idx - py[q] * W : 0;
#pragma unroll
    for (int bb = 0; bb < K; ++bb) xm[q][bb] = wrap_index(px[q] + bb - R, W);
#pragma unroll
    for (int e = 0; e < KK; ++e) Racc[q][e] = T(0);
  }
  // adjoint halos are pushed into the ring neighbours' buffers as in the forward kernel: one cluster barrier per step
  const int up = (rank + cs - 1) % cs, dn = (rank + 1) % cs;
  const int hb_up = band_start(up + 1, H, cs) - band_start(up, H, cs);
  T *cur = lam0, *nxt = lam1;
  T *nxt_up = cluster.map_shared_rank(lam1, up), *cur_up = cluster.map_shared_rank(lam0, up);
  T *nxt_dn = cluster.map_shared_rank(lam1, dn), *cur_dn = cluster.map_shared_rank(lam0, dn);
  for (int t = p.steps - 1; t >= 0; --t) {
    if (t > 0) stage_u(t - 1, unxt); // `unxt` was read last during step t + 1, before the barrier that ended it
    const T *ubuf = ucur;
    const T *inj = (p.grad_all && t > 0) ? p.grad_states + (size_t)(t - 1) * stride + (size_t)b * img + (size_t)y0 * W
                                         : nullptr;
    T *gu = (t == 0 && p.grad_u0) ? p.grad_u0 + (size_t)b * img + (size_t)y0 * W : nullptr;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
      if (!ok[q]) continue;
      const T lp = cur[(R + py[q]) * W + px[q]];
      T acc = T(0);
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const T *ur = ubuf + (py[q] + a) * W;
        // transposed stencil: source point s = q - (a - R, bb - R): band row py + 2R - a (halo-extended index),
        // column wrapped; the tap that maps s to q is (a, bb), evaluated at s
        const int srow = py[q] + 2 * R - a;
        const T *lr = cur + srow * W;
        const T *wr = w + ((size_t)srow * KK + a * K) * W;
#pragma unroll
        for (int bb = 0; bb < K; ++bb) {
          Racc[q][a * K + bb] = fma_(lp, ur[xm[q][bb]], Racc[q][a * K + bb]);
          const int sx = xm[q][K - 1 - bb]; // px - (bb - R), wrapped
          acc = fma_(wr[(size_t)bb * W + sx], lr[sx], acc);
        }
      }
      if (p.fcoef) { // source term: lam_t += dt f'(u_t) lam_{t+1}; its node gradient is reduced from lam_all afterwards
        T fv, dfv;
        lagrange_eval_rt(ubuf[(R + py[q]) * W + px[q]], p.fcoef, p.f_lo, p.f_inv_h, p.f_mesh, p.f_deg, fv, dfv);
        acc = fma_(p.dt * dfv, lp, acc);
        if (p.lam_all) p.lam_all[(size_t)t * stride + (size_t)b * img + (size_t)(y0 + py[q]) * W + px[q]] = p.dt * lp;
      }
      if (inj) acc += inj[py[q] * W + px[q]];
      nxt[(R + py[q]) * W + px[q]] = acc;
      if (t > 0) {
        if (py[q] < R) nxt_up[(R + hb_up + py[q]) * W + px[q]] = acc;
        if (py[q] >= hb - R) nxt_dn[(py[q] - (hb - R)) * W + px[q]] = acc;
      }
      if (gu) gu[py[q] * W + px[q]] = acc;
    }
    cp_async_wait<0>(); // this thread's share of u_{t-1} has landed; the barrier publishes everybody's
    cluster.sync();
    T *tmp = cur;
    cur = nxt;
    nxt = tmp;
    tmp = cur_up, cur_up = nxt_up, nxt_up = tmp;
    tmp = cur_dn, cur_dn = nxt_dn, nxt_dn = tmp;
    tmp = ucur, ucur = unxt, unxt = tmp;
  }
  // tap correlations of this trajectory's band
  T *Rb = p.R + (size_t)b * KK * img + (size_t)y0 * W;
#pragma unroll
  for (int q = 0; q < PTS; ++q)
    if (ok[q]) {
#pragma unroll
      for (int e = 0; e < KK; ++e) Rb[(size_t)e * img + py[q] * W + px[q]] = Racc[q][e];
    }
}

// ---- finishing kernels: parameter gradients from the tap correlations (fixed order: deterministic) ---------
// Rsum[e][p] = sum_b R[b][e][p];  grad_coe[n][p] = dt * sum_e K_n[e] * Rsum[e][p]
// A block owns kDcoePts consecutive points: first one thread per (tap e, point) sums over the batch in order (the B
// loads of a thread are independent, the points of a tap contiguous), the sums stay in shared memory in double;
// then one thread per (field n, point) combines the KK taps in order.  (The first version ran ONE thread per point
// through all KK*B loads and N*KK multiply-adds: 2500 threads on the whole GPU at BASELINE config 2, 69 us -- the
// third-longest kernel of that pass.  Same addends in the same order: bit-identical.)
constexpr int kDcoePts = 16, kDcoeThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kDcoeThreads) lin_cluster_dcoe_kernel(const T *R, int B, int KK, size_t img,
                                                                        const T *kernels, int N, T dt, T *Rsum,
                                                                        T *grad_coe) {
  __shared__ double rs[PDENET_MAX_K * PDENET_MAX_K][kDcoePts];
  __shared__ double kw[PDENET_MAX_FILTERS * PDENET_MAX_K * PDENET_MAX_K];
  const size_t p0 = (size_t)blockIdx.x * kDcoePts;
  if (grad_coe)
    for (int i = threadIdx.x; i < N * KK; i += kDcoeThreads) kw[i] = (double)kernels[i];
  for (int i = threadIdx.x; i < KK * kDcoePts; i += kDcoeThreads) {
    const int e = i / kDcoePts, q = i % kDcoePts;
    const size_t pidx = p0 + q;
    double s = 0.0;
    if (pidx < img) {
      const T *r = R + (size_t)e * img + pidx;
      const size_t bstride = (size_t)KK * img;
#pragma unroll 8
      for (int b = 0; b < B; ++b) s += (double)r[(size_t)b * bstride];
      Rsum[(size_t)e * img + pidx] = (T)s;
    }
    rs[e][q] = s;
  }
  if (!grad_coe) return;
  __syncthreads();
  for (int i = threadIdx.x; i < N * kDcoePts; i += kDcoeThreads) {
    const int n = i / kDcoePts, q = i % kDcoePts;
    const size_t pidx = p0 + q;
    if (pidx >= img) continue;
    double s = 0.0;
    for (int e = 0; e < KK; ++e) s += kw[n * KK + e] * rs[e][q];
    grad_coe[(size_t)n * img + pidx] = (T)((double)dt * s);
  }
}
// constant coefficients: grad_coe[n] = sum_p field_grad[n][p]  (one block per n, fixed order)
template <typename T> __global__ void __launch_bounds__(256) lin_cluster_dcoe_const_kernel(const T *field, size_t img, T *out) {
  const T *r = field + (size_t)blockIdx.x * img;
  double s = 0.0;
  for (size_t i = threadIdx.x; i < img; i += blockDim.x) s += (double)r[i];
  __shared__ double red[8];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    out[blockIdx.x] = (T)tot;
  }
}
// block (n, e): n = 0 -> dK0[e] = sum_p Rsum[e][p];  n >= 1 -> dK_n[e] = dt * sum_p coe[n-1][p] * Rsum[e][p]
template <typename T>
__global__ void __launch_bounds__(256) lin_cluster_dk_kernel(const T *Rsum, const T *coe, int coe_const, size_t img, int KK,
                                                             T dt, T *grad_k0, T *grad_kernels) {
  const int n = blockIdx.y, e = blockIdx.x;
  const T *r = Rsum + (size_t)e * img;
  const T *c = (n > 0 && !coe_const) ? coe + (size_t)(n - 1) * img : nullptr;
  const double cc = (n > 0 && coe_const) ? (double)coe[n - 1] : 1.0;
  double s = 0.0;
  for (size_t i = threadIdx.x; i < img; i += blockDim.x) s += (c ? (double)c[i] : cc) * (double)r[i];
  __shared__ double red[8];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    if (n == 0) {
      if (grad_k0) grad_k0[e] = (T)tot;
    } else if (grad_kernels) {
      grad_kernels[(size_t)(n - 1) * KK + e] = (T)((double)dt * tot);
    }
  }
}

} // namespace pdenet
