// Second reverse kernel ("kernel B"): given the derivative adjoints g written by kernel A,
//   lam_t(q)      = lam_{t+1}(q) [+ seed_t(q)] + sum_n sum_{a,b} K_n[a,b] * g_{c,n}(q - (a-r, b-r))
//   dK_n[a,b]    += sum_{b,c,p} g_{c,n}(p) * u_c(p + (a-r, b-r))
// (data gradient = transposed correlation, filter gradient =
// global reduction).  One work item = (trajectory, channel, 64-row tile); the N filters are
// streamed through a double-buffered cp.async pipeline (g_n tile + halo), the u tile is staged once
// per item.  Thread micro-tile: RW=8 rows x VEC columns (16 bytes), so every shared-memory read
// is a conflict-free LDS.128 and each loaded value feeds up to 7*VEC..7*7 FMAs from registers:
//   B1 (data gradient):   49 flipped weights in registers, RW*VEC accumulators
//   B2 (filter gradient): RW*VEC g values in registers, 49 accumulators, then one recursive-halving
//                         warp reduction per (item, filter) into a per-warp global partial slot.
#pragma once
#include "common.cuh"

namespace pdenet {

template <typename T> struct AdjParams {
  const T *g; // [B][C*N][g_rows][W]
  int g_rows, g_off;
  const T *u; // [B][C][u_rows][W]
  int u_rows, u_off;
  int wrap_y;
  const T *lam_in; // [B][C][li_rows][W]; NULL = 0
  int li_rows, li_off;
  const T *inject; // [B][C][inj_rows][W]; NULL = none
  int inj_rows, inj_off;
  T *lam_out; // [B][C][lo_rows][W]; NULL = skip the data gradient
  int lo_rows, lo_off;
  const T *kernels;   // [N][K][K]
  const T *kernels0;  // if non-NULL: filter 0 is kernels0 [K][K] and `kernels` holds filters 1..N-1 (linear model)
  double *dk_partial; // [gridDim.x * 8][N][VLB] accumulated; NULL = skip the filter gradient
  int B, C, N, H, W;
  int tiles_x, tiles_y, n_items;
};

template <typename T, int K> struct AdjCfg {
  static constexpr int R = (K - 1) / 2;
  static constexpr int NT = 256;
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int RW = 8;
  static constexpr int TW = 32 * VEC, TH = RW * (NT / 32);
  static constexpr int NV = (VEC + K - 1 + VEC - 1) / VEC; // 16-byte vectors per window row
  static constexpr int SR = TH + K - 1, SC = TW + K - 1;
  static constexpr int PITCH =
      round_up(SC, VEC) > (31 * VEC + NV * VEC) ? round_up(SC, VEC) : (31 * VEC + NV * VEC);
  static constexpr int KK = K * K;
  static constexpr int KKP = round_up(KK, VEC);
  static constexpr int VLB = KK <= 16 ? 16 : (KK <= 32 ? 32 : 64);
  static constexpr int TILE_ELEMS = SR * PITCH;
  static __host__ size_t smem_bytes(int N) { return (size_t)(N * KKP + 4 * TILE_ELEMS) * sizeof(T); }
};

template <typename T, int K> __global__ void __launch_bounds__(256, 1) adjoint_kernel(const AdjParams<T> p) {
  using Cfg = AdjCfg<T, K>;
  constexpr int R = Cfg::R, NT = Cfg::NT, VEC = Cfg::VEC, RW = Cfg::RW, TW = Cfg::TW, TH = Cfg::TH;
  constexpr int NV = Cfg::NV, SR = Cfg::SR, SC = Cfg::SC, PITCH = Cfg::PITCH, KK = Cfg::KK, KKP = Cfg::KKP;
  constexpr int VLB = Cfg::VLB, TILE = Cfg::TILE_ELEMS;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *wflip = reinterpret_cast<T *>(smem_raw); // [N][KKP]: wflip[n][a'*K+b'] = K_n[K-1-a'][K-1-b']
  T *ubuf = wflip + p.N * KKP;                // [2][SR][PITCH]
  T *gbuf = ubuf + 2 * TILE;                  // [2][SR][PITCH]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, N = p.N, C = p.C;
  const bool do_lam = p.lam_out != nullptr, do_dk = p.dk_partial != nullptr;

  for (int idx = tid; idx < N * KKP; idx += NT) {
    const int n = idx / KKP, e = idx % KKP;
    T w = T(0);
    if (e < KK) {
      if (p.kernels0)
        w = n == 0 ? p.kernels0[KK - 1 - e] : p.kernels[(n - 1) * KK + (KK - 1 - e)];
      else
        w = p.kernels[n * KK + (KK - 1 - e)];
    }
    wflip[idx] = w;
  }

  auto decode = [&](int item, int &b, int &c, int &ty, int &tx) {
    tx = item % p.tiles_x;
    int t = item / p.tiles_x;
    ty = t % p.tiles_y;
    t /= p.tiles_y;
    c = t % C;
    b = t / C;
  };
  auto stage_u = [&](int item, int ub) {
    int b, c, ty, tx;
    decode(item, b, c, ty, tx);
    stage_tile_async<T, SR, SC, PITCH, NT>(ubuf + ub * TILE, p.u + (size_t)(b * C + c) * p.u_rows * W, p.u_rows,
                                           p.u_off, p.wrap_y, ty * TH - R, tx * TW - R, H, W);
  };
  auto stage_g = [&](int item, int n, int gb) {
    int b, c, ty, tx;
    decode(item, b, c, ty, tx);
    stage_tile_async<T, SR, SC, PITCH, NT>(gbuf + gb * TILE, p.g + (size_t)((b * C + c) * N + n) * p.g_rows * W,
                                           p.g_rows, p.g_off, p.wrap_y, ty * TH - R, tx * TW - R, H, W);
  };

  int item = blockIdx.x, ub = 0, gb = 0;
  if (item < p.n_items) {
    if (do_dk) stage_u(item, 0);
    stage_g(item, 0, 0);
  }
  cp_async_commit();

  const int lrow = warp * RW, lcol = lane * VEC;

  for (; item < p.n_items; item += gridDim.x) {
    int b, c, ty, tx;
    decode(item, b, c, ty, tx);
    const int y0 = ty * TH + lrow, x0 = tx * TW + lcol;

    T lacc[RW][VEC];
#pragma unroll
    for (int j = 0; j < RW; ++j)
#pragma unroll
      for (int i = 0; i < VEC; ++i) lacc[j][i] = T(0);

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      // ---- prefetch the next stage (next filter of this item, or the first of the next item)
      if (n + 1 < N) {
        stage_g(item, n + 1, gb ^ 1);
      } else if (item + (int)gridDim.x < p.n_items) {
        if (do_dk) stage_u(item + gridDim.x, ub ^ 1);
        stage_g(item + gridDim.x, 0, gb ^ 1);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();

      const T *gs = gbuf + gb * TILE + lrow * PITCH + lcol;
      const T *us = ubuf + ub * TILE + lrow * PITCH + lcol;
      T gc[RW][VEC];

      // ---- B1: data gradient (correlation of g_n with the flipped filter)
      {
        T wreg[KKP];
#pragma unroll
        for (int q = 0; q < KKP / VEC; ++q)
          ld_vec(wflip + n * KKP + q * VEC, &wreg[q * VEC]);
#pragma unroll
        for (int rho = 0; rho < RW + K - 1; ++rho) {
          T row[NV * VEC];
#pragma unroll
          for (int v = 0; v < NV; ++v)
            ld_vec(gs + rho * PITCH + v * VEC, &row[v * VEC]);
          if (rho >= R && rho < R + RW) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) gc[rho - R][i] = row[i + R];
          }
          if (do_lam) {
#pragma unroll
            for (int a = 0; a < K; ++a) {
              const int j = rho - a;
              if (j >= 0 && j < RW) {
#pragma unroll
                for (int i = 0; i < VEC; ++i)
#pragma unroll
                  for (int bb = 0; bb < K; ++bb) lacc[j][i] = fma_(wreg[a * K + bb], row[i + bb], lacc[j][i]);
              }
            }
          }
        }
      }

      // ---- B2: filter gradient
      if (do_dk) {
#pragma unroll
        for (int j = 0; j < RW; ++j)
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if (!((y0 + j < H) && (x0 + i < W))) gc[j][i] = T(0);
        T kacc[VLB];
#pragma unroll
        for (int e = 0; e < VLB; ++e) kacc[e] = T(0);
#pragma unroll
        for (int rho = 0; rho < RW + K - 1; ++rho) {
          T row[NV * VEC];
#pragma unroll
          for (int v = 0; v < NV; ++v)
            ld_vec(us + rho * PITCH + v * VEC, &row[v * VEC]);
#pragma unroll
          for (int a = 0; a < K; ++a) {
            const int j = rho - a;
            if (j >= 0 && j < RW) {
#pragma unroll
              for (int i = 0; i < VEC; ++i)
#pragma unroll
                for (int bb = 0; bb < K; ++bb) kacc[a * K + bb] = fma_(gc[j][i], row[i + bb], kacc[a * K + bb]);
            }
          }
        }
        warp_reduce_vec<T, VLB>(kacc);
        double *slot = p.dk_partial + ((size_t)(blockIdx.x * (NT / 32) + warp) * N + n) * VLB;
        if constexpr (VLB >= 32) {
#pragma unroll
          for (int i = 0; i < VLB / 32; ++i) {
            const int e = lane * (VLB / 32) + i;
            if (e < KK) slot[e] += (double)kacc[i];
          }
        } else {
          constexpr int DUP = 32 / VLB;
          const int e = lane / DUP;
          if ((lane % DUP) == 0 && e < KK) slot[e] += (double)kacc[0];
        }
      }
      __syncthreads(); // stage buffers gb (and, on the last filter, ub) are free again
      gb ^= 1;
    }
    ub ^= 1;

    // ---- item epilogue: lam_out = lam_in + seed + correlation sum
    if (do_lam) {
#pragma unroll
      for (int j = 0; j < RW; ++j) {
        const int y = y0 + j;
        if (y < H) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const int x = x0 + i;
            if (x < W) {
              T v = lacc[j][i];
              if (p.lam_in) v += p.lam_in[((size_t)(b * C + c) * p.li_rows + (y + p.li_off)) * W + x];
              if (p.inject) v += p.inject[((size_t)(b * C + c) * p.inj_rows + (y + p.inj_off)) * W + x];
              p.lam_out[((size_t)(b * C + c) * p.lo_rows + (y + p.lo_off)) * W + x] = v;
            }
          }
        }
      }
    }
  }
  cp_async_wait<0>();
}

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
