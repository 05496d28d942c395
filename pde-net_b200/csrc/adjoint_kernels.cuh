// Reverse kernels "B1" and "B2": given the derivative adjoints g written by kernel A,
//   B1 (adjoint_kernel):      lam_t(q) = lam_{t+1}(q) [+ seed_t(q)]
//                                        + sum_n sum_{a,b} K_n[a,b] * g_{c,n}(q - (a-r, b-r))
//   B2 (filter_grad_kernel):  dK_n[a,b] += sum_{b,c,p} g_{c,n}(p) * u_c(p + (a-r, b-r))
// (data gradient = transposed correlation, filter gradient =
// global reduction).  Thread micro-tile in both: RW=8 rows x VEC columns (16 bytes), so every
// shared-memory read is a conflict-free LDS.128 and each loaded value feeds up to 7*VEC..7*7 FMAs
// from registers.
//   B1: one work item = (trajectory, channel, tile); the N filters are streamed through a
//       double-buffered cp.async pipeline (g_n tile + halo); 49 flipped weights in registers,
//       RW*VEC accumulators per thread.
//   B2: warp w of a CTA owns filter n0+w for the whole launch and keeps its 49 accumulators in
//       registers across every tile the CTA visits (the u tile + halo is staged once per item and
//       shared by all warps; g is read straight from global memory, no halo needed), so the
//       cross-lane reduction happens once per launch instead of once per (tile, filter).
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace pdenet {

// NT_ = 256: 64-row tiles, two CTAs per SM.  NT_ = 64: the SMALL tile (16 rows, eight CTAs per SM) for problems
// with too few large tiles to fill / balance the SMs.
template <typename T, int K, int NT_ = 256> struct AdjCfg {
  static constexpr int R = (K - 1) / 2;
  static constexpr int NT = NT_;
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int RW = 8;
  static constexpr int TW = 32 * VEC, TH = RW * (NT / 32);
  // With 16-byte staging the tile is shifted left by XS columns so that the window start
  // (tile origin - R) is a multiple of VEC in global memory as well as in shared memory.
  static constexpr int XS = (VEC - R % VEC) % VEC;
  static constexpr int NV = (VEC + K - 1 + VEC - 1) / VEC; // 16-byte vectors per window row
  static constexpr int SR = TH + K - 1, SC = TW + K - 1;
  static constexpr int PITCH =
      round_up(SC, VEC) > (31 * VEC + NV * VEC) ? round_up(SC, VEC) : (31 * VEC + NV * VEC);
  static constexpr int KK = K * K;
  static constexpr int KKP = round_up(KK, VEC);
  static constexpr int VLB = KK <= 16 ? 16 : (KK <= 32 ? 32 : 64);
  static constexpr int TILE_ELEMS = SR * PITCH;
  static __host__ size_t smem_bytes(int) { return (size_t)(2 * TILE_ELEMS) * sizeof(T); }
};

template <typename T, int K, int NT_ = 256>
__global__ void __launch_bounds__(NT_, NT_ == 512 ? 1 : (NT_ == 256 ? 2 : 8)) adjoint_kernel(const AdjParams<T> p) {
  using Cfg = AdjCfg<T, K, NT_>;
  constexpr int R = Cfg::R, NT = Cfg::NT, VEC = Cfg::VEC, RW = Cfg::RW, TW = Cfg::TW, TH = Cfg::TH;
  constexpr int NV = Cfg::NV, SR = Cfg::SR, SC = Cfg::SC, PITCH = Cfg::PITCH, KK = Cfg::KK;
  constexpr int TILE = Cfg::TILE_ELEMS;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *gbuf = reinterpret_cast<T *>(smem_raw); // [2][SR][PITCH]
  const T *cw = const_taps<T>(p.slot);       // [N][K][K] taps in constant memory (uniform operands)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, N = p.N, C = p.C;

  auto decode = [&](int item, int &b, int &c, int &ty, int &tx) {
    tx = item % p.tiles_x;
    int t = item / p.tiles_x;
    ty = t % p.tiles_y;
    t /= p.tiles_y;
    c = t % C;
    b = t / C;
  };
  const int xs = p.vec_ok ? Cfg::XS : 0; // periodic in x: a shifted tiling covers the same points
  const size_t plane = (size_t)p.g_rows * W;
  // Staging plan (16-byte path): the global offsets of this thread's units of the tile are
  // computed once per item (periodic wrap, clamping) and reused for all N filters of the item --
  // per stage only an add and a cp.async per unit remain.  Unit q = tid + k*NT -> row q / UNITS.
  constexpr int UNITS = (SC + VEC - 1) / VEC, NQ = SR * UNITS, NU = (NQ + NT - 1) / NT;
  int goff[NU];
  auto wrap_near = [](int v, int n) {
    if (v < 0)
      v += n;
    else if (v >= n)
      v -= n;
    return ((unsigned)v < (unsigned)n) ? v : wrap_index(v, n);
  };
  auto make_plan = [&](int ty_, int tx_) {
    const int ybase = ty_ * TH - R, xbase = tx_ * TW - xs - R;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int q = tid + k * NT, row = q / UNITS, unit = q - row * UNITS;
      int gy = ybase + row;
      if (p.wrap_y) gy = wrap_near(gy, H);
      gy += p.g_off;
      gy = gy < 0 ? 0 : (gy >= p.g_rows ? p.g_rows - 1 : gy);
      goff[k] = gy * W + wrap_near(xbase + unit * VEC, W);
    }
  };
  auto stage_plan = [&](const T *img, int gb) {
    T *dst = gbuf + gb * TILE;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int q = tid + k * NT, row = q / UNITS, unit = q - row * UNITS;
      if (q < NQ) cp_async_16_zfill(dst + row * PITCH + unit * VEC, img + goff[k], true);
    }
  };
  // stage filter n of the item whose first g plane is `img0` and whose tile is (ty, tx); on the
  // 16-byte path the plan must describe that tile
  auto stage_g = [&](const T *img0, int ty, int tx, int n, int gb) {
    const T *img = img0 + (size_t)n * plane;
    if (p.vec_ok)
      stage_plan(img, gb);
    else
      stage_tile_async<T, SR, SC, PITCH, NT>(gbuf + gb * TILE, img, p.g_rows, p.g_off, p.wrap_y, ty * TH - R,
                                             tx * TW - R, H, W);
  };

  int item = blockIdx.x, gb = 0;
  int b = 0, c = 0, ty = 0, tx = 0;     // current item
  int nb_ = 0, nc_ = 0, nty = 0, ntx = 0; // next item (decoded once, when its first tile is prefetched)
  if (item < p.n_items) {
    decode(item, b, c, ty, tx);
    if (p.vec_ok) make_plan(ty, tx);
    stage_g(p.g + (size_t)(b * C + c) * N * plane, ty, tx, 0, 0);
  }
  cp_async_commit();

  const int lrow = warp * RW, lcol = lane * VEC;

  for (; item < p.n_items; item += gridDim.x) {
    const int y0 = ty * TH + lrow, x0 = tx * TW - xs + lcol; // x0 in [-xs, W - xs)
    const T *img0 = p.g + (size_t)(b * C + c) * N * plane;

    T lacc[RW][VEC];
#pragma unroll
    for (int j = 0; j < RW; ++j)
#pragma unroll
      for (int i = 0; i < VEC; ++i) lacc[j][i] = T(0);

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      // One barrier per stage: after it, this stage's tile has landed for every thread AND every
      // warp has finished the previous stage, so the other buffer is free for the next prefetch
      // (next filter of this item, or the first filter of the next item).
      cp_async_wait<0>();
      __syncthreads();
      if (n + 1 < N) {
        stage_g(img0, ty, tx, n + 1, gb ^ 1);
      } else if (item + (int)gridDim.x < p.n_items) {
        decode(item + gridDim.x, nb_, nc_, nty, ntx);
        if (p.vec_ok) make_plan(nty, ntx); // every tile of the current item has been issued
        stage_g(p.g + (size_t)(nb_ * C + nc_) * N * plane, nty, ntx, 0, gb ^ 1);
      }
      cp_async_commit();

      const T *gs = gbuf + gb * TILE + lrow * PITCH + lcol;

      // ---- B1: data gradient (correlation of g_n with the flipped filter)
      {
        const T *wn0 = cw + n * KK; // this filter's taps; flipped index below = transposed correlation
        // window rows in pairs: each tap fetched into a uniform register feeds 2*VEC FMAs
        static_assert((RW + K - 1) % 2 == 0, "window rows are processed in pairs");
#pragma unroll
        for (int rp = 0; rp < (RW + K - 1) / 2; ++rp) {
          const T *wn = wn0 + p.zz[rp];
          T row[2][NV * VEC];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int v = 0; v < NV; ++v) ld_vec(gs + (2 * rp + h) * PITCH + v * VEC, &row[h][v * VEC]);
          // tap column outermost: consecutive FMAs hit different accumulators (no dependent chains)
#pragma unroll
          for (int bb = 0; bb < K; ++bb)
#pragma unroll
            for (int a = 0; a < K; ++a) {
              const T w = wn[KK - 1 - (a * K + bb)];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int j = 2 * rp + h - a;
                if (j >= 0 && j < RW) {
#pragma unroll
                  for (int i = 0; i < VEC; ++i) lacc[j][i] = fma_(row[h][i + bb], w, lacc[j][i]);
                }
              }
            }
        }
      }

      gb ^= 1;
    }

    // ---- item epilogue: lam_out = lam_in + seed + correlation sum.  Loads are issued for a
    // whole row group before any use (clamped addresses instead of branches) so they overlap.
    {
      const size_t img = (size_t)(b * C + c);
      const T *li = p.lam_in ? p.lam_in + (img * p.li_rows + p.li_off) * W : nullptr;
      const T *inj = p.inject ? p.inject + (img * p.inj_rows + p.inj_off) * W : nullptr;
      T *lo = p.lam_out + (img * p.lo_rows + p.lo_off) * W;
      int xx[VEC];
      bool okx[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int xg = x0 + i;
        okx[i] = xg < W - xs;
        xx[i] = okx[i] ? (xg < 0 ? xg + W : xg) : 0;
      }
#pragma unroll
      for (int j0 = 0; j0 < RW; j0 += 4) {
        T add[4][VEC];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int y = y0 + j0 + jj;
          const size_t ro = (size_t)(y < H ? y : 0) * W;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            T v = T(0);
            if (li) v = li[ro + xx[i]];
            if (inj) v += inj[ro + xx[i]];
            add[jj][i] = v;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int y = y0 + j0 + jj;
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if (y < H && okx[i]) {
              const T v = lacc[j0 + jj][i] + add[jj][i];
              lo[(size_t)y * W + xx[i]] = v;
              if (p.mirror_rows) { // fused adjoint halo exchange: peer stores over NVLink
                const size_t rb = img * p.lo_rows + p.lo_off;
                if (y < p.mirror_rows) p.peer_up[(rb + H + y) * W + xx[i]] = v;
                if (y >= H - p.mirror_rows) p.peer_dn[(rb + y - H) * W + xx[i]] = v;
              }
            }
        }
      }
    }
    b = nb_;
    c = nc_;
    ty = nty;
    tx = ntx;
  }
  cp_async_wait<0>();
}

// --------------------------------------------------------------------------------------------
// MT_ = 8: 64-row tiles.  MT_ = 2: the SMALL tile (16 rows) for problems with too few large tiles.
template <typename T, int K, int MT_ = 8> struct FGradCfg {
  static constexpr int R = (K - 1) / 2;
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int RW = 8;
  static constexpr int MT = MT_; // micro-tiles (of RW rows) per lane per tile
  static constexpr int TW = 32 * VEC, TH = RW * MT;
  static constexpr int NV = (VEC + K - 1 + VEC - 1) / VEC;
  static constexpr int SR = TH + K - 1, SC = TW + K - 1;
  static constexpr int PITCH =
      round_up(SC, VEC) > (31 * VEC + NV * VEC) ? round_up(SC, VEC) : (31 * VEC + NV * VEC);
  static constexpr int KK = K * K;
  static constexpr int VLB = KK <= 16 ? 16 : (KK <= 32 ? 32 : 64);
  static constexpr int TILE_ELEMS = SR * PITCH;
  static constexpr int NT = 32 * kFGradWarps; // see params.cuh
  static constexpr int GST_ELEMS = RW * 32 * VEC; // per-warp g micro-tile staging (single buffer: it is
                                                  // drained into registers before the next copy is issued)
  static __host__ size_t smem_bytes() { return (size_t)(2 * TILE_ELEMS + (NT / 32) * GST_ELEMS) * sizeof(T); }
};

template <typename T, int K, int MT_ = 8>
__global__ void __launch_bounds__(FGradCfg<T, K, MT_>::NT, kFGradCtasPerSm) filter_grad_kernel(const FGradParams<T> p) {
  using Cfg = FGradCfg<T, K, MT_>;
  constexpr int NTHREADS = Cfg::NT;
  constexpr int R = Cfg::R, VEC = Cfg::VEC, RW = Cfg::RW, MT = Cfg::MT, TW = Cfg::TW, TH = Cfg::TH;
  constexpr int NV = Cfg::NV, SR = Cfg::SR, SC = Cfg::SC, PITCH = Cfg::PITCH, KK = Cfg::KK, VLB = Cfg::VLB;
  constexpr int TILE = Cfg::TILE_ELEMS;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *ubuf = reinterpret_cast<T *>(smem_raw); // [2][SR][PITCH]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T *gst = ubuf + 2 * TILE + warp * Cfg::GST_ELEMS + lane * VEC; // [RW][32 lanes][VEC], this lane's slots
  const int H = p.H, W = p.W, N = p.N, C = p.C;
  const int parts = (NTHREADS / 32) / p.nw, part = warp / p.nw;
  const int n = p.n0 + warp % p.nw;

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
    stage_tile_async<T, SR, SC, PITCH, NTHREADS>(ubuf + ub * TILE, p.u + (size_t)(b * C + c) * p.u_rows * W,
                                                 p.u_rows, p.u_off, p.wrap_y, ty * TH - R, tx * TW - R, H, W);
  };

  T kacc[VLB];
#pragma unroll
  for (int e = 0; e < VLB; ++e) kacc[e] = T(0);
  double *slot = p.dk_partial + ((size_t)(blockIdx.x * p.slot_stride + part) * N + n) * VLB;
  // fp32: the per-lane accumulators are widened once per ITEM (<= MT*RW*VEC products each in single
  // precision), the cross-lane sum and everything after it run in double -- the error of the filter
  // gradient then no longer grows with the grid size or the batch.  fp64 reduces once per launch.
  auto flush_f32 = [&]() {
#pragma unroll
    for (int c0 = 0; c0 < VLB; c0 += 16) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = (double)kacc[c0 + i];
        kacc[c0 + i] = T(0);
      }
      warp_reduce_vec<double, 16>(v); // v[0] = element lane/2, duplicated over two lanes
      const int e = c0 + (lane >> 1);
      if ((lane & 1) == 0 && e < KK) slot[e] += v[0];
    }
  };

  int item = blockIdx.x, ub = 0;
  if (item < p.n_items) stage_u(item, 0);
  cp_async_commit();

  for (; item < p.n_items; item += gridDim.x) {
    // One barrier per item: every cp.async group issued so far (this item's u tile, issued one item
    // ago) has landed, and every warp has finished the previous item, so the other u buffer is
    // free for the next prefetch.
    cp_async_wait<0>();
    __syncthreads();

    {
      int b, c, ty, tx;
      decode(item, b, c, ty, tx);
      const int x0 = tx * TW + lane * VEC;
      const T *gimg = p.g + (size_t)((b * C + c) * N + n) * p.g_rows * W;
      const T *us = ubuf + ub * TILE + lane * VEC;

      // g micro-tile: global -> this warp's staging slots (cp.async, zero-filled outside the
      // domain), one micro-tile ahead of the compute.  The row pointer advances by one 64-bit add per
      // row and the 16-byte / scalar choice is made once per micro-tile (the first version rebuilt a
      // 64-bit address and branched on vec_ok for every row: ~200 integer / branch instructions per
      // 784 FMAs).
      const bool okx = x0 < W;
      const T *gcol = gimg + (size_t)p.g_off * W + (okx ? x0 : 0);
      const size_t wstride = (size_t)W;
      auto stage_g = [&](int mt) {
        const int ybase = ty * TH + mt * RW;
        const T *grow = gcol + (size_t)ybase * wstride;
        const int rows_ok = okx ? H - ybase : 0; // rows j < rows_ok are inside the domain
        if (p.vec_ok) {
#pragma unroll
          for (int j = 0; j < RW; ++j) {
            const bool ok = j < rows_ok;
            cp_async_16_zfill(gst + j * 32 * VEC, ok ? grow : gcol, ok);
            grow += wstride;
          }
        } else {
#pragma unroll
          for (int j = 0; j < RW; ++j) {
            const bool ok = j < rows_ok;
            T *dst = gst + j * 32 * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) dst[i] = (ok && x0 + i < W) ? grow[i] : T(0);
            grow += wstride;
          }
        }
      };

      // cp.async group order per item: G(first), U(next item), then G(mt+1) right after G(mt) has
      // been drained into registers -- the staging slots are single-buffered and lane-private.
      stage_g(part);
      cp_async_commit();
      if (item + (int)gridDim.x < p.n_items) stage_u(item + gridDim.x, ub ^ 1);
      cp_async_commit();
      bool first = true;
#pragma unroll 1
      for (int mt = part; mt < MT; mt += parts) {
        if (ty * TH + mt * RW >= H) break; // warp-uniform: nothing left in this tile
        if (first)
          cp_async_wait<1>(); // G(first) landed; U(next) may still be in flight
        else
          cp_async_wait<0>();
        first = false;
        T gc[RW][VEC];
#pragma unroll
        for (int j = 0; j < RW; ++j) ld_vec(gst + j * 32 * VEC, &gc[j][0]);
        if (mt + parts < MT) stage_g(mt + parts);
        cp_async_commit();
        const T *um = us + mt * RW * PITCH;
#pragma unroll
        for (int rho = 0; rho < RW + K - 1; ++rho) {
          T row[NV * VEC];
#pragma unroll
          for (int v = 0; v < NV; ++v) ld_vec(um + rho * PITCH + v * VEC, &row[v * VEC]);
          // column of the micro-tile outermost: every accumulator is touched once per pass, so
          // dependent FMAs are ~K*K instructions apart
#pragma unroll
          for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int a = 0; a < K; ++a) {
              const int j = rho - a;
              if (j >= 0 && j < RW) {
#pragma unroll
                for (int bb = 0; bb < K; ++bb) kacc[a * K + bb] = fma_(gc[j][i], row[i + bb], kacc[a * K + bb]);
              }
            }
        }
      }
    }
    if constexpr (sizeof(T) == 4) flush_f32();
    ub ^= 1;
  }
  cp_async_wait<0>();

  if constexpr (sizeof(T) == 8) {
    warp_reduce_vec<T, VLB>(kacc);
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
}

} // namespace pdenet
