// Fused per-step kernels: stencil bank -> SymNet -> Euler (forward), and the first reverse kernel
// ("kernel A": recompute stencils + SymNet, back-propagate the SymNet, emit the derivative
// adjoints g and reduce the SymNet parameter gradients).
//
// Restates SURVEY.md section 8(a) rows a5-a9 (forward) and the SymNet half of a11 (reverse).
// Thread mapping: one warp = 32 consecutive x; a thread owns a vertical strip of P points (4 in
// fp64, 8 in fp32) and keeps all N derivative accumulators of that strip in registers, so one
// staged u value feeds N*min(P,K) FMAs; the filter taps come from constant memory through uniform
// registers (`DFMA R, R, UR, R`: two vector-register operands per FMA, the only way to the full
// FP64 rate on B200 -- DESIGN.md section 3) and each feeds P FMAs.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace pdenet {

// per-warp transpose scratch used by the SymNet parameter-gradient reduction (see transpose_sum)
template <typename T> struct TransposeCfg {
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int PITCH = 32 + VEC; // elements per row; keeps rows 16-byte aligned
  static constexpr int ELEMS = 32 * PITCH;
};

#ifndef PDENET_STEP_SPW
#define PDENET_STEP_SPW 2   // strips per warp and large tile
#endif
#ifndef PDENET_STEP_SPW_KEPT
#define PDENET_STEP_SPW_KEPT 1   // ... of kernel A on kept features: it stages no tile (no halo to amortise), and 64 x 16-point
                                 // items measured 4-5 % faster than 64 x 32 (1.595 -> 1.52 ms per 32 trajectories, cfg4 fp64)
#endif
// strips per warp of a mode's large tile (0 = PDENET_STEP_SPW)
__host__ __device__ constexpr int step_spw_override(int mode) { return mode_is_kept(mode) ? PDENET_STEP_SPW_KEPT : 0; }
template <typename T, int K, int N, int C, int NT_ = 256, int PO = 0, int SPWO = 0> struct StepCfg {
  static constexpr int R = (K - 1) / 2;
  static constexpr int NT = NT_;
  // points per thread strip: 4 in fp64; 8 in fp32 (FFMA issues every cycle, so each uniform tap
  // load has to feed twice as many FMAs to stay off the issue-slot limit) when the C*N features of
  // 8 points still fit the register file
  // (PO overrides: see step_p_override)
  static constexpr int P = PO ? PO : ((sizeof(T) == 4 && C * N <= 12) ? 8 : 4);
  // warp layout of a tile: WA warps across (32 columns each), WD down, SPW strips of P rows per warp.
  // NT >= 256: the large tile, 64 x (NT/64 * 2P).  NT == 128: the SMALL tile, 32 x 4P (four warps stacked),
  // for problems whose large-tile count would leave SMs idle or badly balanced (BASELINE configs 1 and 3).
  static constexpr int WA = NT >= 256 ? 2 : 1, WD = NT / 32 / WA, SPW = NT >= 256 ? (SPWO ? SPWO : PDENET_STEP_SPW) : 1;
  static constexpr int TW = 32 * WA, TH = WD * SPW * P;
  static constexpr int SR = TH + K - 1, SC = TW + K - 1;
  static constexpr int PITCH = SC;
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int NP = round_up(N, VEC);
  static constexpr int M = C * N;
  static constexpr int TILE_ELEMS = C * SR * PITCH;
  static __host__ __device__ constexpr int sym_elems(int Psym) { return round_up(C * Psym, VEC); }
  // per-warp SymNet-gradient accumulators (double, first in shared memory so they stay aligned)
  static __host__ __device__ constexpr int grad_elems(int Psym) { return round_up(C * Psym, 2); }
  // MODE_BWD_A_KEPT stages no tiles (the shared memory it leaves free stays L1 cache)
  static constexpr int KEPT_ROWS = (M + C) * P; // rows of 32 columns one warp reads per strip: features + adjoint
  static __host__ __device__ constexpr int tile_region(bool kept) { return kept ? 0 : 2 * TILE_ELEMS; }
  static __host__ size_t smem_bytes(int Psym, bool with_grad, bool kept = false) {
    size_t e = sym_elems(Psym) + tile_region(kept) + (with_grad ? (NT / 32) * TransposeCfg<T>::ELEMS : 0);
    return e * sizeof(T) + (with_grad ? (size_t)(NT / 32) * grad_elems(Psym) * sizeof(double) + 64 * sizeof(int) : 0);
  }
};

// --------------------------------------------------------------------------------------------
// SymNet forward for PP points (row a7).  d: derivative features in registers, prm: this output
// channel's flat parameters in shared memory (broadcast reads).  Layers are looped at run time;
// the few per-layer scalars (p, q, product units) live in small local arrays.
// LT >= 0: the depth is a compile-time constant (layer loops fully unrolled, parameter offsets become
// immediates, the per-layer scalars stay in registers); LT < 0: run-time depth L_rt.
// MD (deduced) >= M: the feature array may carry extra rows after the M SymNet inputs (upwind twins).
template <typename T, int M, int P, int PP, int J0, bool KEEP, int LT, bool NEED_OUT = true, int MD = M>
__device__ __forceinline__ int symnet_forward(const T (&d)[MD][P], const T *prm, int L_rt, T (&out)[PP],
                                              T (*un)[PP], T (*pl)[PP], T (*ql)[PP]) {
  const int L = LT >= 0 ? LT : L_rt;
  int off = 0;
#pragma unroll(LT >= 0 ? 8 : 1)
  for (int i = 0; i < L; ++i) {
    const int mi = M + i;
    T pv[PP], qv[PP];
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) {
      pv[jj] = prm[off + 2 * mi];
      qv[jj] = prm[off + 2 * mi + 1];
    }
#pragma unroll
    for (int f = 0; f < M; ++f) {
      const T w0 = prm[off + f], w1 = prm[off + mi + f];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) {
        pv[jj] = fma_(w0, d[f][J0 + jj], pv[jj]);
        qv[jj] = fma_(w1, d[f][J0 + jj], qv[jj]);
      }
    }
#pragma unroll(LT >= 0 ? 8 : 1)
    for (int i2 = 0; i2 < i; ++i2) {
      const T w0 = prm[off + M + i2], w1 = prm[off + mi + M + i2];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) {
        pv[jj] = fma_(w0, un[i2][jj], pv[jj]);
        qv[jj] = fma_(w1, un[i2][jj], qv[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) {
      un[i][jj] = pv[jj] * qv[jj];
      if (KEEP) {
        pl[i][jj] = pv[jj];
        ql[i][jj] = qv[jj];
      }
    }
    off += 2 * mi + 2;
  }
  if (NEED_OUT) {
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) out[jj] = prm[off + M + L];
#pragma unroll
    for (int f = 0; f < M; ++f) {
      const T wf = prm[off + f];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) out[jj] = fma_(wf, d[f][J0 + jj], out[jj]);
    }
#pragma unroll(LT >= 0 ? 8 : 1)
    for (int i2 = 0; i2 < L; ++i2) {
      const T wf = prm[off + M + i2];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) out[jj] = fma_(wf, un[i2][jj], out[jj]);
    }
  }
  return off; // offset of the final layer's weights
}

// Cross-lane sum of parameter-gradient partials through a per-warp shared-memory transpose:
// every lane stores its partial of element e at scratch[e][lane] (conflict-free rows), then lane
// l < NE sums row l (LDS.128, padded pitch => conflict-free).  ~2.7x fewer instructions than a
// shuffle tree / recursive halving for the same number of adds, and the partial vector never
// has to live in registers.  Deterministic (fixed order).
template <typename T, int NE> __device__ __forceinline__ T transpose_sum(const T *scratch) {
  constexpr int VEC = TransposeCfg<T>::VEC, PITCH = TransposeCfg<T>::PITCH;
  const int lane = threadIdx.x & 31;
  __syncwarp();
  T s[4] = {T(0), T(0), T(0), T(0)};
  if (NE == 32 || lane < NE) {
    const T *row = scratch + lane * PITCH;
#pragma unroll
    for (int q = 0; q < 32 / VEC; ++q) {
      T t[VEC];
      ld_vec(row + q * VEC, t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[(q * VEC + i) & 3] += t[i];
    }
  }
  __syncwarp();
  return (s[0] + s[1]) + (s[2] + s[3]);
}

// SymNet reverse for PP points of one half-strip: accumulates g (adjoint of the derivative
// features) over output channels and reduces the parameter gradients (rows a7/a11).
// VL (16 or 32) >= M + L + 1 is the padded length of one parameter-gradient vector
// [features 0..M), product units [M..M+L), bias at VL-1].
template <typename T, int M, int P, int PP, int J0, int VL, int C, int LT, int MD = M>
__device__ __forceinline__ void symnet_reverse(const T (&d)[MD][P], const T *syms, int Psym, int L_rt,
                                               const T (&lamv)[C][PP], T (&gacc)[M][PP], double *sg, T *scratch) {
  const int L = LT >= 0 ? LT : L_rt;
  constexpr int UCAP = (VL - 1 - M) < kMaxLayers ? (VL - 1 - M) : kMaxLayers; // static bound on L
  constexpr int TP = TransposeCfg<T>::PITCH;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int f = 0; f < M; ++f)
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) gacc[f][jj] = T(0);

#pragma unroll
  for (int co = 0; co < C; ++co) {
    const T *prm = syms + co * Psym;
    double *sgc = sg + co * Psym;
    T un[kMaxLayers][PP], pl[kMaxLayers][PP], ql[kMaxLayers][PP], eta[kMaxLayers][PP];
    T fout[PP];
    int off = symnet_forward<T, M, P, PP, J0, true, LT, false>(d, prm, L, fout, un, pl, ql);
    // ---- final linear layer: out = wf . h + bf, seed = dt * lambda
    {
      T *col = scratch + lane;
#pragma unroll
      for (int f = 0; f < M; ++f) {
        const T wf = prm[off + f];
        T v = T(0);
#pragma unroll
        for (int jj = 0; jj < PP; ++jj) {
          gacc[f][jj] = fma_(lamv[co][jj], wf, gacc[f][jj]);
          v = fma_(lamv[co][jj], d[f][J0 + jj], v);
        }
        col[f * TP] = v;
      }
#pragma unroll
      for (int i2 = 0; i2 < UCAP; ++i2) {
        if (i2 < L) {
          const T wf = prm[off + M + i2];
          T v = T(0);
#pragma unroll
          for (int jj = 0; jj < PP; ++jj) {
            eta[i2][jj] = lamv[co][jj] * wf;
            v = fma_(lamv[co][jj], un[i2][jj], v);
          }
          col[(M + i2) * TP] = v;
        }
      }
      {
        T v = T(0);
#pragma unroll
        for (int jj = 0; jj < PP; ++jj) v += lamv[co][jj];
        col[(VL - 1) * TP] = v;
      }
      const T tot = transpose_sum<T, VL>(scratch);
      if (lane < M + L)
        sgc[off + lane] += (double)tot;
      else if (lane == VL - 1)
        sgc[off + M + L] += (double)tot;
    }
    // ---- hidden product layers, last to first
#pragma unroll(LT >= 0 ? 8 : 1)
    for (int i = L - 1; i >= 0; --i) {
      const int mi = M + i;
      off -= 2 * mi + 2;
      T pb[PP], qb[PP];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) {
        pb[jj] = eta[i][jj] * ql[i][jj];
        qb[jj] = eta[i][jj] * pl[i][jj];
      }
      // row 0 (p) then row 1 (q) of W_i: propagate, and put the two partial vectors side by side
      // in the transpose scratch (rows [0,VL) and [VL,2VL) when VL == 16)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const T(&eb)[PP] = r == 0 ? pb : qb;
        T *col = scratch + lane + (VL == 16 ? r * VL * TP : 0);
#pragma unroll
        for (int f = 0; f < M; ++f) {
          const T w = prm[off + r * mi + f];
          T v = T(0);
#pragma unroll
          for (int jj = 0; jj < PP; ++jj) {
            gacc[f][jj] = fma_(eb[jj], w, gacc[f][jj]);
            v = fma_(eb[jj], d[f][J0 + jj], v);
          }
          col[f * TP] = v;
        }
#pragma unroll
        for (int i2 = 0; i2 < UCAP; ++i2) {
          if (i2 < i) {
            const T w = prm[off + r * mi + M + i2];
            T v = T(0);
#pragma unroll
            for (int jj = 0; jj < PP; ++jj) {
              eta[i2][jj] = fma_(eb[jj], w, eta[i2][jj]);
              v = fma_(eb[jj], un[i2][jj], v);
            }
            col[(M + i2) * TP] = v;
          }
        }
        {
          T v = T(0);
#pragma unroll
          for (int jj = 0; jj < PP; ++jj) v += eb[jj];
          col[(VL - 1) * TP] = v;
        }
        if (VL == 32) {
          const T tot = transpose_sum<T, 32>(scratch);
          if (lane < mi)
            sgc[off + r * mi + lane] += (double)tot;
          else if (lane == VL - 1)
            sgc[off + 2 * mi + r] += (double)tot;
        }
      }
      if (VL == 16) {
        const T tot = transpose_sum<T, 32>(scratch);
        const int r = lane >> 4, e = lane & 15;
        if (e < mi)
          sgc[off + r * mi + e] += (double)tot;
        else if (e == VL - 1)
          sgc[off + 2 * mi + r] += (double)tot;
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// Static-depth fast path of the SymNet reverse: all P points of the strip at once, in two sweeps.
//   sweep 1 (per output channel): forward (p_i, q_i, units) and the 2*LT+1 error signals
//            S = (dt*lambda, pbar_i, qbar_i); the few unit-weight / bias gradients go to one
//            "tail" transpose;
//   sweep 2 (per derivative feature f): g_f = sum_{co,s} S * W[co,s,f] is stored straight away and
//            the C*(2LT+1) parameter-gradient partials sum_j S*d_f fill FG features' worth of rows
//            of the transpose scratch before each cross-lane row sum.
// Compared with the half-strip path this halves the transposes and parameter loads per point and
// never holds the g accumulators in registers.
template <int M, int C, int LT> struct RevFullCfg {
  static constexpr int NS = 2 * LT + 1;               // error signals per output channel
  static constexpr int RPF = C * NS;                  // scratch rows per feature
  static constexpr int FG = 32 / (RPF > 0 ? RPF : 1); // features per transpose
  static constexpr int NTAIL = C * (LT + 1) * (LT + 1);
  static constexpr bool OK = LT >= 0 && RPF <= 32 && NTAIL <= 32 && FG >= 1;
};

// Tables (shared memory, built once per kernel): feat_base[l] = parameter index (within the
// [C][Psym] block) that scratch row l of a feature transpose maps to for feature 0 of the group
// (add the group's first feature); tail_idx[l] = parameter index of tail row l.
template <int M, int C, int LT> __device__ void build_rev_tables(int Psym, int *feat_base, int *tail_idx) {
  using RC = RevFullCfg<M, C, LT>;
  int offs[LT + 1];
  int off = 0;
  for (int i = 0; i < LT; ++i) {
    offs[i] = off;
    off += 2 * (M + i) + 2;
  }
  offs[LT] = off;
  for (int l = 0; l < 32; ++l) {
    int v = -1;
    if (l < RC::FG * RC::RPF) {
      const int fl = l / RC::RPF, rem = l % RC::RPF, co = rem / RC::NS, sidx = rem % RC::NS;
      if (sidx == 0) {
        v = co * Psym + offs[LT] + fl;
      } else {
        const int i = (sidx - 1) / 2, r = (sidx - 1) % 2;
        v = co * Psym + offs[i] + r * (M + i) + fl;
      }
    }
    feat_base[l] = v;
    tail_idx[l] = -1;
  }
  int t = 0;
  for (int co = 0; co < C; ++co) {
    for (int i2 = 0; i2 < LT; ++i2) tail_idx[t++] = co * Psym + offs[LT] + M + i2;
    tail_idx[t++] = co * Psym + offs[LT] + M + LT;
    for (int i = LT - 1; i >= 0; --i)
      for (int r = 0; r < 2; ++r) {
        for (int i2 = 0; i2 < i; ++i2) tail_idx[t++] = co * Psym + offs[i] + r * (M + i) + M + i2;
        tail_idx[t++] = co * Psym + offs[i] + 2 * (M + i) + r;
      }
  }
}

// XC: the P points of the strip are consecutive along x (g is stored with 16-byte vector stores at gp, gp+VEC, ...)
// instead of P rows apart.
template <typename T, int M, int P, int C, int LT, bool XC = false>
__device__ __forceinline__ void symnet_reverse_full(const T (&d)[M][P], const T *syms, int Psym,
                                                    const T (&lam)[C][P], const bool (&valid)[P], T *gp,
                                                    size_t plane, int W, double *sg, T *scratch,
                                                    const int *feat_base, const int *tail_idx) {
  using RC = RevFullCfg<M, C, LT>;
  constexpr int NS = RC::NS, RPF = RC::RPF, FG = RC::FG, NTAIL = RC::NTAIL;
  constexpr int TP = TransposeCfg<T>::PITCH;
  constexpr int LA = LT > 0 ? LT : 1;
  const int lane = threadIdx.x & 31;
  T *col = scratch + lane;
  T S[C][NS][P];
  int t = 0; // tail row (compile-time after unrolling; same enumeration as build_rev_tables)
#pragma unroll
  for (int co = 0; co < C; ++co) {
    const T *prm = syms + co * Psym;
    T pl[LA][P], ql[LA][P], un[LA][P], eta[LA][P];
    int off = 0;
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      const int mi = M + i;
      T pv[P], qv[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        pv[j] = prm[off + 2 * mi];
        qv[j] = prm[off + 2 * mi + 1];
      }
#pragma unroll
      for (int f = 0; f < M; ++f) {
        const T w0 = prm[off + f], w1 = prm[off + mi + f];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          pv[j] = fma_(w0, d[f][j], pv[j]);
          qv[j] = fma_(w1, d[f][j], qv[j]);
        }
      }
#pragma unroll
      for (int i2 = 0; i2 < i; ++i2) {
        const T w0 = prm[off + M + i2], w1 = prm[off + mi + M + i2];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          pv[j] = fma_(w0, un[i2][j], pv[j]);
          qv[j] = fma_(w1, un[i2][j], qv[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        pl[i][j] = pv[j];
        ql[i][j] = qv[j];
        un[i][j] = pv[j] * qv[j];
      }
      off += 2 * mi + 2;
    }
    // final layer: signal 0 = dt*lambda; unit weights and bias of w_f go to the tail
#pragma unroll
    for (int j = 0; j < P; ++j) S[co][0][j] = lam[co][j];
#pragma unroll
    for (int i2 = 0; i2 < LT; ++i2) {
      const T wf = prm[off + M + i2];
      T v = T(0);
#pragma unroll
      for (int j = 0; j < P; ++j) {
        eta[i2][j] = lam[co][j] * wf;
        v = fma_(lam[co][j], un[i2][j], v);
      }
      col[(t++) * TP] = v;
    }
    {
      T v = T(0);
#pragma unroll
      for (int j = 0; j < P; ++j) v += lam[co][j];
      col[(t++) * TP] = v;
    }
#pragma unroll
    for (int i = LT - 1; i >= 0; --i) {
      const int mi = M + i;
      off -= 2 * mi + 2;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        S[co][1 + 2 * i][j] = eta[i][j] * ql[i][j];
        S[co][2 + 2 * i][j] = eta[i][j] * pl[i][j];
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int i2 = 0; i2 < i; ++i2) {
          const T w = prm[off + r * mi + M + i2];
          T v = T(0);
#pragma unroll
          for (int j = 0; j < P; ++j) {
            eta[i2][j] = fma_(S[co][1 + 2 * i + r][j], w, eta[i2][j]);
            v = fma_(S[co][1 + 2 * i + r][j], un[i2][j], v);
          }
          col[(t++) * TP] = v;
        }
        T v = T(0);
#pragma unroll
        for (int j = 0; j < P; ++j) v += S[co][1 + 2 * i + r][j];
        col[(t++) * TP] = v;
      }
    }
  }
  {
    const T tot = transpose_sum<T, 32>(scratch);
    if (lane < NTAIL) sg[tail_idx[lane]] += (double)tot;
  }
  // ---- sweep 2: derivative features
  int offs[LA + 1];
  {
    int off = 0;
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      offs[i] = off;
      off += 2 * (M + i) + 2;
    }
    offs[LT > 0 ? LT : 0] = off;
  }
  const int fb = feat_base[lane];
#pragma unroll
  for (int f0 = 0; f0 < M; f0 += FG) {
#pragma unroll
    for (int fl = 0; fl < FG; ++fl) {
      const int f = f0 + fl;
      if (f < M) {
        T gv[P];
#pragma unroll
        for (int j = 0; j < P; ++j) gv[j] = T(0);
#pragma unroll
        for (int co = 0; co < C; ++co) {
          const T *prm = syms + co * Psym;
#pragma unroll
          for (int sidx = 0; sidx < NS; ++sidx) {
            const int woff = sidx == 0 ? offs[LT > 0 ? LT : 0]
                                       : offs[(sidx - 1) / 2] + ((sidx - 1) % 2) * (M + (sidx - 1) / 2);
            const T w = prm[woff + f];
            T v = T(0);
#pragma unroll
            for (int j = 0; j < P; ++j) {
              gv[j] = fma_(S[co][sidx][j], w, gv[j]);
              v = fma_(S[co][sidx][j], d[f][j], v);
            }
            col[(fl * RPF + co * NS + sidx) * TP] = v;
          }
        }
        if constexpr (XC) {
          if (valid[0]) {
            constexpr int VEC = VecOf<T>::N;
#pragma unroll
            for (int j = 0; j < P; j += VEC) {
              typename VecOf<T>::type v;
              if constexpr (VEC == 2) {
                v.x = gv[j];
                v.y = gv[j + 1];
              } else {
                v.x = gv[j];
                v.y = gv[j + 1];
                v.z = gv[j + 2];
                v.w = gv[j + 3];
              }
              *reinterpret_cast<typename VecOf<T>::type *>(gp + j) = v;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < P; ++j)
            if (valid[j]) gp[(size_t)j * W] = gv[j];
        }
        gp += plane;
      }
    }
    const T tot = transpose_sum<T, 32>(scratch);
    const int rows = (M - f0 < FG ? M - f0 : FG) * RPF;
    if (lane < rows) sg[fb + f0] += (double)tot;
  }
}

// One half-strip (2 points) of kernel A's epilogue: seed, SymNet reverse, store g.
template <typename T, int M, int P, int J0, int VL, int C, int LT>
__device__ __forceinline__ void reverse_half(const T (&d)[M][P], const T *syms, int Psym, int L,
                                             const bool (&valid)[P], const T (&lam_raw)[C][P],
                                             const StepParams<T> &p, int b, int y0, int x, double *sg,
                                             T *scratch) {
  constexpr int PP = 2;
  const int W = p.W;
  T lamv[C][PP];
#pragma unroll
  for (int co = 0; co < C; ++co)
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) lamv[co][jj] = p.dt * lam_raw[co][J0 + jj];
  T gacc[M][PP];
  symnet_reverse<T, M, P, PP, J0, VL, C, LT>(d, syms, Psym, L, lamv, gacc, sg, scratch);
  // running pointer over the feature planes: one 64-bit add per plane instead of a full index
  T *gp = p.g + ((size_t)b * M * p.g_rows + (y0 + J0 + p.g_off)) * W + x;
  const size_t plane = (size_t)p.g_rows * W;
#pragma unroll
  for (int f = 0; f < M; ++f) {
#pragma unroll
    for (int jj = 0; jj < PP; ++jj)
      if (valid[J0 + jj]) gp[(size_t)jj * W] = gacc[f][jj];
    gp += plane;
  }
}

// --------------------------------------------------------------------------------------------
// Pseudo-upwind selection (row f-1; restated from the description of arXiv 1812.04426, see the
// header include/pdenet_b200.h for the conventions).  Register layout of the feature array in the
// upwind modes: rows [0, MS) are the SymNet inputs f = c*NC + n (NC central filters per channel),
// rows MS + 2c + dir are the flipped twins of channel c's first-order filters (dir 0: order (0,1),
// bank index 1; dir 1: order (1,0), bank index 2).  For output channel `co` the local "velocity" of
// input (c, dir) is a = d SymNet_co / d x_f at the CENTRAL features; a < 0 selects the flipped twin.
// upwind_select evaluates a for PP points, swaps the selected twins INTO the SymNet rows (so the
// unchanged SymNet code sees the selected inputs) and returns the masks; calling upwind_swap with
// the same masks restores the array.
template <typename T, int MS, int C, int NC, int P, int PP, int J0, int MD>
__device__ __forceinline__ void upwind_swap(T (&d)[MD][P], const bool (&mask)[2 * C][PP]) {
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int dir = 0; dir < 2; ++dir)
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) {
        const T a = d[c * NC + 1 + dir][J0 + jj], b = d[MS + 2 * c + dir][J0 + jj];
        d[c * NC + 1 + dir][J0 + jj] = mask[2 * c + dir][jj] ? b : a;
        d[MS + 2 * c + dir][J0 + jj] = mask[2 * c + dir][jj] ? a : b;
      }
}

template <typename T, int MS, int C, int NC, int P, int PP, int J0, int LT, int MD>
__device__ __forceinline__ void upwind_select(T (&d)[MD][P], const T *prm, int L_rt, bool (&mask)[2 * C][PP]) {
  const int L = LT >= 0 ? LT : L_rt;
  T un[kMaxLayers][PP], pl[kMaxLayers][PP], ql[kMaxLayers][PP], eta[kMaxLayers][PP], fout[PP];
  const int off_f = symnet_forward<T, MS, P, PP, J0, true, LT, false>(d, prm, L, fout, un, pl, ql);
  // eta_i = d out / d unit_i, then pl/ql are overwritten with d out / d p_i, d out / d q_i
#pragma unroll(LT >= 0 ? 8 : 1)
  for (int i2 = 0; i2 < L; ++i2) {
    const T wf = prm[off_f + MS + i2];
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) eta[i2][jj] = wf;
  }
  int off = off_f;
#pragma unroll(LT >= 0 ? 8 : 1)
  for (int i = L - 1; i >= 0; --i) {
    const int mi = MS + i;
    off -= 2 * mi + 2;
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) {
      const T pb = eta[i][jj] * ql[i][jj], qb = eta[i][jj] * pl[i][jj];
      pl[i][jj] = pb;
      ql[i][jj] = qb;
    }
#pragma unroll(LT >= 0 ? 8 : 1)
    for (int i2 = 0; i2 < i; ++i2) {
      const T w0 = prm[off + MS + i2], w1 = prm[off + mi + MS + i2];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) eta[i2][jj] = fma_(pl[i][jj], w0, fma_(ql[i][jj], w1, eta[i2][jj]));
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const int f = c * NC + 1 + dir;
      T a[PP];
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) a[jj] = prm[off_f + f];
      int o2 = 0;
#pragma unroll(LT >= 0 ? 8 : 1)
      for (int i = 0; i < L; ++i) {
        const int mi = MS + i;
        const T w0 = prm[o2 + f], w1 = prm[o2 + mi + f];
#pragma unroll
        for (int jj = 0; jj < PP; ++jj) a[jj] = fma_(pl[i][jj], w0, fma_(ql[i][jj], w1, a[jj]));
        o2 += 2 * mi + 2;
      }
#pragma unroll
      for (int jj = 0; jj < PP; ++jj) mask[2 * c + dir][jj] = a[jj] < T(0);
    }
  upwind_swap<T, MS, C, NC, P, PP, J0>(d, mask);
}

// One half-strip (2 points) of the upwind reverse kernel: per output channel select, SymNet reverse on
// the selected inputs, scatter g to the central / flipped planes (first channel stores, later ones add).
template <typename T, int MS, int C, int NC, int P, int J0, int VL, int LT, int MD>
__device__ __forceinline__ void upwind_reverse_half(T (&d)[MD][P], const T *syms, int Psym, int L,
                                                    const bool (&valid)[P], const T (&lam_raw)[C][P],
                                                    const StepParams<T> &p, int b, int y0, int x, double *sgrad_warp,
                                                    T *scratch) {
  constexpr int PP = 2, NF = NC + 2;
  const int W = p.W;
  const size_t plane = (size_t)p.g_rows * W;
  T *g0 = p.g + ((size_t)b * (C * NF) * p.g_rows + (y0 + J0 + p.g_off)) * W + x;
#pragma unroll 1
  for (int co = 0; co < C; ++co) {
    const T *prm = syms + co * Psym;
    bool mask[2 * C][PP];
    upwind_select<T, MS, C, NC, P, PP, J0, LT>(d, prm, L, mask);
    T lamv[1][PP];
#pragma unroll
    for (int jj = 0; jj < PP; ++jj) {
      T lv = T(0);
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) lv = (c2 == co) ? lam_raw[c2][J0 + jj] : lv;
      lamv[0][jj] = p.dt * lv;
    }
    T gacc[MS][PP];
    symnet_reverse<T, MS, P, PP, J0, VL, 1, LT>(d, prm, Psym, L, lamv, gacc, sgrad_warp + co * Psym, scratch);
    // g planes in bank order: channel c, filter n -> plane c*NF + n (n >= NC: flipped twins)
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int n = 0; n < NF; ++n) {
        T *gp = g0 + (size_t)(c * NF + n) * plane;
#pragma unroll
        for (int jj = 0; jj < PP; ++jj) {
          T v;
          if (n >= NC)
            v = mask[2 * c + (n - NC)][jj] ? gacc[c * NC + 1 + (n - NC)][jj] : T(0);
          else if (n == 1 || n == 2)
            v = mask[2 * c + (n - 1)][jj] ? T(0) : gacc[c * NC + n][jj];
          else
            v = gacc[c * NC + n][jj];
          if (valid[J0 + jj]) {
            if (co == 0)
              gp[(size_t)jj * W] = v;
            else
              gp[(size_t)jj * W] += v;
          }
        }
      }
    upwind_swap<T, MS, C, NC, P, PP, J0>(d, mask);
  }
}

// --------------------------------------------------------------------------------------------
// NTO: 0 = the large tile with step_threads(MODE, C*N) threads, one CTA per SM; 128 = the small tile, two CTAs per SM
template <typename T, int K, int N, int C, int MODE, int VL, int LT, int NTO = 0>
__global__ void __launch_bounds__(NTO ? NTO : step_threads_t<T>(MODE, C *N), NTO ? 2 : step_min_blocks<T>(MODE, C *N))
    step_kernel(const StepParams<T> p) {
  using Cfg = StepCfg<T, K, N, C, (NTO ? NTO : step_threads_t<T>(MODE, C * N)), step_p_override<T>(MODE, C * N),
                      step_spw_override(MODE)>;
  constexpr int R = Cfg::R, TW = Cfg::TW, TH = Cfg::TH, P = Cfg::P, SR = Cfg::SR, SC = Cfg::SC;
  constexpr int PITCH = Cfg::PITCH, M = Cfg::M, NT = Cfg::NT;
  constexpr bool LINEAR = (MODE == MODE_LIN_FWD || MODE == MODE_LIN_BWD_A);
  constexpr bool KEPT = mode_is_kept(MODE);       // features come from HBM: no tile staging, no block barrier
  constexpr bool KX = (MODE == MODE_BWD_A_KEPTX); // ... with the thread's P points consecutive along x
  constexpr bool UPW = mode_is_upwind(MODE);      // bank = NC central filters + the 2 flipped first-order twins
  constexpr bool WITH_GRAD = (MODE == MODE_BWD_A || KEPT || MODE == MODE_UPW_BWD_A);
  constexpr int NC = UPW ? N - 2 : N;             // filters feeding the SymNet per channel
  constexpr int MS = C * NC;                      // SymNet inputs (== M unless upwind)

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sgrad = reinterpret_cast<double *>(smem_raw); // [NT/32][grad_elems]  (MODE_BWD_A only)
  T *syms = reinterpret_cast<T *>(sgrad + (WITH_GRAD ? (NT / 32) * Cfg::grad_elems(p.Psym) : 0)); // [C][Psym]
  const T *cw = const_taps<T>(p.slot); // [N][K][K] filter taps in constant memory (uniform operands)
  T *tiles = syms + (LINEAR ? 0 : Cfg::sym_elems(p.Psym)); // [2][C][SR][PITCH]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, L = p.L, Psym = p.Psym;
  const size_t feat_plane = (size_t)H * W;

  // ---- one-time setup: SymNet parameters to shared memory, zeroed gradient accumulators
  if (!LINEAR)
    for (int idx = tid; idx < Cfg::sym_elems(Psym); idx += NT) syms[idx] = idx < C * Psym ? p.symnet[idx] : T(0);
  if (WITH_GRAD)
    for (int idx = tid; idx < (NT / 32) * Cfg::grad_elems(Psym); idx += NT) sgrad[idx] = 0.0;

  auto stage = [&](int item, int buf) {
    const int tx = item % p.tiles_x, t2 = item / p.tiles_x, ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
#pragma unroll
    for (int c = 0; c < C; ++c)
      stage_tile_async<T, SR, SC, PITCH, NT>(tiles + (buf * C + c) * SR * PITCH,
                                             p.u + (size_t)(b * C + c) * p.u_rows * W, p.u_rows, p.u_off,
                                             p.wrap_y, ty * TH - R, tx * TW - R, H, W);
  };

  int item = blockIdx.x, buf = 0;
  if constexpr (!KEPT) {
    if (item < p.n_items) stage(item, 0);
    cp_async_commit();
  }

  const int wx = warp % Cfg::WA, wy = warp / Cfg::WA;
  double *sg = sgrad + warp * Cfg::grad_elems(Psym);
  constexpr int region = Cfg::tile_region(KEPT);
  T *scratch = tiles + region + warp * TransposeCfg<T>::ELEMS; // per-warp transpose scratch (reverse)
  int *rev_tables = reinterpret_cast<int *>(tiles + region + (NT / 32) * TransposeCfg<T>::ELEMS);
  constexpr bool REV_FULL = WITH_GRAD && !UPW && RevFullCfg<M, C, LT>::OK;
  static_assert(!KX || REV_FULL, "the x-contiguous kept-feature mode exists for the static-depth reverse sweep only");
  if constexpr (REV_FULL) {
    if (tid == 0) build_rev_tables<M, C, (LT >= 0 ? LT : 0)>(Psym, rev_tables, rev_tables + 32);
  }

  constexpr int SPW = Cfg::SPW;             // strips per warp and item
  if constexpr (KEPT) __syncthreads();      // setup (SymNet parameters, zeroed accumulators, tables) visible to all
  for (; item < p.n_items; item += gridDim.x) {
    // One barrier per item: this item's tile has landed for every thread and every warp has
    // finished the previous item, so the other buffer can take the next prefetch.
    if constexpr (!KEPT) {
      cp_async_wait<0>();
      __syncthreads();
      const int next = item + gridDim.x;
      if (next < p.n_items) stage(next, buf ^ 1);
      cp_async_commit();
    }

    const int tx = item % p.tiles_x, t2 = item / p.tiles_x, ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
    const T *tile = tiles + buf * Cfg::TILE_ELEMS;
    const int col = wx * 32 + lane;

    // KX geometry: a strip is RPS rows of the full tile width, lane -> (row lane / LPRW, P columns from (lane % LPRW) * P)
    constexpr int LPRW = TW / P, RPS = 32 / LPRW;
    static_assert(!KX || (Cfg::WD * SPW * Cfg::WA * RPS == TH && P % VecOf<T>::N == 0), "KEPTX strip geometry");
#pragma unroll 1
    for (int s = 0; s < SPW; ++s) {
      const int row0 = KX ? ((wy * SPW + s) * Cfg::WA + wx) * RPS : (wy * SPW + s) * P;
      if (ty * TH + row0 >= H) break; // warp-uniform: the rest of this warp's strips lie below the grid (ragged last tile row)
      const int y0 = ty * TH + row0 + (KX ? lane / LPRW : 0);
      const int x = KX ? tx * TW + (lane % LPRW) * P : tx * TW + col;
      bool valid[P];
#pragma unroll
      for (int j = 0; j < P; ++j) valid[j] = KX ? ((x < W) && (y0 < H)) : ((x < W) && (y0 + j < H));
      // reverse mode: issue the adjoint loads now so their latency hides behind the stencils
      T lam_raw[WITH_GRAD ? C : 1][P];
      T d[M][P];
      if constexpr (KX) {
        // P contiguous values per plane: vector loads, one address per plane (W % P == 0: a lane is all in or all out)
        constexpr int VEC = VecOf<T>::N;
#pragma unroll
        for (int co = 0; co < C; ++co) {
          const T *lp = p.lam + ((size_t)(b * C + co) * p.lam_rows + (y0 + p.lam_off)) * W + x;
#pragma unroll
          for (int j = 0; j < P; j += VEC) {
            if (valid[0]) {
              ld_vec(lp + j, &lam_raw[co][j]);
            } else {
#pragma unroll
              for (int i = 0; i < VEC; ++i) lam_raw[co][j + i] = T(0);
            }
          }
        }
        const T *fp = p.feats + ((size_t)b * M * H + y0) * W + x;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int f = 0; f < M; ++f) {
#pragma unroll
          for (int j = 0; j < P; j += VEC) {
            if (valid[0]) {
              ld_vec(fp + j, &d[f][j]);
            } else {
#pragma unroll
              for (int i = 0; i < VEC; ++i) d[f][j + i] = T(0);
            }
          }
          fp += plane;
        }
        // L1 prefetch of the warp's NEXT strip (possibly in the CTA's next item): RPS rows of TW columns per plane
        {
          const int ns = (s + 1 < SPW) ? s + 1 : 0, nit = (s + 1 < SPW) ? item : item + (int)gridDim.x;
          if (nit < p.n_items) {
            const int ntx = nit % p.tiles_x, nt2 = nit / p.tiles_x, nty = nt2 % p.tiles_y, nb = nt2 / p.tiles_y;
            const int nx0 = ntx * TW, ny0 = nty * TH + ((wy * SPW + ns) * Cfg::WA + wx) * RPS;
            constexpr int LPROW = TW * (int)sizeof(T) / 128;   // 128-byte lines per tile row
            constexpr int LPP = RPS * LPROW;                   // lines per plane and strip
            constexpr int NL = (M + C) * LPP;
#pragma unroll
            for (int q = 0; q < (NL + 31) / 32; ++q) {
              const int line = q * 32 + lane, f = line / LPP, rem = line % LPP, row = rem / LPROW;
              const int hx = (rem % LPROW) * (128 / (int)sizeof(T));
              if (line < NL && ny0 + row < H && nx0 + hx < W) {
                const T *src = f < M ? p.feats + (((size_t)nb * M + f) * H + ny0 + row) * W + nx0 + hx
                                     : p.lam + (((size_t)nb * C + (f - M)) * p.lam_rows + ny0 + row + p.lam_off) * W +
                                           nx0 + hx;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
              }
            }
          }
        }
      }
      if constexpr (WITH_GRAD && !KX) {
#pragma unroll
        for (int co = 0; co < C; ++co) {
          const T *lp = p.lam + ((size_t)(b * C + co) * p.lam_rows + (y0 + p.lam_off)) * W + x;
#pragma unroll
          for (int j = 0; j < P; ++j) lam_raw[co][j] = valid[j] ? lp[(size_t)j * W] : T(0);
        }
      }
      // ---- derivative features of every channel at the strip's P points (rows a5-a6)
      if constexpr (KEPT && !KX) {
        const T *fp = p.feats + ((size_t)b * M * H + y0) * W + x;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int f = 0; f < M; ++f) {
#pragma unroll
          for (int j = 0; j < P; ++j) d[f][j] = valid[j] ? fp[(size_t)j * W] : T(0);
          fp += plane;
        }
        // No tile, no stencils: the strip's features come straight from HBM (coalesced 256-byte
        // rows).  While this strip computes, the lines of the warp's NEXT strip (possibly in the
        // CTA's next item) are pulled into L1, so its loads hit the cache.
        {
          const int ns = (s + 1 < SPW) ? s + 1 : 0, nit = (s + 1 < SPW) ? item : item + (int)gridDim.x;
          if (nit < p.n_items) {
            const int ntx = nit % p.tiles_x, nt2 = nit / p.tiles_x, nty = nt2 % p.tiles_y, nb = nt2 / p.tiles_y;
            const int nx0 = ntx * TW + wx * 32, ny0 = nty * TH + (wy * SPW + ns) * P;
            constexpr int LPR = 32 * (int)sizeof(T) / 128; // 128-byte lines per 32-column row
            constexpr int NL = Cfg::KEPT_ROWS * LPR;
#pragma unroll
            for (int q = 0; q < (NL + 31) / 32; ++q) {
              const int line = q * 32 + lane, row = line / LPR, hx = (line % LPR) * (128 / (int)sizeof(T));
              const int f = row / P, j = row % P;
              if (line < NL && ny0 + j < H && nx0 + hx < W) {
                const T *src = f < M ? p.feats + (((size_t)nb * M + f) * H + ny0 + j) * W + nx0 + hx
                                     : p.lam + (((size_t)nb * C + (f - M)) * p.lam_rows + ny0 + j + p.lam_off) * W +
                                           nx0 + hx;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
              }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < (KEPT ? 0 : C); ++c) {
        const T *ut = tile + (c * SR + row0) * PITCH + col;
        T acc[N][P];
        if constexpr (sizeof(T) == 4) {
          // fp32: two filters per packed FMA (FFMA2).  The taps of filters (2np, 2np+1) are one 64-bit uniform load
          // from the filter-innermost copy of the bank, the window value is broadcast inside the instruction: half
          // the FMA issue slots of the scalar loop for the same, bit-identical arithmetic (the fp32 kernels are
          // issue-bound).
          constexpr int NP2 = (N + 1) / 2;
          const float2 *cw2 = const_taps_paired(p.slot);
          float2 acc2[NP2][P];
#pragma unroll
          for (int np = 0; np < NP2; ++np)
#pragma unroll
            for (int j = 0; j < P; ++j) acc2[np][j] = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int bb = 0; bb < K; ++bb) {
            float win[P + K - 1];
#pragma unroll
            for (int i = 0; i < P + K - 1; ++i) win[i] = ut[i * PITCH + bb];
#pragma unroll
            for (int a = 0; a < K; ++a)
#pragma unroll
              for (int np = 0; np < NP2; ++np) {
                const float2 w2 = cw2[(a * K + bb) * NP2 + np]; // warp-uniform -> LDCU.64 + uniform-register pair
#pragma unroll
                for (int j = 0; j < P; ++j) acc2[np][j] = ffma2_bcast(win[j + a], w2, acc2[np][j]);
              }
          }
#pragma unroll
          for (int n = 0; n < N; ++n)
#pragma unroll
            for (int j = 0; j < P; ++j) acc[n][j] = (n & 1) ? acc2[n / 2][j].y : acc2[n / 2][j].x;
        } else {
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
          for (int j = 0; j < P; ++j) acc[n][j] = T(0);
#pragma unroll 1
        for (int bb = 0; bb < K; ++bb) {
          T win[P + K - 1];
#pragma unroll
          for (int i = 0; i < P + K - 1; ++i) win[i] = ut[i * PITCH + bb];
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int n = 0; n < N; ++n) {
              const T w = cw[(n * K + a) * K + bb]; // warp-uniform -> LDCU + uniform-register operand
#pragma unroll
              for (int j = 0; j < P; ++j) acc[n][j] = fma_(win[j + a], w, acc[n][j]);
            }
        }
        }
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
          for (int j = 0; j < P; ++j) d[UPW ? (n < NC ? c * NC + n : MS + 2 * c + (n - NC)) : c * N + n][j] = acc[n][j];
        if constexpr (MODE == MODE_FWD) {
          // keep this channel's features for the reverse pass (written once, read once: streaming
          // stores); issued per channel, while few registers are live, with 32-bit row offsets
          if (p.feats) {
            T *fp = p.feats + (((size_t)b * M + c * N) * H + y0) * W + x;
#pragma unroll
            for (int n = 0; n < N; ++n) {
#pragma unroll
              for (int j = 0; j < P; ++j)
                if (valid[j]) __stcs(fp + j * W, acc[n][j]);
              fp += feat_plane;
            }
          }
        }
      }

      if constexpr (MODE == MODE_FEATS) {
#pragma unroll
        for (int f = 0; f < M; ++f)
#pragma unroll
          for (int j = 0; j < P; ++j)
            if (valid[j]) p.out[((size_t)(b * M + f) * p.out_rows + (y0 + j + p.out_off)) * W + x] = d[f][j];
      } else if constexpr (MODE == MODE_FWD) {
        // ---- SymNet + Euler (rows a7-a9)
#pragma unroll 1
        for (int co = 0; co < C; ++co) {
          T un[kMaxLayers][P], fout[P];
          symnet_forward<T, M, P, P, 0, false, LT>(d, syms + co * Psym, L, fout, un, un, un);
          const T *uc = tile + (co * SR + row0 + R) * PITCH + col + R;
          T *op = p.out + ((size_t)(b * C + co) * p.out_rows + (y0 + p.out_off)) * W + x;
#pragma unroll
          for (int j = 0; j < P; ++j)
            if (valid[j]) {
              const T v = p.rhs_only ? fout[j] : fma_(p.dt, fout[j], uc[j * PITCH]);
              op[(size_t)j * W] = v;
              if (p.mirror_rows) { // fused halo exchange: peer stores over NVLink
                const int y = y0 + j;
                const size_t img = (size_t)(b * C + co) * p.out_rows;
                if (y < p.mirror_rows) p.peer_up[(img + p.out_off + H + y) * W + x] = v;
                if (y >= H - p.mirror_rows) p.peer_dn[(img + p.out_off + y - H) * W + x] = v;
              }
            }
        }
      } else if constexpr (MODE == MODE_UPW_FWD) {
        // ---- pseudo-upwind step (row f-1): per output channel select the first-order inputs, SymNet, Euler
#pragma unroll 1
        for (int co = 0; co < C; ++co) {
          bool mask[2 * C][P];
          upwind_select<T, MS, C, NC, P, P, 0, LT>(d, syms + co * Psym, L, mask);
          T un[kMaxLayers][P], fout[P];
          symnet_forward<T, MS, P, P, 0, false, LT>(d, syms + co * Psym, L, fout, un, un, un);
          upwind_swap<T, MS, C, NC, P, P, 0>(d, mask);
          const T *uc = tile + (co * SR + row0 + R) * PITCH + col + R;
          T *op = p.out + ((size_t)(b * C + co) * p.out_rows + (y0 + p.out_off)) * W + x;
#pragma unroll
          for (int j = 0; j < P; ++j)
            if (valid[j]) op[(size_t)j * W] = p.rhs_only ? fout[j] : fma_(p.dt, fout[j], uc[j * PITCH]);
        }
      } else if constexpr (MODE == MODE_UPW_BWD_A) {
        upwind_reverse_half<T, MS, C, NC, P, 0, VL, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
        upwind_reverse_half<T, MS, C, NC, P, 2, VL, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
        if constexpr (P == 8) {
          upwind_reverse_half<T, MS, C, NC, P, 4, VL, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
          upwind_reverse_half<T, MS, C, NC, P, 6, VL, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
        }
      } else if constexpr (WITH_GRAD && !LINEAR) {
        // ---- SymNet reverse, two points at a time (register budget), rows a7/a11
        if constexpr (REV_FULL) {
          T lamv[C][P];
#pragma unroll
          for (int co = 0; co < C; ++co)
#pragma unroll
            for (int j = 0; j < P; ++j) lamv[co][j] = p.dt * lam_raw[co][j];
          T *gp = p.g + ((size_t)b * M * p.g_rows + (y0 + p.g_off)) * W + x;
          symnet_reverse_full<T, M, P, C, (LT >= 0 ? LT : 0), KX>(d, syms, Psym, lamv, valid, gp, (size_t)p.g_rows * W,
                                                                 W, sg, scratch, rev_tables, rev_tables + 32);
        } else {
          reverse_half<T, M, P, 0, VL, C, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
          reverse_half<T, M, P, 2, VL, C, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
          if constexpr (P == 8) {
            reverse_half<T, M, P, 4, VL, C, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
            reverse_half<T, M, P, 6, VL, C, LT>(d, syms, Psym, L, valid, lam_raw, p, b, y0, x, sg, scratch);
          }
        }
      } else if constexpr (MODE == MODE_LIN_FWD) {
        // row a12: u+ = D0*u + dt * sum_n coe_n (D_n*u); filter 0 is D0, C == 1.  Loads of the
        // coefficient fields are issued for the whole strip before use (running pointers).
        const size_t plane = (size_t)H * W;
        const int xc = x < W ? x : 0; // lanes right of the grid read column 0 (discarded): stays in bounds
        const T *cp = p.symnet + (size_t)y0 * W + xc;
        T acc[P];
#pragma unroll
        for (int j = 0; j < P; ++j) acc[j] = T(0);
#pragma unroll
        for (int n = 1; n < N; ++n) {
#pragma unroll
          for (int j = 0; j < P; ++j) acc[j] = fma_(cp[(size_t)(valid[j] ? j : 0) * W], d[n][j], acc[j]);
          cp += plane;
        }
        T *op = p.out + ((size_t)b * p.out_rows + (y0 + p.out_off)) * W + x;
#pragma unroll
        for (int j = 0; j < P; ++j)
          if (valid[j]) op[(size_t)j * W] = fma_(p.dt, acc[j], d[0][j]);
      } else if constexpr (MODE == MODE_LIN_BWD_A) {
        // g_0 = lam, g_n = dt*coe_n*lam;  dcoe_n(y,x) += dt*lam*d_n  (per trajectory, summed later)
        const size_t plane = (size_t)H * W, gplane = (size_t)p.g_rows * W;
        const T *lp = p.lam + ((size_t)b * p.lam_rows + (y0 + p.lam_off)) * W + x;
        T lam[P], lamdt[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          lam[j] = valid[j] ? lp[(size_t)j * W] : T(0);
          lamdt[j] = p.dt * lam[j];
        }
        T *gp = p.g + ((size_t)b * N * p.g_rows + (y0 + p.g_off)) * W + x;
#pragma unroll
        for (int j = 0; j < P; ++j)
          if (valid[j]) gp[(size_t)j * W] = lam[j];
        const int xc = x < W ? x : 0; // invalid lanes load from column 0 (in bounds) and never store
        const T *cp = p.symnet + (size_t)y0 * W + xc;
        T *dc = p.dcoe_partial + (size_t)b * (N - 1) * plane + (size_t)y0 * W + xc;
        // batches of 4 filters: all loads of a batch are issued before its stores (the compiler
        // cannot move loads across the stores itself -- the pointers may alias)
        constexpr int NB = 4;
#pragma unroll
        for (int n0 = 1; n0 < N; n0 += NB) {
          T cv[NB][P], dv[NB][P];
#pragma unroll
          for (int q = 0; q < NB; ++q)
            if (n0 + q < N) {
#pragma unroll
              for (int j = 0; j < P; ++j) {
                const size_t o = (size_t)q * plane + (size_t)(valid[j] ? j : 0) * W;
                cv[q][j] = cp[o];
                dv[q][j] = dc[o];
              }
            }
#pragma unroll
          for (int q = 0; q < NB; ++q)
            if (n0 + q < N) {
              gp += gplane;
#pragma unroll
              for (int j = 0; j < P; ++j)
                if (valid[j]) {
                  gp[(size_t)j * W] = lamdt[j] * cv[q][j];
                  dc[(size_t)q * plane + (size_t)j * W] = fma_(lamdt[j], d[n0 + q][j], dv[q][j]);
                }
            }
          cp += (size_t)NB * plane;
          dc += (size_t)NB * plane;
        }
      }
    }
    buf ^= 1;
  }
  if constexpr (!KEPT) cp_async_wait<0>();

  if constexpr (WITH_GRAD) {
    __syncthreads();
    const int CP = C * Psym;
    for (int idx = tid; idx < CP; idx += NT) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tot += sgrad[w * Cfg::grad_elems(Psym) + idx];
      if (!p.skip_sym_grad) p.sym_partial[(size_t)blockIdx.x * CP + idx] += tot;
    }
  }
}

} // namespace pdenet
