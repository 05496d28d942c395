// Kernel parameter blocks and launch geometry shared by the kernels (inst.cu) and the C-ABI (api.cu).
#pragma once
#include "common.cuh"

namespace pdenet {

enum StepMode {
  MODE_FWD = 0,
  MODE_FEATS = 1,
  MODE_BWD_A = 2,
  MODE_LIN_FWD = 3,
  MODE_LIN_BWD_A = 4,
  MODE_BWD_A_KEPT = 5, // reverse kernel A reading the features the forward step kept (no stencil recompute)
  // pseudo-upwind scheme (row f-1): the bank carries two extra filters per channel, the flipped
  // first-order filters, as its last two entries; per output channel and grid point the SymNet sees
  // either a first-order feature or its flipped twin, chosen by the sign of d SymNet / d feature
  MODE_UPW_FWD = 6,
  MODE_UPW_BWD_A = 7,
  // MODE_BWD_A_KEPT with the strip of a thread laid along x (P consecutive points) instead of along y: the kernel
  // runs no stencils, so nothing needs the vertical strip, and P contiguous values are ONE 16-byte (fp32) or two
  // (fp64) vector accesses with one address instead of P scalar accesses P rows apart (17 % of kernel A's warp
  // instructions were 64-bit address arithmetic).  Needs W % P == 0 (and 16-byte aligned planes).
  MODE_BWD_A_KEPTX = 8
};
__host__ __device__ constexpr bool mode_is_kept(int mode) { return mode == MODE_BWD_A_KEPT || mode == MODE_BWD_A_KEPTX; }
__host__ __device__ constexpr bool mode_is_upwind(int mode) { return mode == MODE_UPW_FWD || mode == MODE_UPW_BWD_A; }

template <typename T> struct StepParams {
  // input state [B][C][u_rows][W]
  const T *u;
  int u_rows, u_off, wrap_y;
  int slot; // constant-memory slot holding the N scaled filters [N][K][K] (linear modes: filter 0 = D0)
  const T *symnet;  // [C][Psym]            (linear modes: coe [N-1][H][W])
  // forward output [B][C][out_rows][W] / features [B][C*N][out_rows][W]
  T *out;
  int out_rows, out_off;
  // slab mode with peer halo writes (NVLink peer memory): the first / last `mirror_rows` output
  // rows are also stored into the ring neighbours' output buffers (same layout), where they are
  // the neighbours' bottom / top ghost rows.  mirror_rows == 0: off.
  T *peer_up, *peer_dn;
  int mirror_rows;
  // reverse: adjoint of the step's output [B][C][lam_rows][W]
  const T *lam;
  int lam_rows, lam_off;
  // reverse: derivative adjoints g [B][C*N][g_rows][W]
  T *g;
  int g_rows, g_off;
  // kept derivative features [B][C*N][H][W] of this step's input state (periodic rollout only):
  // MODE_FWD stores them when non-NULL, MODE_BWD_A_KEPT reads them instead of recomputing
  T *feats;
  double *sym_partial; // [gridDim.x][C*Psym], accumulated across launches
  int skip_sym_grad;   // 1: ghost-row launch of a slab -- emit g only, contribute no parameter gradient
  int rhs_only;        // MODE_FWD: 1 = write F(u) = SymNet(FD2d(u)) itself instead of the Euler update (row a8)
  T *dcoe_partial;     // linear modes: [B][N-1][H][W], accumulated across launches
  int B, H, W, L, Psym;
  T dt;
  int tiles_x, tiles_y, n_items;
};

// Threads per CTA: the forward-type kernels fit 168 registers when C*N <= 12, so they run 12
// warps per SM; the reverse kernel A (and the widest banks) need up to 255 registers -> 8 warps.
__host__ __device__ constexpr int step_threads(int mode, int M) {
  return ((mode == MODE_FWD || mode == MODE_FEATS || mode == MODE_LIN_FWD || mode == MODE_UPW_FWD) && M <= 12) ? 384
                                                                                                               : 256;
}

// Small-tile variants (see StepCfg): two 128-thread CTAs per SM; per-CTA partial-sum slots are sized for them.
constexpr int kSmallTileCtasPerSm = 2;
constexpr int kSymSlotsPerSm = kSmallTileCtasPerSm;
// large-tile items per SM below which the small tiles are used (pdenet_small_tile_factor() sets it; 0 = never;
// initial value from PDENET_SMALL_TILE_FACTOR, default 4).  Defined in api.cu.
int small_tile_factor();

// fp32 kernel A on kept features: strips of 4 points (not 8) and 12 warps -- it runs no stencils, so nothing asks
// for the long strip, and the halved register footprint (d[12][4] + S[10][4] floats) lets 12 warps hide its
// global-load and shared-memory latencies (8 warps at 254 registers: 0.33 of the FP32 peak)
#ifndef PDENET_KEPT_F32_THREADS
#define PDENET_KEPT_F32_THREADS 384
#define PDENET_KEPT_F32_BLOCKS 1
#endif
template <typename T> __host__ __device__ constexpr int step_p_override(int mode, int M) {
  return (mode_is_kept(mode) && sizeof(T) == 4 && M <= 12) ? 4 : 0;
}
template <typename T> __host__ __device__ constexpr int step_threads_t(int mode, int M) {
  return (mode_is_kept(mode) && sizeof(T) == 4 && M <= 12) ? PDENET_KEPT_F32_THREADS : step_threads(mode, M);
}
template <typename T> __host__ __device__ constexpr int step_min_blocks(int mode, int M) {
  return (mode_is_kept(mode) && sizeof(T) == 4 && M <= 12) ? PDENET_KEPT_F32_BLOCKS : 1;
}

template <typename T> struct AdjParams {
  const T *g; // [B][C*N][g_rows][W]
  int g_rows, g_off;
  int wrap_y;
  const T *lam_in; // [B][C][li_rows][W]; NULL = 0
  int li_rows, li_off;
  const T *inject; // [B][C][inj_rows][W]; NULL = none
  int inj_rows, inj_off;
  T *lam_out; // [B][C][lo_rows][W]; NULL = skip the data gradient
  int lo_rows, lo_off;
  T *peer_up, *peer_dn; // slab mode: mirror the first / last mirror_rows rows into the neighbours' lam_out
  int mirror_rows;
  int slot;           // constant-memory slot holding the N filters [N][K][K]
  int zz[32];         // all zero; an offset the compiler cannot see through (one per window row), so the
                      // taps are re-fetched per row into uniform registers instead of being kept
                      // live in 98 vector registers
  int vec_ok;         // 1: g may be staged in 16-byte units (W % VEC == 0, base 16-byte aligned)
  int B, C, N, H, W;
  int tiles_x, tiles_y, n_items;
};

// Filter-gradient kernel (B2) geometry.  Warps of a CTA own (filter, row part) pairs.  12 warps in ONE
// CTA per SM put exactly 3 warps on each of the 4 SM sub-partitions; the first version ran 2 CTAs of 6
// warps, which lands 4 / 4 / 2 / 2 warps on the sub-partitions (warp w of a CTA goes to sub-partition
// w mod 4): the two busy ones bound the kernel at 75 % while the others wait at the item barrier (ncu:
// barrier stall 1.04 per issued instruction).
#ifndef PDENET_B2_WARPS
#define PDENET_B2_WARPS 12
#endif
#ifndef PDENET_B1_NT
#define PDENET_B1_NT 256 // threads of the large-tile data-gradient kernel (256: two CTAs per SM; 512: one)
#endif
constexpr int kAdjNT = PDENET_B1_NT;
#ifndef PDENET_B2_MT
#define PDENET_B2_MT 16  // micro-tiles (8 rows each) per large tile: 128-row tiles (64-row tiles: one block barrier per item twice as often, 1.74 vs 1.67 ms)
#endif
constexpr int kFGradWarps = PDENET_B2_WARPS;
constexpr int kFGradMT = PDENET_B2_MT;
// medium tile (32 rows) for launches with too few large tiles but plenty of medium ones (BASELINE config 3: 256 large
// items on 148 CTAs are 1.7 waves; 2048 small items balance but pay 1.4x the halo staging; 1024 medium items: B2
// 74 -> 66 us per launch there)
constexpr int kFGradMediumMT = 4;
// forced tile level of the filter-gradient kernel (-1 = by the policy above; pdenet_filter_grad_tile_level())
int filter_grad_tile_level();
constexpr int kFGradCtasPerSm = kFGradWarps <= 6 ? 2 : 1;

template <typename T> struct FGradParams {
  const T *g; // [B][C*N][g_rows][W]
  int g_rows, g_off;
  const T *u; // [B][C][u_rows][W]
  int u_rows, u_off;
  int wrap_y;
  double *dk_partial; // [gridDim.x * slot_stride][N][VLB], accumulated across launches
  int slot_stride;    // partial slots per CTA = the largest `parts` of any launch of this bank (host: fgrad_max_parts)
  int n0, nw;         // this launch handles filters n0 .. n0+nw-1; warp w -> filter n0 + w % nw,
                      // row part w / nw of `parts` = NWARPS / nw (nw divides NWARPS)
  int vec_ok;         // 1 if g rows may be read with 16-byte loads (W % VEC == 0, base aligned)
  int B, C, N, H, W;
  int tiles_x, tiles_y, n_items;
};

} // namespace pdenet
