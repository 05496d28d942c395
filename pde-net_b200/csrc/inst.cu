// One translation unit per (PDENET_T, PDENET_K): explicit instantiation of the launchers.
#define PDENET_DEFINE_CONST
#include "launch.cuh"

#include "adjoint_kernels.cuh"
#include "step_kernels.cuh"

#ifndef PDENET_T
#error "compile with -DPDENET_T=double|float -DPDENET_K=3|5|7"
#endif

namespace pdenet {

// opt in to the large dynamic shared-memory carve-out (idempotent, cheap)
template <typename KernelT> static int allow_smem(KernelT kern, size_t smem) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 0 : -4;
}

constexpr size_t kMaxDynamicSmem = 227 * 1024;

template <typename T, int K, int N, int C, int MODE, int VL, int LT = -1, int NTO = 0>
static int launch_step_one(const StepParams<T> &p_in, int sms, cudaStream_t st) {
  using Cfg = StepCfg<T, K, N, C, (NTO ? NTO : step_threads_t<T>(MODE, C * N)), step_p_override<T>(MODE, C * N),
                      step_spw_override(MODE)>;
  StepParams<T> p = p_in;
  p.tiles_x = (p.W + Cfg::TW - 1) / Cfg::TW;
  p.tiles_y = (p.H + Cfg::TH - 1) / Cfg::TH;
  p.n_items = p.B * p.tiles_x * p.tiles_y;
  const int slots = sms * (NTO ? kSmallTileCtasPerSm : step_min_blocks<T>(MODE, C * N));
  const int grid = p.n_items < slots ? p.n_items : slots;
  constexpr bool REV = MODE == MODE_BWD_A || mode_is_kept(MODE) || MODE == MODE_UPW_BWD_A;
  const size_t smem = Cfg::smem_bytes(p.Psym, REV, mode_is_kept(MODE));
  if (smem > kMaxDynamicSmem) return -2; // e.g. 3-channel banks with deep SymNets in the recomputing reverse kernel
  auto kern = step_kernel<T, K, N, C, MODE, VL, LT, NTO>;
  if (allow_smem(kern, smem) != 0) return -4;
  kern<<<grid, Cfg::NT, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// Small tiles (32 x 16 in fp64) are compiled for the headline bank (N = 6, C = 2, 2-layer SymNet) and used when
// the large tiles would be too few to fill / balance the SMs: fewer than kSmallTileFactor items per SM.
template <typename T, int K>
static bool want_small_tiles(int mode, int N, int C, int VL, const StepParams<T> &p, int sms) {
  if (!(N == 6 && C == 2 && p.L == 2 && VL == 16) || p.mirror_rows || !p.wrap_y) return false;
  if (!(mode == MODE_FWD || mode == MODE_BWD_A || mode == MODE_BWD_A_KEPT)) return false;
  using Big = StepCfg<T, K, 6, 2, step_threads(MODE_FWD, 12)>;
  const long long items = (long long)p.B * ((p.W + Big::TW - 1) / Big::TW) * ((p.H + Big::TH - 1) / Big::TH);
  return items < (long long)small_tile_factor() * sms;
}

// kernel A on kept features with x-contiguous strips (MODE_BWD_A_KEPTX): 4 points per thread along x, vector accesses
template <typename T> static bool keptx_ok(const StepParams<T> &p, int VL) {
  auto al = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // fp32 only: measured 0.93 -> 0.82 ms per 32 trajectories there, no gain in fp64 (1.62 ms either way: that kernel is
  // bound by its shared-memory transposes at 8 warps per SM, not by address arithmetic), where the vertical strips
  // also keep the SymNet gradient bit-identical to the recomputing path
  return sizeof(T) == 4 && VL == 16 && p.L == 2 && p.W % 4 == 0 && al(p.feats) && al(p.g) && al(p.lam);
}

template <typename T, int K>
int launch_step(int mode_in, int N, int C, int VL, const StepParams<T> &p, int sms, cudaStream_t st) {
  const int mode = (mode_in == MODE_BWD_A_KEPT && keptx_ok(p, VL)) ? MODE_BWD_A_KEPTX : mode_in;
  if (want_small_tiles<T, K>(mode_in, N, C, VL, p, sms)) {
    if (mode == MODE_FWD) return launch_step_one<T, K, 6, 2, MODE_FWD, 16, 2, 128>(p, sms, st);
    if (mode == MODE_BWD_A) return launch_step_one<T, K, 6, 2, MODE_BWD_A, 16, 2, 128>(p, sms, st);
    if (mode == MODE_BWD_A_KEPT) return launch_step_one<T, K, 6, 2, MODE_BWD_A_KEPT, 16, 2, 128>(p, sms, st);
    if constexpr (sizeof(T) == 4)
      if (mode == MODE_BWD_A_KEPTX) return launch_step_one<T, K, 6, 2, MODE_BWD_A_KEPTX, 16, 2, 128>(p, sms, st);
  }
#define PDENET_CASE(NN, CC)                                                                                       \
  if (N == NN && C == CC) {                                                                                      \
    if (mode == MODE_FWD) {                                                                                      \
      if (p.L == 2) return launch_step_one<T, K, NN, CC, MODE_FWD, 16, 2>(p, sms, st);                           \
      return launch_step_one<T, K, NN, CC, MODE_FWD, 16>(p, sms, st);                                            \
    }                                                                                                            \
    if (mode == MODE_FEATS) return launch_step_one<T, K, NN, CC, MODE_FEATS, 16>(p, sms, st);                   \
    if (mode == MODE_BWD_A) {                                                                                    \
      if constexpr (NN * CC + 3 <= 16)                                                                           \
        if (VL == 16 && p.L == 2) return launch_step_one<T, K, NN, CC, MODE_BWD_A, 16, 2>(p, sms, st);           \
      if constexpr (NN * CC + 1 < 16)                                                                            \
        if (VL == 16) return launch_step_one<T, K, NN, CC, MODE_BWD_A, 16>(p, sms, st);                         \
      if (VL == 32) return launch_step_one<T, K, NN, CC, MODE_BWD_A, 32>(p, sms, st);                           \
    }                                                                                                            \
    if (mode == MODE_BWD_A_KEPTX) {                                                                              \
      if constexpr (NN * CC + 3 <= 16 && sizeof(T) == 4)                                                         \
        return launch_step_one<T, K, NN, CC, MODE_BWD_A_KEPTX, 16, 2>(p, sms, st);                               \
    }                                                                                                            \
    if (mode == MODE_BWD_A_KEPT || mode == MODE_BWD_A_KEPTX) {                                                   \
      if constexpr (NN * CC + 3 <= 16)                                                                           \
        if (VL == 16 && p.L == 2) return launch_step_one<T, K, NN, CC, MODE_BWD_A_KEPT, 16, 2>(p, sms, st);      \
      if constexpr (NN * CC + 1 < 16)                                                                            \
        if (VL == 16) return launch_step_one<T, K, NN, CC, MODE_BWD_A_KEPT, 16>(p, sms, st);                    \
      if (VL == 32) return launch_step_one<T, K, NN, CC, MODE_BWD_A_KEPT, 32>(p, sms, st);                      \
    }                                                                                                            \
  }
  PDENET_FOR_NC(PDENET_CASE)
#undef PDENET_CASE
  // pseudo-upwind: NN = bank size incl. the two flipped twins, the SymNet sees CC*(NN-2) inputs
#define PDENET_CASE(NN, CC)                                                                                       \
  if (N == NN && C == CC) {                                                                                      \
    if (mode == MODE_UPW_FWD) {                                                                                  \
      if (p.L == 2) return launch_step_one<T, K, NN, CC, MODE_UPW_FWD, 16, 2>(p, sms, st);                       \
      return launch_step_one<T, K, NN, CC, MODE_UPW_FWD, 16>(p, sms, st);                                        \
    }                                                                                                            \
    if (mode == MODE_UPW_BWD_A) {                                                                                \
      if constexpr ((NN - 2) * CC + 3 <= 16)                                                                     \
        if (VL == 16 && p.L == 2) return launch_step_one<T, K, NN, CC, MODE_UPW_BWD_A, 16, 2>(p, sms, st);       \
      if constexpr ((NN - 2) * CC + 1 < 16)                                                                      \
        if (VL == 16) return launch_step_one<T, K, NN, CC, MODE_UPW_BWD_A, 16>(p, sms, st);                     \
      if (VL == 32) return launch_step_one<T, K, NN, CC, MODE_UPW_BWD_A, 32>(p, sms, st);                       \
    }                                                                                                            \
  }
  PDENET_FOR_UPW(PDENET_CASE)
#undef PDENET_CASE
#define PDENET_CASE(NN)                                                                                           \
  if (N == NN && C == 1) {                                                                                       \
    if (mode == MODE_LIN_FWD) return launch_step_one<T, K, NN, 1, MODE_LIN_FWD, 16>(p, sms, st);                \
    if (mode == MODE_LIN_BWD_A) return launch_step_one<T, K, NN, 1, MODE_LIN_BWD_A, 16>(p, sms, st);            \
  }
  PDENET_FOR_LIN(PDENET_CASE)
#undef PDENET_CASE
  return -2;
}

template <typename T, int K> size_t step_smem_bytes(int mode, int N, int C, int Psym) {
  const bool rev = mode == MODE_BWD_A || mode == MODE_BWD_A_KEPT;
#define PDENET_CASE(NN, CC)                                                                                       \
  if (N == NN && C == CC)                                                                                        \
    return StepCfg<T, K, NN, CC, step_threads(MODE_BWD_A, NN * CC)>::smem_bytes(Psym, rev, mode == MODE_BWD_A_KEPT);
  PDENET_FOR_NC(PDENET_CASE)
#undef PDENET_CASE
  return 0;
}

// `small` selects the 16-row tile variants; the caller sizes tiles_x/tiles_y/n_items with adjoint_tile /
// filter_grad_tile of the same variant
template <typename T, int K, int NT_> static int launch_adjoint_v(const AdjParams<T> &p, int grid, cudaStream_t st) {
  using Cfg = AdjCfg<T, K, NT_>;
  const size_t smem = Cfg::smem_bytes(p.N);
  if (smem > 227 * 1024) return -2;
  auto kern = adjoint_kernel<T, K, NT_>;
  if (allow_smem(kern, smem) != 0) return -4;
  kern<<<grid, Cfg::NT, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
template <typename T, int K> int launch_adjoint(const AdjParams<T> &p, int grid, int small, cudaStream_t st) {
  return small ? launch_adjoint_v<T, K, 64>(p, grid, st) : launch_adjoint_v<T, K, kAdjNT>(p, grid, st);
}

template <typename T, int K, int MT_> static int launch_filter_grad_v(const FGradParams<T> &p, int grid, cudaStream_t st) {
  using Cfg = FGradCfg<T, K, MT_>;
  const size_t smem = Cfg::smem_bytes();
  auto kern = filter_grad_kernel<T, K, MT_>;
  if (allow_smem(kern, smem) != 0) return -4;
  kern<<<grid, Cfg::NT, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
// `small`: 0 = large tile (kFGradMT micro-tiles of 8 rows), 1 = small (16 rows), 2 = medium (kFGradMediumMT: 32 rows)
template <typename T, int K> int launch_filter_grad(const FGradParams<T> &p, int grid, int small, cudaStream_t st) {
  if (small == 2) return launch_filter_grad_v<T, K, kFGradMediumMT>(p, grid, st);
  return small ? launch_filter_grad_v<T, K, 2>(p, grid, st) : launch_filter_grad_v<T, K, kFGradMT>(p, grid, st);
}
template <typename T, int K> void filter_grad_tile(int small, int &tw, int &th) {
  tw = FGradCfg<T, K>::TW;
  th = small == 2 ? FGradCfg<T, K, kFGradMediumMT>::TH : (small ? FGradCfg<T, K, 2>::TH : FGradCfg<T, K, kFGradMT>::TH);
}

template <typename T, int K> int upload_taps(int slot, int first, const T *taps, int n_filters, cudaStream_t st) {
  if (slot < 0 || slot >= kConstSlots || (size_t)(first + n_filters) * K * K * sizeof(T) > (size_t)kConstSlotBytes)
    return -2;
  const cudaError_t e =
      cudaMemcpyToSymbolAsync(c_slots, taps, (size_t)n_filters * K * K * sizeof(T),
                              (size_t)slot * kConstSlotBytes + (size_t)first * K * K * sizeof(T),
                              cudaMemcpyDeviceToDevice, st);
  return e == cudaSuccess ? 0 : -4;
}

// raw stream-ordered copy into a constant slot at a byte offset (the filter-innermost fp32 copy of the bank)
template <typename T, int K> int upload_raw(int slot, size_t offset, const void *src, size_t bytes, cudaStream_t st) {
  if (slot < 0 || slot >= kConstSlots || offset + bytes > (size_t)kConstSlotBytes) return -2;
  return cudaMemcpyToSymbolAsync(c_slots, src, bytes, (size_t)slot * kConstSlotBytes + offset, cudaMemcpyDeviceToDevice,
                                 st) == cudaSuccess
             ? 0
             : -4;
}

template <typename T, int K> int adjoint_vlb() { return AdjCfg<T, K>::VLB; }
template <typename T, int K> void adjoint_tile(int small, int &tw, int &th) {
  tw = AdjCfg<T, K>::TW;
  th = small ? AdjCfg<T, K, 64>::TH : AdjCfg<T, K, kAdjNT>::TH;
}

template int launch_step<PDENET_T, PDENET_K>(int, int, int, int, const StepParams<PDENET_T> &, int, cudaStream_t);
template size_t step_smem_bytes<PDENET_T, PDENET_K>(int, int, int, int);
template int launch_adjoint<PDENET_T, PDENET_K>(const AdjParams<PDENET_T> &, int, int, cudaStream_t);
template int launch_filter_grad<PDENET_T, PDENET_K>(const FGradParams<PDENET_T> &, int, int, cudaStream_t);
template void filter_grad_tile<PDENET_T, PDENET_K>(int, int &, int &);
template int adjoint_vlb<PDENET_T, PDENET_K>();
template int upload_taps<PDENET_T, PDENET_K>(int, int, const PDENET_T *, int, cudaStream_t);
template int upload_raw<PDENET_T, PDENET_K>(int, size_t, const void *, size_t, cudaStream_t);
template void adjoint_tile<PDENET_T, PDENET_K>(int, int &, int &);

} // namespace pdenet
