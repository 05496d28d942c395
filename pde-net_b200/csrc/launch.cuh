// Host-side launchers, instantiated once per (scalar type, filter size) translation unit.
#pragma once
#include "params.cuh"

namespace pdenet {

// (N filters, C channels) pairs compiled for the SymNet (PDE-Net 2.0) kernels: max_order 1..4
// with up to 3 state channels, bounded by the register budget for the C*N features per point.
#define PDENET_FOR_NC(X) X(1, 1) X(3, 1) X(3, 2) X(3, 3) X(6, 1) X(6, 2) X(6, 3) X(10, 1) X(15, 1)
// (bank size incl. the two flipped first-order twins, C channels) compiled for the pseudo-upwind scheme
// (row f-1): max_order 1..2 with up to 2 state channels
#define PDENET_FOR_UPW(X) X(5, 1) X(5, 2) X(8, 1) X(8, 2)
// filter counts (incl. the order-(0,0) filter D0) compiled for the linear (PDE-Net 1) kernels: max_order 1..4
#define PDENET_FOR_LIN(X) X(3) X(6) X(10) X(15)

// returns 0 on success, -2 if (mode, N, C, VL) is not compiled, -4 on a CUDA error; fills the
// tile decomposition from p.B/H/W and launches min(items, sms) persistent CTAs
template <typename T, int K>
int launch_step(int mode, int N, int C, int VL, const StepParams<T> &p, int sms, cudaStream_t st);
template <typename T, int K> int launch_adjoint(const AdjParams<T> &p, int grid, int small, cudaStream_t st);
// dynamic shared memory (bytes) a step kernel of this bank needs in `mode`; 0 if the bank is not compiled
template <typename T, int K> size_t step_smem_bytes(int mode, int N, int C, int Psym);
template <typename T, int K> int launch_filter_grad(const FGradParams<T> &p, int grid, int small, cudaStream_t st);
template <typename T, int K> void filter_grad_tile(int small, int &tw, int &th);
template <typename T, int K> int adjoint_vlb();
// copy n_filters [K][K] filters (device memory) into constant slot `slot`, starting at filter index `first`
template <typename T, int K> int upload_taps(int slot, int first, const T *taps, int n_filters, cudaStream_t st);
template <typename T, int K> int upload_raw(int slot, size_t offset, const void *src, size_t bytes, cudaStream_t st);
template <typename T, int K> void adjoint_tile(int small, int &tw, int &th);

} // namespace pdenet
