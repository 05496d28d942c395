// Shared device helpers for the PDE-Net hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdenet {

constexpr int kMaxLayers = 8;   // PDENET_MAX_LAYERS
constexpr int kMaxFilters = 16; // PDENET_MAX_FILTERS

// Filter taps live in constant memory: a warp-uniform constant index lets the compiler fetch them
// with LDCU into UNIFORM registers and issue `DFMA R, R, UR, R`, which needs only two vector
// register operands.  Measured on B200 (tools/probes): a DFMA with three
// distinct vector-register operands is register-bank limited to ~73 % of the pipe rate, with a
// uniform-register operand it reaches ~95 %.  A small ring of slots (guarded by events on the
// host side) keeps concurrent calls on different streams from overwriting each other's taps.
constexpr int kConstSlots = 8;
constexpr int kConstSlotBytes = kMaxFilters * 49 * 8; // up to 16 filters of 7x7 doubles
#ifdef PDENET_DEFINE_CONST
__constant__ __align__(16) unsigned char c_slots[kConstSlots * kConstSlotBytes];
template <typename T> __device__ __forceinline__ const T *const_taps(int slot) {
  return reinterpret_cast<const T *>(c_slots + slot * kConstSlotBytes);
}
// fp32 only: the second half of a slot holds the bank a second time, FILTER-innermost and padded to an even filter
// count -- [K*K taps][NP2 pairs][2] -- so that the taps of two filters are one 64-bit uniform load and feed one
// packed FFMA2 (see step_kernels.cuh)
__device__ __forceinline__ const float2 *const_taps_paired(int slot) {
  return reinterpret_cast<const float2 *>(c_slots + slot * kConstSlotBytes + kConstSlotBytes / 2);
}
#endif
constexpr int kPairedBytes = kConstSlotBytes / 2;

// packed single-precision FMA (SASS FFMA2): d = (x * w.x + c.x, x * w.y + c.y).  The broadcast of x and a
// uniform-register pair for w are folded into the instruction by ptxas (`FFMA2 R, R.F32, UR.F32x2.HI_LO, R`).
__device__ __forceinline__ float2 ffma2_bcast(float x, float2 w, float2 c) {
  float2 d;
  asm("{\n .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %2};\n mov.b64 rb, {%3, %4};\n mov.b64 rc, {%5, %6};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd;\n}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(x), "f"(w.x), "f"(w.y), "f"(c.x), "f"(c.y));
  return d;
}

// 16-byte vector type per scalar type (LDS.128 / STG.128 granularity).
template <typename T> struct VecOf;
template <> struct VecOf<double> {
  using type = double2;
  static constexpr int N = 2;
};
template <> struct VecOf<float> {
  using type = float4;
  static constexpr int N = 4;
};

// 16-byte shared/global load into consecutive elements of a register array
__device__ __forceinline__ void ld_vec(const double *s, double *d) {
  const double2 t = *reinterpret_cast<const double2 *>(s);
  d[0] = t.x;
  d[1] = t.y;
}
__device__ __forceinline__ void ld_vec(const float *s, float *d) {
  const float4 t = *reinterpret_cast<const float4 *>(s);
  d[0] = t.x;
  d[1] = t.y;
  d[2] = t.z;
  d[3] = t.w;
}

// Field geometry used by every kernel: a field is [B][CH][rows][W] and its logical row 0 sits at
// buffer row `off`.  With wrap_y==1 the field is periodic along H (rows == H, off == 0); with
// wrap_y==0 the caller guarantees that every row a valid output needs (logical -halo ..
// Hc+halo-1) exists in the buffer (slab decomposition with ghost rows).

__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ int wrap_index(int v, int n) {
  v %= n;
  return v < 0 ? v + n : v;
}

// ---- cp.async (LDGSTS) -----------------------------------------------------------------------
__device__ __forceinline__ void cp_async_elem(double *smem, const double *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_elem(float *smem, const float *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
// 16-byte copy, zero-filled when !valid (src-size 0); streams past L1
__device__ __forceinline__ void cp_async_16_zfill(void *smem, const void *gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage a ROWS x COLS window of one image (origin (y0, x0) in logical coordinates, periodic in x,
// periodic in y iff wrap_y) into shared memory with row pitch PITCH.  All NT threads take part;
// lanes run along x so global reads are coalesced.
template <typename T, int ROWS, int COLS, int PITCH, int NT>
__device__ __forceinline__ void stage_tile_async(T *dst, const T *__restrict__ img, int rows_buf, int row_off,
                                                 int wrap_y, int y0, int x0, int H, int W) {
  constexpr int NCB = (COLS + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int gx[NCB];
#pragma unroll
  for (int j = 0; j < NCB; ++j) gx[j] = wrap_index(x0 + lane + 32 * j, W);
  int gyw = y0 + warp; // logical row, wrapped incrementally (one mod per call, not per row)
  if (wrap_y) gyw = wrap_index(gyw, H);
  for (int row = warp; row < ROWS; row += NT / 32) {
    int gy = gyw + row_off;
    // slab mode: rows outside the buffer can only belong to masked (out-of-domain) outputs
    gy = gy < 0 ? 0 : (gy >= rows_buf ? rows_buf - 1 : gy);
    gyw += NT / 32;
    if (wrap_y && gyw >= H) {
      gyw -= H;
      if (gyw >= H) gyw %= H;
    }
    const T *srow = img + (size_t)gy * W;
    T *drow = dst + row * PITCH;
#pragma unroll
    for (int j = 0; j < NCB; ++j) {
      const int lx = lane + 32 * j;
      if (lx < COLS) cp_async_elem(drow + lx, srow + gx[j]);
    }
  }
}

// Same window, staged in 16-byte units.  Requires x0 % VEC == 0, W % VEC == 0, a 16-byte aligned
// image base and PITCH % VEC == 0; may stage up to VEC-1 columns past COLS (still inside PITCH).
template <typename T, int ROWS, int COLS, int PITCH, int NT>
__device__ __forceinline__ void stage_tile_async_vec(T *dst, const T *__restrict__ img, int rows_buf, int row_off,
                                                     int wrap_y, int y0, int x0, int H, int W) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int UNITS = (COLS + VEC - 1) / VEC;
  static_assert(UNITS * VEC <= PITCH && PITCH % VEC == 0, "pitch too small for vector staging");
  constexpr int NCB = (UNITS + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int gx[NCB];
#pragma unroll
  for (int j = 0; j < NCB; ++j) gx[j] = wrap_index(x0 + (lane + 32 * j) * VEC, W);
  int gyw = y0 + warp;
  if (wrap_y) gyw = wrap_index(gyw, H);
  for (int row = warp; row < ROWS; row += NT / 32) {
    int gy = gyw + row_off;
    gy = gy < 0 ? 0 : (gy >= rows_buf ? rows_buf - 1 : gy);
    gyw += NT / 32;
    if (wrap_y && gyw >= H) {
      gyw -= H;
      if (gyw >= H) gyw %= H;
    }
    const T *srow = img + (size_t)gy * W;
    T *drow = dst + row * PITCH;
#pragma unroll
    for (int j = 0; j < NCB; ++j) {
      const int u = lane + 32 * j;
      if (u < UNITS) cp_async_16_zfill(drow + u * VEC, srow + gx[j], true);
    }
  }
}

// ---- warp-level vector reduction by recursive halving ---------------------------------------
// v[0..V) per lane (V a power of two)  ->  V >= 32: every lane ends with V/32 fully reduced
// elements, v[i] = element lane*(V/32)+i;  V < 32: v[0] = element lane/(32/V) (duplicated over
// 32/V lanes).  Cost ~V shuffle+add instead of 5V.  Deterministic (fixed order).
template <typename T, int V, int LEN, int MASK> struct WarpReduceVec {
  static __device__ __forceinline__ void run(T (&v)[V], int lane) {
    if constexpr (MASK >= 1) {
      if constexpr (LEN > 1) {
        constexpr int half = LEN / 2;
        const bool upper = (lane & MASK) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const T send = upper ? v[i] : v[i + half];
          const T keep = upper ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
        }
        WarpReduceVec<T, V, half, MASK / 2>::run(v, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], MASK);
        WarpReduceVec<T, V, 1, MASK / 2>::run(v, lane);
      }
    }
  }
};
template <typename T, int V> __device__ __forceinline__ void warp_reduce_vec(T (&v)[V]) {
  static_assert((V & (V - 1)) == 0, "V must be a power of two");
  WarpReduceVec<T, V, V, 16>::run(v, threadIdx.x & 31);
}

template <typename T> __device__ __forceinline__ T fma_(T a, T b, T c);
template <> __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }

} // namespace pdenet
