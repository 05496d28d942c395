// C-ABI of the PDE-Net hot path (see include/pdenet_b200.h).  No torch types, no global state;
// everything is ordered on the caller's stream.
#include "../../include/pdenet_b200.h"
#include "launch.cuh"
#include "linear_cluster_kernels.cuh"
#include "small_kernels.cuh"

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <mutex>
#include <vector>

namespace pdenet {

// small-tile policy (params.cuh): process-global, settable through pdenet_small_tile_factor()
static std::atomic<int> g_small_tile_factor{[] {
  const char *e = getenv("PDENET_SMALL_TILE_FACTOR");
  return e ? atoi(e) : 4;
}()};
int small_tile_factor() { return g_small_tile_factor.load(std::memory_order_relaxed); }
static std::atomic<int> g_fgrad_level{-1};
int filter_grad_tile_level() { return g_fgrad_level.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// measurement hooks: launch counters (always) + per-launch event pairs (when enabled)
struct ProfRec {
  int kind;
  cudaEvent_t e0, e1;
};
struct Prof {
  std::atomic<bool> on{false};
  std::atomic<long long> launches[PDENET_PROF_KINDS];
  std::mutex m; // guards recs
  std::vector<ProfRec> recs;
  Prof() {
    for (auto &l : launches) l.store(0);
  }
};
static Prof g_prof;

struct ProfScope {
  ProfRec r;
  cudaStream_t st;
  bool live;
  ProfScope(int kind, cudaStream_t s) : st(s), live(g_prof.on.load(std::memory_order_relaxed)) {
    g_prof.launches[kind].fetch_add(1, std::memory_order_relaxed);
    r.kind = kind;
    if (live) {
      cudaEventCreate(&r.e0);
      cudaEventCreate(&r.e1);
      cudaEventRecord(r.e0, st);
    }
  }
  ~ProfScope() {
    if (live) {
      cudaEventRecord(r.e1, st);
      std::lock_guard<std::mutex> lk(g_prof.m);
      g_prof.recs.push_back(r);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// extern templates: the heavy kernels live in inst_<type>_k<K>.o
#define PDENET_EXTERN(TT, KK)                                                                                     \
  extern template int launch_step<TT, KK>(int, int, int, int, const StepParams<TT> &, int, cudaStream_t);        \
  extern template int launch_adjoint<TT, KK>(const AdjParams<TT> &, int, int, cudaStream_t);                     \
  extern template int upload_taps<TT, KK>(int, int, const TT *, int, cudaStream_t);                              \
  extern template int upload_raw<TT, KK>(int, size_t, const void *, size_t, cudaStream_t);                       \
  extern template size_t step_smem_bytes<TT, KK>(int, int, int, int);                                            \
  extern template int launch_filter_grad<TT, KK>(const FGradParams<TT> &, int, int, cudaStream_t);               \
  extern template void filter_grad_tile<TT, KK>(int, int &, int &);                                              \
  extern template int adjoint_vlb<TT, KK>();                                                                     \
  extern template void adjoint_tile<TT, KK>(int, int &, int &);
PDENET_EXTERN(double, 3) PDENET_EXTERN(double, 5) PDENET_EXTERN(double, 7)
PDENET_EXTERN(float, 3) PDENET_EXTERN(float, 5) PDENET_EXTERN(float, 7)
#undef PDENET_EXTERN

// ---------------------------------------------------------------------------------------------
// Ring of constant-memory tap slots (see common.cuh).  A call takes the next slot, makes its
// stream wait for the slot's previous user, uploads its taps (stream-ordered device-to-device
// copy into the __constant__ bank of the matching translation unit) and records the slot's event
// after its last kernel.  This is the library's only cross-call state.
struct SlotRing {
  std::mutex m;
  std::condition_variable cv;
  int next[16] = {0};
  bool leased[16][kConstSlots] = {{false}};
  cudaEvent_t ev[16][kConstSlots];
  bool have[16] = {false};
};
static SlotRing g_slots;

// A lease covers the whole host-side enqueue of a call: the slot is marked `leased` under the
// mutex until the destructor has recorded the slot's release event, so a concurrent call on another
// host thread can never pick a slot whose owner is still enqueuing kernels (it takes the next free
// slot, or blocks on the condition variable if all eight are out).  Device-side ordering between
// consecutive users of a slot is the event: the new user's stream waits for the previous user's
// last kernel before its upload overwrites the taps.
struct SlotLease {
  int slot = -1, dev = 0;
  cudaStream_t st;
  explicit SlotLease(cudaStream_t s) : st(s) {
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return;
    std::unique_lock<std::mutex> lk(g_slots.m);
    if (!g_slots.have[dev]) {
      for (int i = 0; i < kConstSlots; ++i) cudaEventCreateWithFlags(&g_slots.ev[dev][i], cudaEventDisableTiming);
      g_slots.have[dev] = true;
    }
    for (;;) {
      for (int i = 0; i < kConstSlots && slot < 0; ++i) {
        const int c = (g_slots.next[dev] + i) % kConstSlots;
        if (!g_slots.leased[dev][c]) slot = c;
      }
      if (slot >= 0) break;
      g_slots.cv.wait(lk);
    }
    g_slots.leased[dev][slot] = true;
    g_slots.next[dev] = (slot + 1) % kConstSlots;
    // inside a CUDA-graph capture the guard events cannot be used (they were recorded outside the
    // capture); the captured work is stream-ordered within the graph, which is all it needs
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    capturing = cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    if (!capturing) cudaStreamWaitEvent(st, g_slots.ev[dev][slot], 0);
  }
  ~SlotLease() {
    if (slot < 0) return;
    if (!capturing) cudaEventRecord(g_slots.ev[dev][slot], st);
    {
      std::lock_guard<std::mutex> lk(g_slots.m);
      g_slots.leased[dev][slot] = false;
    }
    g_slots.cv.notify_one();
  }
  SlotLease(const SlotLease &) = delete;
  SlotLease &operator=(const SlotLease &) = delete;
  bool capturing = false;
};

#define PDENET_DISPATCH_K(k, EXPR)                                                                                \
  switch (k) {                                                                                                    \
  case 3: {                                                                                                       \
    constexpr int KK_ = 3;                                                                                        \
    EXPR;                                                                                                         \
  } break;                                                                                                        \
  case 5: {                                                                                                       \
    constexpr int KK_ = 5;                                                                                        \
    EXPR;                                                                                                         \
  } break;                                                                                                        \
  case 7: {                                                                                                       \
    constexpr int KK_ = 7;                                                                                        \
    EXPR;                                                                                                         \
  } break;                                                                                                        \
  default:                                                                                                        \
    return PDENET_E_UNSUPPORTED;                                                                                  \
  }

static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

static int symnet_param_count(int m, int L) {
  int s = 0;
  for (int i = 0; i < L; ++i) s += 2 * (m + i) + 2;
  return s + (m + L) + 1;
}

static bool nc_supported(int N, int C) {
#define PDENET_CASE(NN, CC)                                                                                       \
  if (N == NN && C == CC) return true;
  PDENET_FOR_NC(PDENET_CASE)
#undef PDENET_CASE
  return false;
}
static bool lin_supported(int NF) {
#define PDENET_CASE(NN)                                                                                           \
  if (NF == NN) return true;
  PDENET_FOR_LIN(PDENET_CASE)
#undef PDENET_CASE
  return false;
}
// vector length of the SymNet-gradient warp reduction: needs M + L + 1 <= VL
static int pick_vl(int M, int L) {
  // the 16-long gradient vectors are only instantiated for banks with M + 1 < 16 features (inst.cu)
  if (M + L + 1 <= 16 && M + 1 < 16) return 16;
  if (M + L + 1 <= 32) return 32;
  return 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// moment <-> kernel map (rows a1-a3)
struct Mat7 {
  double m[PDENET_MAX_K * PDENET_MAX_K];
};
struct Scales {
  double s[PDENET_MAX_FILTERS];
};

// A[i][j] = (j - c)^i / i!   (row a1)
static void moment_matrix(int k, long double *A) {
  const int c = (k - 1) / 2;
  for (int i = 0; i < k; ++i) {
    long double fact = 1;
    for (int q = 2; q <= i; ++q) fact *= q;
    for (int j = 0; j < k; ++j) A[i * k + j] = powl((long double)(j - c), (long double)i) / fact;
  }
}
static void invert(int k, const long double *A, long double *Ai) {
  long double a[PDENET_MAX_K][2 * PDENET_MAX_K];
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      a[i][j] = A[i * k + j];
      a[i][k + j] = i == j ? 1.0L : 0.0L;
    }
  for (int col = 0; col < k; ++col) {
    int piv = col;
    for (int r = col + 1; r < k; ++r)
      if (fabsl(a[r][col]) > fabsl(a[piv][col])) piv = r;
    if (piv != col)
      for (int j = 0; j < 2 * k; ++j) {
        long double t = a[col][j];
        a[col][j] = a[piv][j];
        a[piv][j] = t;
      }
    const long double d = a[col][col];
    for (int j = 0; j < 2 * k; ++j) a[col][j] /= d;
    for (int r = 0; r < k; ++r)
      if (r != col) {
        const long double f = a[r][col];
        if (f != 0)
          for (int j = 0; j < 2 * k; ++j) a[r][j] -= f * a[col][j];
      }
  }
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) Ai[i * k + j] = a[i][k + j];
}

// out[n][x][y] = s[n] * sum_{i,j} Q[x][i] * in[n][i][j] * Q[y][j], optionally masked
template <typename T>
__global__ void sandwich_kernel(const T *in, T *out, int k, Mat7 Q, Scales sc, const unsigned char *mask) {
  __shared__ double tmp[PDENET_MAX_K * PDENET_MAX_K];
  const int n = blockIdx.x, t = threadIdx.x, kk = k * k;
  const int x = t / k, y = t % k;
  if (t < kk) {
    // tmp[x][j] = sum_i Q[x][i] in[i][j]   (here y plays j)
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += Q.m[x * k + i] * (double)in[n * kk + i * k + y];
    tmp[t] = s;
  }
  __syncthreads();
  if (t < kk) {
    double s = 0.0;
    for (int j = 0; j < k; ++j) s += tmp[x * k + j] * Q.m[y * k + j];
    s *= sc.s[n];
    if (mask && mask[n * kk + t]) s = 0.0;
    out[n * kk + t] = (T)s;
  }
}

enum { SW_M2K = 0, SW_K2M = 1, SW_M2K_BWD = 2 };

template <typename T>
static int sandwich(int which, const T *in, const double *scale_host, const unsigned char *mask, T *out, int N,
                    int k, void *stream) {
  if (!in || !out || N <= 0 || N > PDENET_MAX_FILTERS || k <= 0 || k > PDENET_MAX_K || (k % 2) == 0)
    return PDENET_E_BADARG;
  long double A[PDENET_MAX_K * PDENET_MAX_K], Ai[PDENET_MAX_K * PDENET_MAX_K];
  moment_matrix(k, A);
  invert(k, A, Ai);
  Mat7 Q;
  Scales sc;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      if (which == SW_M2K) Q.m[i * k + j] = (double)Ai[i * k + j];
      if (which == SW_K2M) Q.m[i * k + j] = (double)A[i * k + j];
      if (which == SW_M2K_BWD) Q.m[i * k + j] = (double)Ai[j * k + i];
    }
  for (int n = 0; n < N; ++n) {
    const double s = scale_host ? scale_host[n] : 1.0;
    sc.s[n] = which == SW_K2M ? 1.0 / s : s;
  }
  ProfScope ps(4, (cudaStream_t)stream);
  sandwich_kernel<T><<<N, 64, 0, (cudaStream_t)stream>>>(in, out, k, Q, sc, mask);
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---------------------------------------------------------------------------------------------
// fp32: second, filter-innermost copy of the bank (pairs of filters for the packed FFMA2 of the step kernels),
// built from the same sources in a per-slot device scratch and copied into the second half of the slot.
// Bank = [k0 (optional)] + n filters of `taps` + [2 twins (optional)].
__device__ __align__(16) unsigned char g_pair_scratch[kConstSlots][kPairedBytes];

__global__ void pack_pairs_kernel(const float *k0, const float *taps, int n, const float *twins, int kk, float *out) {
  const int nbank = (k0 ? 1 : 0) + n + (twins ? 2 : 0), np2 = (nbank + 1) / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kk * np2 * 2) return;
  const int h = idx & 1, np = (idx >> 1) % np2, e = (idx >> 1) / np2, f = 2 * np + h;
  float v = 0.f;
  if (f < nbank) {
    int g = f;
    if (k0) {
      if (g == 0) {
        v = k0[e];
        g = -1;
      } else {
        g -= 1;
      }
    }
    if (g >= 0) v = g < n ? taps[g * kk + e] : twins[(g - n) * kk + e];
  }
  out[idx] = v;
}

template <typename T>
static int upload_pairs(const SlotLease &lease, const T *k0, const T *taps, int n, const T *twins, int k, cudaStream_t st) {
  if constexpr (sizeof(T) == 4) {
    void *scr = nullptr;
    if (cudaGetSymbolAddress(&scr, g_pair_scratch) != cudaSuccess) return PDENET_E_CUDA;
    float *out = (float *)((unsigned char *)scr + (size_t)lease.slot * kPairedBytes);
    const int nbank = (k0 ? 1 : 0) + n + (twins ? 2 : 0), cnt = k * k * ((nbank + 1) / 2) * 2;
    if ((size_t)cnt * sizeof(float) > (size_t)kPairedBytes) return PDENET_E_UNSUPPORTED;
    pack_pairs_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(k0, taps, n, twins, k * k, out);
    if (cudaGetLastError() != cudaSuccess) return PDENET_E_CUDA;
    int rc = 0;
    PDENET_DISPATCH_K(k, rc = (upload_raw<T, KK_>(lease.slot, kPairedBytes, out, (size_t)cnt * sizeof(float), st)));
    return rc;
  } else {
    return 0;
  }
}

// upload [n][k][k] taps (optionally preceded by one extra filter k0) into a leased constant slot
template <typename T>
static int upload(const SlotLease &lease, const T *k0, const T *taps, int n, int k, cudaStream_t st) {
  if (lease.slot < 0) return PDENET_E_CUDA;
  int rc = 0, first = 0;
  if (k0) {
    PDENET_DISPATCH_K(k, rc = (upload_taps<T, KK_>(lease.slot, 0, k0, 1, st)));
    if (rc) return rc;
    first = 1;
  }
  PDENET_DISPATCH_K(k, rc = (upload_taps<T, KK_>(lease.slot, first, taps, n, st)));
  if (rc) return rc;
  return upload_pairs<T>(lease, k0, taps, n, nullptr, k, st);
}

// ---- pseudo-upwind bank (row f-1): bank order puts order (0,1) at index 1 (derivative along W) and
// (1,0) at index 2 (along H); their flipped twins  K1-[a,b] = -K1[a,k-1-b],  K2-[a,b] = -K2[k-1-a,b]
// (mirror along the derivative axis, negate: the same derivative with the opposite bias) follow the
// N central filters.  The twins of a call are built in a per-slot device scratch and uploaded from there.
__device__ __align__(16) unsigned char g_upwind_scratch[kConstSlots][2 * PDENET_MAX_K * PDENET_MAX_K * 8];

template <typename T> __global__ void upwind_twins_kernel(const T *kernels, int k, T *twins) {
  const int t = threadIdx.x, kk = k * k;
  if (t >= 2 * kk) return;
  const int which = t / kk, e = t % kk, a = e / k, b = e % k;
  twins[t] = which == 0 ? -kernels[1 * kk + a * k + (k - 1 - b)] : -kernels[2 * kk + (k - 1 - a) * k + b];
}
// grad[n] = dk[n] (n < N) minus the mirrored gradient of the twin for n = 1, 2
template <typename T> __global__ void upwind_fold_kernel(const T *dk_bank, int N, int k, T *grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, kk = k * k;
  if (i >= N * kk) return;
  const int n = i / kk, e = i % kk, a = e / k, b = e % k;
  T v = dk_bank[i];
  if (n == 1) v -= dk_bank[N * kk + a * k + (k - 1 - b)];
  if (n == 2) v -= dk_bank[(N + 1) * kk + (k - 1 - a) * k + b];
  grad[i] = v;
}
static bool upwind_supported(int N, int C) { return (N == 3 || N == 6) && (C == 1 || C == 2); }

// upload the N central filters followed by the two twins into the leased slot
template <typename T>
static int upload_upwind(const SlotLease &lease, const T *kernels, int N, int k, cudaStream_t st) {
  if (lease.slot < 0) return PDENET_E_CUDA;
  void *scr = nullptr;
  if (cudaGetSymbolAddress(&scr, g_upwind_scratch) != cudaSuccess) return PDENET_E_CUDA;
  T *twins = (T *)((unsigned char *)scr + (size_t)lease.slot * sizeof(g_upwind_scratch[0]));
  upwind_twins_kernel<T><<<1, 128, 0, st>>>(kernels, k, twins);
  if (cudaGetLastError() != cudaSuccess) return PDENET_E_CUDA;
  int rc = 0;
  PDENET_DISPATCH_K(k, rc = (upload_taps<T, KK_>(lease.slot, 0, kernels, N, st)));
  if (rc) return rc;
  PDENET_DISPATCH_K(k, rc = (upload_taps<T, KK_>(lease.slot, N, twins, 2, st)));
  if (rc) return rc;
  return upload_pairs<T>(lease, nullptr, kernels, N, twins, k, st);
}

static bool bad_geom(int B, int C, int H, int W, int k, int N) {
  return B <= 0 || C <= 0 || H <= 0 || W <= 0 || k <= 0 || (k % 2) == 0 || N <= 0;
}

template <typename T, int K>
static int run_step(int mode, int N, int C, int VL, StepParams<T> &p, int B, int H, int W, cudaStream_t st) {
  p.B = B;
  p.H = H;
  p.W = W;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  ProfScope ps((mode == MODE_BWD_A || mode_is_kept(mode) || mode == MODE_LIN_BWD_A || mode == MODE_UPW_BWD_A) ? 1
                                                                                                                  : 0,
               st);
  return launch_step<T, K>(mode, N, C, VL, p, sms, st);
}

template <typename T, int K>
static int run_adjoint(AdjParams<T> &p, int B, int C, int N, int H, int W, cudaStream_t st) {
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  int tw, th;
  adjoint_tile<T, K>(0, tw, th);
  p.B = B;
  p.C = C;
  p.N = N;
  p.H = H;
  p.W = W;
  // small (16-row) tiles when the 64-row tiles would be too few to fill / balance the SMs
  const long long big_items = (long long)B * C * ((W + tw - 1) / tw) * ((H + th - 1) / th);
  const int small = p.wrap_y && !p.mirror_rows && big_items < (long long)small_tile_factor() * sms * 2;
  if (small) adjoint_tile<T, K>(1, tw, th);
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  p.n_items = B * C * p.tiles_x * p.tiles_y;
  const int vec = 16 / (int)sizeof(T);
  p.vec_ok = (W % vec == 0) && (((uintptr_t)p.g) % 16 == 0) ? 1 : 0;
  const int cap = sms * (small ? 8 : (kAdjNT == 512 ? 1 : 2)); // resident CTAs per SM
  const int grid = p.n_items < cap ? p.n_items : cap;
  ProfScope ps(2, st);
  return launch_adjoint<T, K>(p, grid, small, st);
}

// Filter gradient: warps own filters; N filters are covered by launches of nw in {6, 3, 2, 1}
// filters each (nw divides the warps of a CTA; the other warps split the tile rows).  Geometry
// (kFGradWarps, kFGradCtasPerSm): params.cuh.
static int fgrad_nw(int left) { // largest divisor of the CTA's warp count that does not exceed the filters left
  int nw = 1;
  for (int d = 1; d <= kFGradWarps; ++d)
    if (kFGradWarps % d == 0 && d <= left) nw = d;
  return nw;
}
// row parts per CTA of the launch that splits the rows most finely (= partial slots a CTA needs): the partial
// buffer and its final reduction cover sms * CTAs * this many slots, not one per warp (N = 6: 2 instead of 12 --
// the reduction of a small problem was its longest kernel)
static int fgrad_max_parts(int N) {
  int mp = 1;
  for (int n0 = 0; n0 < N;) {
    const int nw = fgrad_nw(N - n0);
    if (kFGradWarps / nw > mp) mp = kFGradWarps / nw;
    n0 += nw;
  }
  return mp;
}
static int dk_slot_count(int sms, int N) { return sms * kFGradCtasPerSm * fgrad_max_parts(N); }

template <typename T, int K>
static int run_filter_grad(FGradParams<T> &p, int B, int C, int N, int H, int W, cudaStream_t st) {
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  int tw, th;
  filter_grad_tile<T, K>(0, tw, th);
  p.B = B;
  p.C = C;
  p.N = N;
  p.H = H;
  p.W = W;
  const long long big_items = (long long)B * C * ((W + tw - 1) / tw) * ((H + th - 1) / th);
  const long long enough = (long long)small_tile_factor() * sms * kFGradCtasPerSm;
  int small = 0; // tile level: 0 large, 2 medium, 1 small -- the largest tile that still gives `enough` items
  if (p.wrap_y && big_items < enough) {
    filter_grad_tile<T, K>(2, tw, th);
    const long long med_items = (long long)B * C * ((W + tw - 1) / tw) * ((H + th - 1) / th);
    small = med_items >= enough ? 2 : 1;
  }
  if (p.wrap_y && filter_grad_tile_level() >= 0) small = filter_grad_tile_level();
  filter_grad_tile<T, K>(small, tw, th);
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  p.n_items = B * C * p.tiles_x * p.tiles_y;
  const int vec = 16 / (int)sizeof(T);
  p.vec_ok = (W % vec == 0) && (((uintptr_t)p.g) % 16 == 0) ? 1 : 0;
  const int cap = sms * kFGradCtasPerSm;
  const int grid = p.n_items < cap ? p.n_items : cap;
  p.slot_stride = fgrad_max_parts(N);
  int n0 = 0;
  while (n0 < N) {
    const int nw = fgrad_nw(N - n0);
    p.n0 = n0;
    p.nw = nw;
    ProfScope ps(3, st);
    const int rc = launch_filter_grad<T, K>(p, grid, small, st);
    if (rc) return rc;
    n0 += nw;
  }
  return 0;
}

// ---- forward rollout (rows a5-a10) -----------------------------------------------------------
template <typename T>
static int rollout_fwd(const T *u0, const T *kernels, const T *symnet, T *states, T *feats, int B, int C, int H,
                       int W, int k, int N, int L, int steps, double dt, void *stream, int rhs_only = 0,
                       bool upwind = false) {
  if (!u0 || !kernels || !symnet || !states || bad_geom(B, C, H, W, k, N) || L < 0 || steps < 0)
    return PDENET_E_BADARG;
  if (L > PDENET_MAX_LAYERS || !nc_supported(N, C) || (upwind && (!upwind_supported(N, C) || feats)))
    return PDENET_E_UNSUPPORTED;
  const size_t stride = (size_t)B * C * H * W;
  SlotLease lease((cudaStream_t)stream);
  if (upwind) {
    if (int rc = upload_upwind<T>(lease, kernels, N, k, (cudaStream_t)stream)) return rc;
  } else if (int rc = upload<T>(lease, nullptr, kernels, N, k, (cudaStream_t)stream)) {
    return rc;
  }
  StepParams<T> p;
  memset(&p, 0, sizeof(p));
  p.slot = lease.slot;
  p.symnet = symnet;
  p.L = L;
  p.Psym = symnet_param_count(C * N, L);
  p.dt = (T)dt;
  p.wrap_y = 1;
  p.u_rows = p.out_rows = H;
  p.rhs_only = rhs_only;
  for (int t = 0; t < steps; ++t) {
    p.u = t == 0 ? u0 : states + (size_t)(t - 1) * stride;
    p.out = states + (size_t)t * stride;
    p.feats = feats ? feats + (size_t)t * stride * N : nullptr; // features of the step's input state u_t
    int rc = 0;
    if (upwind) {
      PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_UPW_FWD, N + 2, C, 16, p, B, H, W, (cudaStream_t)stream)));
    } else {
      PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_FWD, N, C, 16, p, B, H, W, (cudaStream_t)stream)));
    }
    if (rc) return rc;
  }
  return PDENET_OK;
}

template <typename T>
static int fd2d_fwd(const T *u, const T *kernels, T *feats, int B, int C, int H, int W, int k, int N, void *stream) {
  if (!u || !kernels || !feats || bad_geom(B, C, H, W, k, N)) return PDENET_E_BADARG;
  if (!nc_supported(N, C)) return PDENET_E_UNSUPPORTED;
  SlotLease lease((cudaStream_t)stream);
  if (int rc = upload<T>(lease, nullptr, kernels, N, k, (cudaStream_t)stream)) return rc;
  StepParams<T> p;
  memset(&p, 0, sizeof(p));
  p.slot = lease.slot;
  p.symnet = kernels; // unused in MODE_FEATS (Psym = 0)
  p.wrap_y = 1;
  p.u_rows = p.out_rows = H;
  p.u = u;
  p.out = feats;
  int rc = 0;
  PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_FEATS, N, C, 16, p, B, H, W, (cudaStream_t)stream)));
  return rc;
}

template <typename T>
static int fd2d_bwd(const T *u, const T *grad_feats, const T *kernels, T *grad_u, T *grad_kernels, void *workspace,
                    size_t ws_bytes, int B, int C, int H, int W, int k, int N, void *stream) {
  if (!u || !grad_feats || !kernels || !workspace || bad_geom(B, C, H, W, k, N) || N > PDENET_MAX_FILTERS)
    return PDENET_E_BADARG;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  const int kk = k * k, vlb = kk <= 16 ? 16 : (kk <= 32 ? 32 : 64), slots = dk_slot_count(sms, N);
  const size_t need = (size_t)slots * N * vlb * sizeof(double);
  if (ws_bytes < need) return PDENET_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double *dkp = (double *)workspace;
  if (grad_kernels && cudaMemsetAsync(dkp, 0, need, st) != cudaSuccess) return PDENET_E_CUDA;
  SlotLease lease(st);
  if (int rc = upload<T>(lease, nullptr, kernels, N, k, st)) return rc;
  AdjParams<T> q;
  memset(&q, 0, sizeof(q));
  q.g = grad_feats;
  q.g_rows = H;
  q.wrap_y = 1;
  q.lam_out = grad_u;
  q.lo_rows = H;
  q.slot = lease.slot;
  int rc = 0;
  if (grad_u) {
    PDENET_DISPATCH_K(k, rc = (run_adjoint<T, KK_>(q, B, C, N, H, W, st)));
    if (rc) return rc;
  }
  if (grad_kernels) {
    FGradParams<T> f;
    memset(&f, 0, sizeof(f));
    f.g = grad_feats;
    f.g_rows = H;
    f.u = u;
    f.u_rows = H;
    f.wrap_y = 1;
    f.dk_partial = dkp;
    PDENET_DISPATCH_K(k, rc = (run_filter_grad<T, KK_>(f, B, C, N, H, W, st)));
    if (rc) return rc;
    const int n = N * kk;
    ProfScope ps(4, st);
    reduce_dk_kernel<T><<<reduce_blocks(n), 128, 0, st>>>(dkp, slots, N, kk, vlb, n, grad_kernels);
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

template <typename T>
static int slab_step_fwd(const T *u_ext, const T *kernels, const T *symnet, T *out, T *feats, T *peer_up, T *peer_dn,
                         int mirror_rows, int B, int C, int Hs, int W, int halo, int out_rows, int out_off, int k,
                         int N, int L, double dt, void *stream) {
  if (mirror_rows < 0 || mirror_rows > Hs || mirror_rows > out_off || (mirror_rows > 0 && (!peer_up || !peer_dn)) ||
      out_rows < Hs + out_off + mirror_rows)
    return PDENET_E_BADARG;
  if (!u_ext || !kernels || !symnet || !out || bad_geom(B, C, Hs, W, k, N) || L < 0 || halo < (k - 1) / 2 ||
      out_rows < Hs + out_off || out_off < 0)
    return PDENET_E_BADARG;
  if (L > PDENET_MAX_LAYERS || !nc_supported(N, C)) return PDENET_E_UNSUPPORTED;
  SlotLease lease((cudaStream_t)stream);
  if (int rc = upload<T>(lease, nullptr, kernels, N, k, (cudaStream_t)stream)) return rc;
  StepParams<T> p;
  memset(&p, 0, sizeof(p));
  p.slot = lease.slot;
  p.symnet = symnet;
  p.L = L;
  p.Psym = symnet_param_count(C * N, L);
  p.dt = (T)dt;
  p.wrap_y = 0;
  p.u = u_ext;
  p.u_rows = Hs + 2 * halo;
  p.u_off = halo;
  p.out = out;
  p.out_rows = out_rows;
  p.out_off = out_off;
  p.peer_up = peer_up;
  p.peer_dn = peer_dn;
  p.mirror_rows = mirror_rows;
  p.feats = feats; // [B][C*N][Hs][W]: features of the owned rows, kept for the reverse step (may be NULL)
  int rc = 0;
  PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_FWD, N, C, 16, p, B, Hs, W, (cudaStream_t)stream)));
  return rc;
}

// ---- reverse slab step (config 5) ------------------------------------------------
struct SlabLayout {
  size_t off_g, off_dk, off_sym, total;
  int dk_slots, sym_slots, vlb;
};
static SlabLayout slab_layout(size_t es, int B, int C, int Hs, int W, int k, int N, int L, int sms) {
  SlabLayout l;
  const int kk = k * k, r = (k - 1) / 2;
  l.vlb = kk <= 16 ? 16 : (kk <= 32 ? 32 : 64);
  l.dk_slots = dk_slot_count(sms, N);
  l.sym_slots = sms * kSymSlotsPerSm;
  size_t o = 0;
  l.off_g = o;
  o = align_up(o + (size_t)B * C * N * (Hs + 2 * r) * W * es, 256);
  l.off_dk = o;
  o = align_up(o + (size_t)l.dk_slots * N * l.vlb * sizeof(double), 256);
  l.off_sym = o;
  o = align_up(o + (size_t)l.sym_slots * C * symnet_param_count(C * N, L) * sizeof(double), 256);
  l.total = o;
  return l;
}

template <typename T>
static int slab_step_bwd(const T *u_ext, const T *lam_ext, const T *feats, const T *kernels, const T *symnet,
                         T *lam_out, T *peer_up, T *peer_dn, int mirror_rows, int out_rows, int out_off, void *workspace, size_t ws_bytes,
                         int B, int C, int Hs, int W, int k, int N, int L, double dt, void *stream) {
  if (mirror_rows < 0 || mirror_rows > Hs || mirror_rows > out_off || (mirror_rows > 0 && (!peer_up || !peer_dn)) ||
      out_rows < Hs + out_off + mirror_rows)
    return PDENET_E_BADARG;
  if (!u_ext || !lam_ext || !kernels || !symnet || !lam_out || !workspace || bad_geom(B, C, Hs, W, k, N) || L < 0 ||
      out_off < 0 || out_rows < Hs + out_off)
    return PDENET_E_BADARG;
  if (L > PDENET_MAX_LAYERS || !nc_supported(N, C)) return PDENET_E_UNSUPPORTED;
  const int M = C * N, VL = pick_vl(M, L), r = (k - 1) / 2;
  if (!VL) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const SlabLayout lay = slab_layout(sizeof(T), B, C, Hs, W, k, N, L, sms);
  if (ws_bytes < lay.total) return PDENET_E_WORKSPACE;
  char *ws = (char *)workspace;
  T *gbuf = (T *)(ws + lay.off_g);
  SlotLease lease(st);
  if (int rc = upload<T>(lease, nullptr, kernels, N, k, st)) return rc;
  // kernel A on the owned rows (with parameter gradients) and on the r ghost rows above / below
  // (g only): B1 needs g on [-r, Hs + r)
  const int row0[3] = {0, -r, Hs}, nrows[3] = {Hs, r, r};
  for (int part = 0; part < 3; ++part) {
    if (nrows[part] <= 0) continue;
    StepParams<T> p;
    memset(&p, 0, sizeof(p));
    p.slot = lease.slot;
    p.symnet = symnet;
    p.L = L;
    p.Psym = symnet_param_count(M, L);
    p.dt = (T)dt;
    p.wrap_y = 0;
    p.u = u_ext;
    p.u_rows = Hs + 4 * r;
    p.u_off = 2 * r + row0[part];
    p.lam = lam_ext;
    p.lam_rows = Hs + 2 * r;
    p.lam_off = r + row0[part];
    p.g = gbuf;
    p.g_rows = Hs + 2 * r;
    p.g_off = r + row0[part];
    p.sym_partial = (double *)(ws + lay.off_sym);
    p.skip_sym_grad = part != 0;
    // owned rows: read the features the forward step kept, if any; ghost rows are always recomputed
    const bool kept = part == 0 && feats;
    p.feats = kept ? const_cast<T *>(feats) : nullptr;
    int rc = 0;
    PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(kept ? MODE_BWD_A_KEPT : MODE_BWD_A, N, C, VL, p, B, nrows[part], W, st)));
    if (rc) return rc;
  }
  AdjParams<T> q;
  memset(&q, 0, sizeof(q));
  q.g = gbuf;
  q.g_rows = Hs + 2 * r;
  q.g_off = r;
  q.wrap_y = 0;
  q.lam_in = lam_ext;
  q.li_rows = Hs + 2 * r;
  q.li_off = r;
  q.lam_out = lam_out;
  q.lo_rows = out_rows;
  q.lo_off = out_off;
  q.peer_up = peer_up;
  q.peer_dn = peer_dn;
  q.mirror_rows = mirror_rows;
  q.slot = lease.slot;
  int rc = 0;
  PDENET_DISPATCH_K(k, rc = (run_adjoint<T, KK_>(q, B, C, N, Hs, W, st)));
  if (rc) return rc;
  FGradParams<T> f;
  memset(&f, 0, sizeof(f));
  f.g = gbuf;
  f.g_rows = Hs + 2 * r;
  f.g_off = r;
  f.u = u_ext;
  f.u_rows = Hs + 4 * r;
  f.u_off = 2 * r;
  f.wrap_y = 0;
  f.dk_partial = (double *)(ws + lay.off_dk);
  PDENET_DISPATCH_K(k, rc = (run_filter_grad<T, KK_>(f, B, C, N, Hs, W, st)));
  return rc;
}

template <typename T>
static int slab_bwd_end(void *workspace, size_t ws_bytes, T *grad_kernels, T *grad_symnet, int B, int C, int Hs,
                        int W, int k, int N, int L, void *stream) {
  if (!workspace || bad_geom(B, C, Hs, W, k, N) || L < 0) return PDENET_E_BADARG;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const SlabLayout lay = slab_layout(sizeof(T), B, C, Hs, W, k, N, L, sms);
  if (ws_bytes < lay.total) return PDENET_E_WORKSPACE;
  char *ws = (char *)workspace;
  if (grad_kernels) {
    const int n = N * k * k;
    ProfScope ps(4, st);
    reduce_dk_kernel<T><<<reduce_blocks(n), 128, 0, st>>>((double *)(ws + lay.off_dk), lay.dk_slots, N, k * k, lay.vlb,
                                                         n, grad_kernels);
  }
  if (grad_symnet) {
    const int n = C * symnet_param_count(C * N, L);
    ProfScope ps(4, st);
    reduce_flat_kernel<T><<<reduce_blocks(n), 128, 0, st>>>((double *)(ws + lay.off_sym), lay.sym_slots, n, grad_symnet);
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- reverse rollout (row a11) ---------------------------------------------------------------
struct BwdLayout {
  size_t off_g, off_lam0, off_lam1, off_dk, off_sym, off_bank, total;
  int chunk, dk_slots, sym_slots, vlb;
};

static BwdLayout bwd_layout(size_t es, int B, int C, int H, int W, int k, int N, int L, int batch_chunk, int sms) {
  BwdLayout l;
  l.chunk = (batch_chunk <= 0 || batch_chunk > B) ? B : batch_chunk;
  const int kk = k * k;
  l.vlb = kk <= 16 ? 16 : (kk <= 32 ? 32 : 64);
  l.dk_slots = dk_slot_count(sms, N);
  l.sym_slots = sms * kSymSlotsPerSm;
  const size_t img = (size_t)H * W;
  size_t o = 0;
  l.off_g = o;
  o = align_up(o + (size_t)l.chunk * C * N * img * es, 256);
  l.off_lam0 = o;
  o = align_up(o + (size_t)B * C * img * es, 256);
  l.off_lam1 = o;
  o = align_up(o + (size_t)B * C * img * es, 256);
  l.off_dk = o;
  o = align_up(o + (size_t)l.dk_slots * N * l.vlb * sizeof(double), 256);
  l.off_sym = o;
  o = align_up(o + (size_t)l.sym_slots * C * symnet_param_count(C * N, L) * sizeof(double), 256);
  l.off_bank = o; // upwind: gradient of the whole bank (N filters incl. the twins) before it is folded
  o = align_up(o + (size_t)N * kk * es, 256);
  l.total = o;
  return l;
}

// upwind: g / dK regions for the bank of Ns + 2 filters, SymNet partials for C*Ns inputs
static BwdLayout bwd_layout_upwind(size_t es, int B, int C, int H, int W, int k, int Ns, int L, int batch_chunk,
                                   int sms) {
  BwdLayout l = bwd_layout(es, B, C, H, W, k, Ns + 2, L, batch_chunk, sms);
  // the SymNet region of the wide layout is an upper bound of what C*Ns inputs need: keep it
  return l;
}

template <typename T>
static int rollout_bwd(const T *u0, const T *states, const T *feats, const T *grad_states, int grad_mode,
                       const T *kernels, const T *symnet, T *grad_u0, T *grad_kernels, T *grad_symnet, void *workspace,
                       size_t ws_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L, int steps,
                       double dt, void *stream, bool identity = true, bool upwind = false) {
  if (!u0 || !grad_states || !kernels || !symnet || !workspace || bad_geom(B, C, H, W, k, N) || L < 0 ||
      steps < 1 || (steps > 1 && !states) || (grad_mode != PDENET_GRAD_LAST && grad_mode != PDENET_GRAD_ALL))
    return PDENET_E_BADARG;
  if (L > PDENET_MAX_LAYERS || !nc_supported(N, C) || (upwind && (!upwind_supported(N, C) || feats)))
    return PDENET_E_UNSUPPORTED;
  const int Ns = N;            // filters feeding the SymNet
  if (upwind) N += 2;          // the bank the stencil / data-gradient / filter-gradient kernels see
  const int M = C * N;
  const int VL = pick_vl(C * Ns, L);
  if (!VL) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  // (upwind: sized for the bank of Ns + 2 filters == bwd_layout_upwind(Ns))
  const BwdLayout lay = bwd_layout(sizeof(T), B, C, H, W, k, N, L, batch_chunk, sms);
  if (ws_bytes < lay.total) return PDENET_E_WORKSPACE;
  char *ws = (char *)workspace;
  T *gbuf = (T *)(ws + lay.off_g);
  T *lamb[2] = {(T *)(ws + lay.off_lam0), (T *)(ws + lay.off_lam1)};
  double *dkp = (double *)(ws + lay.off_dk);
  double *symp = (double *)(ws + lay.off_sym);
  const int Psym = symnet_param_count(C * Ns, L);
  // the two partial regions are adjacent in the workspace (bwd_layout): one memset node for both
  if (cudaMemsetAsync(dkp, 0, (lay.off_sym - lay.off_dk) + (size_t)lay.sym_slots * C * Psym * sizeof(double), st) !=
      cudaSuccess)
    return PDENET_E_CUDA;

  SlotLease lease(st);
  if (upwind) {
    if (int rc = upload_upwind<T>(lease, kernels, Ns, k, st)) return rc;
  } else if (int rc = upload<T>(lease, nullptr, kernels, N, k, st)) {
    return rc;
  }
  const size_t img = (size_t)H * W;
  const size_t stride = (size_t)B * C * img;
  const T *lam_in = grad_mode == PDENET_GRAD_ALL ? grad_states + (size_t)(steps - 1) * stride : grad_states;
  int cur = 0;
  for (int t = steps - 1; t >= 0; --t) {
    const T *ut = t == 0 ? u0 : states + (size_t)(t - 1) * stride;
    T *lam_out = t == 0 ? grad_u0 : lamb[cur];
    const T *inject = (grad_mode == PDENET_GRAD_ALL && t > 0) ? grad_states + (size_t)(t - 1) * stride : nullptr;
    for (int b0 = 0; b0 < B; b0 += lay.chunk) {
      const int nb = (B - b0) < lay.chunk ? (B - b0) : lay.chunk;
      const size_t boff = (size_t)b0 * C * img;
      StepParams<T> p;
      memset(&p, 0, sizeof(p));
      p.slot = lease.slot;
      p.symnet = symnet;
      p.L = L;
      p.Psym = Psym;
      p.dt = (T)dt;
      p.wrap_y = 1;
      p.u = ut + boff;
      p.u_rows = H;
      p.lam = lam_in + boff;
      p.lam_rows = H;
      p.g = gbuf;
      p.g_rows = H;
      p.sym_partial = symp;
      // kept features of u_t (forward ran with a feature buffer): kernel A skips the stencils
      p.feats = feats ? const_cast<T *>(feats) + ((size_t)t * B + b0) * M * img : nullptr;
      int rc = 0;
      PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(upwind ? MODE_UPW_BWD_A : (feats ? MODE_BWD_A_KEPT : MODE_BWD_A), N, C,
                                                  VL, p, nb, H, W, st)));
      if (rc) return rc;
      if (lam_out) {
        AdjParams<T> q;
        memset(&q, 0, sizeof(q));
        q.g = gbuf;
        q.g_rows = H;
        q.wrap_y = 1;
        q.lam_in = identity ? lam_in + boff : nullptr; // F(u) alone (row a8) has no identity path
        q.li_rows = H;
        q.inject = inject ? inject + boff : nullptr;
        q.inj_rows = H;
        q.lam_out = lam_out + boff;
        q.lo_rows = H;
        q.slot = lease.slot;
        PDENET_DISPATCH_K(k, rc = (run_adjoint<T, KK_>(q, nb, C, N, H, W, st)));
        if (rc) return rc;
      }
      if (grad_kernels) {
        FGradParams<T> f;
        memset(&f, 0, sizeof(f));
        f.g = gbuf;
        f.g_rows = H;
        f.u = ut + boff;
        f.u_rows = H;
        f.wrap_y = 1;
        f.dk_partial = dkp;
        PDENET_DISPATCH_K(k, rc = (run_filter_grad<T, KK_>(f, nb, C, N, H, W, st)));
        if (rc) return rc;
      }
    }
    lam_in = lam_out;
    cur ^= 1;
    if (!lam_out && t > 0) return PDENET_E_BADARG; // unreachable: lam_out is only NULL at t == 0
  }
  if (grad_kernels && grad_symnet && !upwind) { // the usual case: both finishing reductions in one launch
    const int n = N * k * k, ns = C * Psym;
    ProfScope ps(4, st);
    reduce_grads_kernel<T><<<reduce_blocks(n + ns), 128, 0, st>>>(dkp, lay.dk_slots, N, k * k, lay.vlb, n, grad_kernels,
                                                                  symp, lay.sym_slots, ns, grad_symnet);
    return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
  }
  if (grad_kernels) {
    const int n = N * k * k;
    ProfScope ps(4, st);
    T *bank = upwind ? (T *)(ws + lay.off_bank) : grad_kernels;
    reduce_dk_kernel<T><<<reduce_blocks(n), 128, 0, st>>>(dkp, lay.dk_slots, N, k * k, lay.vlb, n, bank);
    if (upwind) // dK of a twin is the mirrored, negated contribution to its central filter
      upwind_fold_kernel<T><<<(Ns * k * k + 127) / 128, 128, 0, st>>>(bank, Ns, k, grad_kernels);
  }
  if (grad_symnet) {
    const int n = C * Psym;
    ProfScope ps(4, st);
    reduce_flat_kernel<T><<<reduce_blocks(n), 128, 0, st>>>(symp, lay.sym_slots, n, grad_symnet);
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- PDE-Net 1 linear rollout (row a12) ------------------------------------------------------
// Small grids: one trajectory per thread-block cluster, resident in shared memory over all steps
// (linear_cluster_kernels.cuh).  lin_cluster_plan returns the CTAs per cluster (1 .. 8) or 0 when the
// grid does not fit (then the tiled kernels run).
static std::atomic<int> g_linear_cluster_on{1};

struct LinClusterPlan {
  int cs, pts, hb_max;
  size_t smem;
};
// Clusters of `cs` CTAs of the given kernel that the device holds at once (cudaOccupancyMaxActiveClusters: the
// CTAs of a cluster must share a GPC, so e.g. only 15 clusters of 8 fit a B200 although 15 * 8 < 148), or 0 if the
// query fails.  Cached per (scalar size, k, points per thread, direction, cluster size, shared memory).
template <typename KernelT> static int query_active_clusters(KernelT kern, int cs, size_t smem) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cs);
  cfg.blockDim = dim3(kLinClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
template <typename T> static int active_clusters_t(int k, int pts, int bwd, int cs, size_t smem) {
#define PDENET_OCC(KK_, PTS_)                                                                                     \
  if (k == KK_ && pts == PTS_)                                                                                    \
    return bwd ? query_active_clusters(linear_cluster_bwd_kernel<T, KK_, PTS_>, cs, smem)                         \
               : query_active_clusters(linear_cluster_fwd_kernel<T, KK_, PTS_>, cs, smem);
  PDENET_OCC(3, 1) PDENET_OCC(3, 2) PDENET_OCC(5, 1) PDENET_OCC(5, 2) PDENET_OCC(7, 1)
#undef PDENET_OCC
  return 0;
}
static int active_clusters(size_t es, int k, int pts, int bwd, int cs, size_t smem) {
  static std::mutex mu;
  static std::map<std::tuple<size_t, int, int, int, int, size_t, int>, int> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const auto key = std::make_tuple(es, k, pts, bwd, cs, smem, dev);
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int n = es == 8 ? active_clusters_t<double>(k, pts, bwd, cs, smem) : active_clusters_t<float>(k, pts, bwd, cs, smem);
  cache[key] = n;
  return n;
}

// Cluster size (1..8 CTAs per trajectory, any integer: bands need not be equal) for the small-grid path.  Every size
// whose band fits a CTA (<= 256 * pts points, <= 220 KB, at least r rows) is a candidate; the choice minimises
//   waves * (2 + points per thread),   waves = ceil(B / clusters the device holds at once),
// a model of a launch that is all latency (per step: one cluster barrier + one dependent tap chain per point of a
// thread), ties to the larger cluster.  BASELINE config 2 (B = 16, 50 x 50, 5 x 5): clusters of 8 run as TWO waves
// (15 resident + 1 straggler: measured 81 us forward, the SMs active for 40 of them; 7 do not all fit either),
// clusters of 6 as one (49 us).
static LinClusterPlan lin_cluster_plan(size_t es, int B, int H, int W, int k, int N, int sms, int bwd = 0) {
  LinClusterPlan none = {0, 0, 0, 0};
  if (!g_linear_cluster_on.load(std::memory_order_relaxed) || (k != 3 && k != 5 && k != 7)) return none;
  const int r = (k - 1) / 2, kk = k * k, pts_max = k == 7 ? 1 : 2;
  auto fits = [&](int cs, LinClusterPlan &pl) {
    if (H / cs < r || H / cs < 1) return false; // every band must own the halo rows its neighbours need
    const int hb_max = (H + cs - 1) / cs;
    if ((long long)hb_max * W > (long long)kLinClusterThreads * pts_max) return false;
    const size_t elems = (size_t)(N + 1) * kk + (size_t)(hb_max + 2 * r) * kk * W + 4 * (size_t)(hb_max + 2 * r) * W;
    if (elems * es > (size_t)220 * 1024) return false;
    pl.cs = cs;
    pl.hb_max = hb_max;
    pl.pts = (hb_max * W + kLinClusterThreads - 1) / kLinClusterThreads;
    pl.smem = elems * es;
    return true;
  };
  LinClusterPlan best = none, cand;
  long long best_cost = 0;
  for (int cs = 1; cs <= 8; ++cs)
    if (fits(cs, cand)) {
      int n = active_clusters(es, k, cand.pts, bwd, cs, cand.smem);
      if (n <= 0) n = sms / cs > 0 ? sms / cs : 1; // no device / query failed: one CTA per SM, GPC limits ignored
      const long long waves = ((long long)B + n - 1) / n, cost = waves * (2 + cand.pts);
      if (!best.cs || cost <= best_cost) {
        best = cand;
        best_cost = cost;
      }
    }
  return best;
}

template <typename KernelT, typename ParamsT>
static int launch_cluster(KernelT kern, const ParamsT &p, int grid, int cs, size_t smem, cudaStream_t st) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return PDENET_E_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kLinClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p) == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

#define PDENET_LIN_CLUSTER_DISPATCH(KERNEL, k, pts, ...)                                                          \
  do {                                                                                                            \
    if (k == 3 && pts == 1) return launch_cluster(KERNEL<T, 3, 1>, __VA_ARGS__);                                  \
    if (k == 3 && pts == 2) return launch_cluster(KERNEL<T, 3, 2>, __VA_ARGS__);                                  \
    if (k == 5 && pts == 1) return launch_cluster(KERNEL<T, 5, 1>, __VA_ARGS__);                                  \
    if (k == 5 && pts == 2) return launch_cluster(KERNEL<T, 5, 2>, __VA_ARGS__);                                  \
    if (k == 7 && pts == 1) return launch_cluster(KERNEL<T, 7, 1>, __VA_ARGS__);                                  \
    return PDENET_E_UNSUPPORTED;                                                                                  \
  } while (0)

template <typename T>
static int lin_cluster_fwd_launch(const LinClusterParams<T> &p, int k, const LinClusterPlan &pl, cudaStream_t st) {
  PDENET_LIN_CLUSTER_DISPATCH(linear_cluster_fwd_kernel, k, pl.pts, p, p.B * pl.cs, pl.cs, pl.smem, st);
}
template <typename T>
static int lin_cluster_bwd_launch(const LinClusterParams<T> &p, int k, const LinClusterPlan &pl, cudaStream_t st) {
  PDENET_LIN_CLUSTER_DISPATCH(linear_cluster_bwd_kernel, k, pl.pts, p, p.B * pl.cs, pl.cs, pl.smem, st);
}

template <typename T>
static int linear_fwd(const T *u0, const T *id_kernel, const T *kernels, const T *coe, T *states, int B, int H,
                      int W, int k, int N, int steps, double dt, void *stream) {
  if (!u0 || !id_kernel || !kernels || !coe || !states || bad_geom(B, 1, H, W, k, N) || steps < 0)
    return PDENET_E_BADARG;
  if (!lin_supported(N + 1)) return PDENET_E_UNSUPPORTED;
  const size_t stride = (size_t)B * H * W;
  if (steps > 0) {
    const LinClusterPlan pl = lin_cluster_plan(sizeof(T), B, H, W, k, N, sm_count());
    if (pl.cs) { // small grid: the whole rollout in one launch, one trajectory per cluster
      LinClusterParams<T> q;
      memset(&q, 0, sizeof(q));
      q.u0 = u0;
      q.k0 = id_kernel;
      q.kernels = kernels;
      q.coe = coe;
      q.states = states;
      q.B = B, q.H = H, q.W = W, q.N = N, q.steps = steps, q.cs = pl.cs, q.hb_max = pl.hb_max;
      q.dt = (T)dt;
      ProfScope ps(0, (cudaStream_t)stream);
      return lin_cluster_fwd_launch<T>(q, k, pl, (cudaStream_t)stream);
    }
  }
  SlotLease lease((cudaStream_t)stream);
  if (int rc = upload<T>(lease, id_kernel, kernels, N, k, (cudaStream_t)stream)) return rc;
  StepParams<T> p;
  memset(&p, 0, sizeof(p));
  p.slot = lease.slot;
  p.symnet = coe;
  p.dt = (T)dt;
  p.wrap_y = 1;
  p.u_rows = p.out_rows = H;
  for (int t = 0; t < steps; ++t) {
    p.u = t == 0 ? u0 : states + (size_t)(t - 1) * stride;
    p.out = states + (size_t)t * stride;
    int rc = 0;
    PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_LIN_FWD, N + 1, 1, 16, p, B, H, W, (cudaStream_t)stream)));
    if (rc) return rc;
  }
  return PDENET_OK;
}

struct LinLayout {
  size_t off_g, off_lam0, off_lam1, off_dk, off_dcoe, total;
  int dk_slots, vlb;
};
static LinLayout lin_layout(size_t es, int B, int H, int W, int k, int N, int sms) {
  LinLayout l;
  const int kk = k * k;
  l.vlb = kk <= 16 ? 16 : (kk <= 32 ? 32 : 64);
  l.dk_slots = dk_slot_count(sms, N + 1);
  const size_t img = (size_t)H * W;
  size_t o = 0;
  l.off_g = o;
  o = align_up(o + (size_t)B * (N + 1) * img * es, 256);
  l.off_lam0 = o;
  o = align_up(o + (size_t)B * img * es, 256);
  l.off_lam1 = o;
  o = align_up(o + (size_t)B * img * es, 256);
  l.off_dk = o;
  o = align_up(o + (size_t)l.dk_slots * (N + 1) * l.vlb * sizeof(double), 256);
  l.off_dcoe = o;
  o = align_up(o + (size_t)B * N * img * es, 256);
  // the small-grid cluster path keeps its tap correlations R [B][k*k][H][W] + their batch sum here instead
  const size_t cluster_need = align_up((size_t)(B + 1) * kk * img * es, 256);
  l.total = o > cluster_need ? o : cluster_need;
  return l;
}

// grad_coe[n][pix] = sum_b partial[b][n][pix]  (fixed order)
template <typename T> __global__ void reduce_dcoe_kernel(const T *partial, int B, size_t n, T *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0; // summed in double for either T: the batch can be long
  for (int b = 0; b < B; ++b) s += (double)partial[(size_t)b * n + i];
  out[i] = (T)s;
}

template <typename T>
static int linear_bwd(const T *u0, const T *states, const T *grad_states, int grad_mode, const T *id_kernel,
                      const T *kernels, const T *coe, T *grad_u0, T *grad_id_kernel, T *grad_kernels, T *grad_coe,
                      void *workspace, size_t ws_bytes, int B, int H, int W, int k, int N, int steps, double dt,
                      void *stream) {
  if (!u0 || !grad_states || !id_kernel || !kernels || !coe || !workspace || bad_geom(B, 1, H, W, k, N) ||
      steps < 1 || (steps > 1 && !states) || (grad_mode != PDENET_GRAD_LAST && grad_mode != PDENET_GRAD_ALL))
    return PDENET_E_BADARG;
  if (!lin_supported(N + 1)) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const LinLayout lay = lin_layout(sizeof(T), B, H, W, k, N, sms);
  if (ws_bytes < lay.total) return PDENET_E_WORKSPACE;
  char *ws = (char *)workspace;
  const size_t img = (size_t)H * W;
  const size_t stride = (size_t)B * img;
  {
    const LinClusterPlan pl = lin_cluster_plan(sizeof(T), B, H, W, k, N, sms, 1);
    if (pl.cs) { // small grid: reverse rollout in one launch + two finishing kernels
      const int kk = k * k;
      LinClusterParams<T> q;
      memset(&q, 0, sizeof(q));
      q.u0 = u0;
      q.k0 = id_kernel;
      q.kernels = kernels;
      q.coe = coe;
      q.states = const_cast<T *>(states);
      q.grad_states = grad_states;
      q.grad_all = grad_mode == PDENET_GRAD_ALL;
      q.grad_u0 = grad_u0;
      q.R = (T *)ws;
      q.B = B, q.H = H, q.W = W, q.N = N, q.steps = steps, q.cs = pl.cs, q.hb_max = pl.hb_max;
      q.dt = (T)dt;
      {
        ProfScope ps(1, st);
        if (int rc = lin_cluster_bwd_launch<T>(q, k, pl, st)) return rc;
      }
      if (grad_id_kernel || grad_kernels || grad_coe) {
        T *Rsum = (T *)ws + (size_t)B * kk * img;
        ProfScope ps(4, st);
        lin_cluster_dcoe_kernel<T><<<(unsigned)((img + kDcoePts - 1) / kDcoePts), kDcoeThreads, 0, st>>>(q.R, B, kk, img, kernels, N, (T)dt, Rsum,
                                                                                  grad_coe);
        if (grad_id_kernel || grad_kernels)
          lin_cluster_dk_kernel<T><<<dim3(kk, N + 1), 256, 0, st>>>(Rsum, coe, 0, img, kk, (T)dt, grad_id_kernel,
                                                                    grad_kernels);
      }
      return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
    }
  }
  T *gbuf = (T *)(ws + lay.off_g);
  T *lamb[2] = {(T *)(ws + lay.off_lam0), (T *)(ws + lay.off_lam1)};
  double *dkp = (double *)(ws + lay.off_dk);
  T *dcoe = (T *)(ws + lay.off_dcoe);
  const bool want_dk = grad_id_kernel != nullptr || grad_kernels != nullptr;
  if (cudaMemsetAsync(dkp, 0, (size_t)lay.dk_slots * (N + 1) * lay.vlb * sizeof(double), st) != cudaSuccess ||
      cudaMemsetAsync(dcoe, 0, (size_t)B * N * img * sizeof(T), st) != cudaSuccess)
    return PDENET_E_CUDA;
  SlotLease lease(st);
  if (int rc = upload<T>(lease, id_kernel, kernels, N, k, st)) return rc;
  const T *lam_in = grad_mode == PDENET_GRAD_ALL ? grad_states + (size_t)(steps - 1) * stride : grad_states;
  int cur = 0;
  for (int t = steps - 1; t >= 0; --t) {
    const T *ut = t == 0 ? u0 : states + (size_t)(t - 1) * stride;
    T *lam_out = t == 0 ? grad_u0 : lamb[cur];
    const T *inject = (grad_mode == PDENET_GRAD_ALL && t > 0) ? grad_states + (size_t)(t - 1) * stride : nullptr;
    StepParams<T> p;
    memset(&p, 0, sizeof(p));
    p.slot = lease.slot;
    p.symnet = coe;
    p.dt = (T)dt;
    p.wrap_y = 1;
    p.u = ut;
    p.u_rows = H;
    p.lam = lam_in;
    p.lam_rows = H;
    p.g = gbuf;
    p.g_rows = H;
    p.dcoe_partial = dcoe;
    int rc = 0;
    PDENET_DISPATCH_K(k, rc = (run_step<T, KK_>(MODE_LIN_BWD_A, N + 1, 1, 16, p, B, H, W, st)));
    if (rc) return rc;
    if (lam_out) {
      AdjParams<T> q;
      memset(&q, 0, sizeof(q));
      q.g = gbuf;
      q.g_rows = H;
      q.wrap_y = 1;
      q.lam_in = nullptr; // the D0 filter carries the identity path
      q.inject = inject;
      q.inj_rows = H;
      q.lam_out = lam_out;
      q.lo_rows = H;
      q.slot = lease.slot;
      PDENET_DISPATCH_K(k, rc = (run_adjoint<T, KK_>(q, B, 1, N + 1, H, W, st)));
      if (rc) return rc;
    }
    if (want_dk) {
      FGradParams<T> f;
      memset(&f, 0, sizeof(f));
      f.g = gbuf;
      f.g_rows = H;
      f.u = ut;
      f.u_rows = H;
      f.wrap_y = 1;
      f.dk_partial = dkp;
      PDENET_DISPATCH_K(k, rc = (run_filter_grad<T, KK_>(f, B, 1, N + 1, H, W, st)));
      if (rc) return rc;
    }
    lam_in = lam_out;
    cur ^= 1;
  }
  const int kk = k * k;
  if (grad_id_kernel) {
    ProfScope ps(4, st);
    reduce_dk_kernel<T><<<reduce_blocks(kk), 128, 0, st>>>(dkp, lay.dk_slots, N + 1, kk, lay.vlb, kk, grad_id_kernel);
  }
  if (grad_kernels) {
    // filters 1..N: same partial layout, shifted by one filter
    const int n = N * kk;
    ProfScope ps(4, st);
    reduce_dk_kernel<T><<<reduce_blocks(n), 128, 0, st>>>(dkp + lay.vlb, lay.dk_slots, N + 1, kk, lay.vlb, n,
                                                         grad_kernels);
  }
  if (grad_coe) {
    const size_t n = (size_t)N * img;
    ProfScope ps(4, st);
    reduce_dcoe_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dcoe, B, n, grad_coe);
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- fused MSE loss + seed (row f-3) ------------------------------------------------------------
template <typename T>
static int mse_seed(const T *states, const T *observed, const double *weights_host, T *grad_states, double *loss_out,
                    size_t n, int steps, void *workspace, size_t ws_bytes, void *stream) {
  if (!states || !observed || !grad_states || !loss_out || !workspace || n == 0 || steps <= 0)
    return PDENET_E_BADARG;
  if (steps > 64) return PDENET_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kMseBlocks) blocks = kMseBlocks;
  if (ws_bytes < (size_t)steps * blocks * sizeof(double)) return PDENET_E_WORKSPACE;
  MseWeights w;
  for (int t = 0; t < steps; ++t) w.w[t] = weights_host ? weights_host[t] : 1.0;
  ProfScope ps(4, st);
  mse_seed_kernel<T><<<dim3((unsigned)blocks, (unsigned)steps), 256, 0, st>>>(states, observed, grad_states,
                                                                              (double *)workspace, n, w);
  mse_finish_kernel<<<1, 32, 0, st>>>((const double *)workspace, (int)(steps * blocks), loss_out);
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- 1-D piecewise Lagrange interpolation (row f-2: SingleNonLinear2d's source term) -----------
static int lagrange_blocks(size_t n, int sms) {
  size_t b = (n + kLagrangeThreads * 4 - 1) / (kLagrangeThreads * 4);
  const size_t cap = (size_t)sms * 4;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
template <typename T>
static int lagrange1d_fwd(const T *x, const T *coef, T *y, size_t n, double lo, double hi, int mesh, int degree,
                          void *stream) {
  if (!x || !coef || !y || n == 0 || mesh < 1 || !(hi > lo)) return PDENET_E_BADARG;
  if (degree < 1 || degree > 4) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const T l = (T)lo, ih = (T)((double)mesh / (hi - lo));
  const unsigned grid = (unsigned)((n + 1023) / 1024 < (size_t)sms * 8 ? (n + 1023) / 1024 : (size_t)sms * 8);
  ProfScope ps(4, st);
  switch (degree) {
  case 1: lagrange1d_fwd_kernel<T, 1><<<grid, 256, 0, st>>>(x, coef, y, n, l, ih, mesh); break;
  case 2: lagrange1d_fwd_kernel<T, 2><<<grid, 256, 0, st>>>(x, coef, y, n, l, ih, mesh); break;
  case 3: lagrange1d_fwd_kernel<T, 3><<<grid, 256, 0, st>>>(x, coef, y, n, l, ih, mesh); break;
  default: lagrange1d_fwd_kernel<T, 4><<<grid, 256, 0, st>>>(x, coef, y, n, l, ih, mesh); break;
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}
template <typename T>
static int lagrange1d_bwd(const T *x, const T *gy, const T *coef, T *gx, T *gcoef, size_t n, double lo, double hi,
                          int mesh, int degree, void *workspace, size_t ws_bytes, void *stream) {
  if (!x || !gy || !coef || !gcoef || !workspace || n == 0 || mesh < 1 || !(hi > lo)) return PDENET_E_BADARG;
  if (degree < 1 || degree > 4) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  const int Q = mesh * degree + 1;
  const size_t smem = (size_t)Q * kLagrangeThreads * sizeof(double);
  if (smem > 200 * 1024) return PDENET_E_UNSUPPORTED;
  const int blocks = lagrange_blocks(n, sms);
  if (ws_bytes < (size_t)blocks * Q * sizeof(double)) return PDENET_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const T l = (T)lo, ih = (T)((double)mesh / (hi - lo));
  double *partial = (double *)workspace;
  ProfScope ps(4, st);
#define PDENET_LAG_CASE(D)                                                                                        \
  {                                                                                                              \
    auto kern = lagrange1d_bwd_kernel<T, D>;                                                                     \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)       \
      return PDENET_E_CUDA;                                                                                      \
    kern<<<blocks, kLagrangeThreads, smem, st>>>(x, gy, coef, gx, partial, n, l, ih, mesh);                      \
  }
  switch (degree) {
  case 1: PDENET_LAG_CASE(1) break;
  case 2: PDENET_LAG_CASE(2) break;
  case 3: PDENET_LAG_CASE(3) break;
  default: PDENET_LAG_CASE(4) break;
  }
#undef PDENET_LAG_CASE
  reduce_flat_kernel<T><<<reduce_blocks(Q), 128, 0, st>>>(partial, blocks, Q, gcoef);
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- PDE-Net 1 SingleNonLinear2d (row f-2): the whole rollout fused, small grids ------------------
// u+ = D0*u + dt (sum_n c_n D_n*u + f(u)), constant coefficients c_n, f a piecewise Lagrange polynomial.
// Runs on the cluster-resident path only (the grids this model is used on: 50 x 50); larger grids return
// PDENET_E_UNSUPPORTED and the caller steps through pdenet_linear_rollout_* + pdenet_lagrange1d_* instead.
struct SnlLayout {
  size_t off_R, off_field, off_lam, off_x, off_part, part_bytes, total;
};
static SnlLayout snl_layout(size_t es, int B, int H, int W, int k, int N, int steps, int mesh, int degree, int sms) {
  SnlLayout l;
  const size_t img = (size_t)H * W, kk = (size_t)k * k, n = (size_t)steps * B * img;
  size_t o = 0;
  l.off_R = o;
  o = align_up(o + (B + 1) * kk * img * es, 256);
  l.off_field = o;
  o = align_up(o + (size_t)N * img * es, 256);
  l.off_lam = o;
  o = align_up(o + n * es, 256);
  l.off_x = o;
  o = align_up(o + n * es, 256);
  l.off_part = o;
  l.part_bytes = (size_t)lagrange_blocks(n, sms) * (mesh * degree + 1) * sizeof(double);
  o = align_up(o + l.part_bytes, 256);
  l.total = o;
  return l;
}

template <typename T>
static int snl_fwd(const T *u0, const T *k0, const T *kernels, const T *coe, const T *fcoef, T *states, int B, int H,
                   int W, int k, int N, int steps, double dt, double lo, double hi, int mesh, int degree, void *stream) {
  if (!u0 || !k0 || !kernels || !coe || !fcoef || !states || bad_geom(B, 1, H, W, k, N) || steps < 1 || mesh < 1 ||
      !(hi > lo))
    return PDENET_E_BADARG;
  if (degree < 1 || degree > 4) return PDENET_E_UNSUPPORTED;
  const LinClusterPlan pl = lin_cluster_plan(sizeof(T), B, H, W, k, N, sm_count());
  if (!pl.cs) return PDENET_E_UNSUPPORTED;
  LinClusterParams<T> q;
  memset(&q, 0, sizeof(q));
  q.u0 = u0, q.k0 = k0, q.kernels = kernels, q.coe = coe, q.coe_const = 1, q.states = states;
  q.fcoef = fcoef, q.f_lo = (T)lo, q.f_inv_h = (T)((double)mesh / (hi - lo)), q.f_mesh = mesh, q.f_deg = degree;
  q.B = B, q.H = H, q.W = W, q.N = N, q.steps = steps, q.cs = pl.cs, q.hb_max = pl.hb_max;
  q.dt = (T)dt;
  ProfScope ps(0, (cudaStream_t)stream);
  return lin_cluster_fwd_launch<T>(q, k, pl, (cudaStream_t)stream);
}

template <typename T>
static int snl_bwd(const T *u0, const T *states, const T *grad_states, int grad_mode, const T *k0, const T *kernels,
                   const T *coe, const T *fcoef, T *grad_u0, T *grad_k0, T *grad_kernels, T *grad_coe, T *grad_fcoef,
                   void *workspace, size_t ws_bytes, int B, int H, int W, int k, int N, int steps, double dt, double lo,
                   double hi, int mesh, int degree, void *stream) {
  if (!u0 || !grad_states || !k0 || !kernels || !coe || !fcoef || !workspace || bad_geom(B, 1, H, W, k, N) ||
      steps < 1 || (steps > 1 && !states) || mesh < 1 || !(hi > lo) ||
      (grad_mode != PDENET_GRAD_LAST && grad_mode != PDENET_GRAD_ALL))
    return PDENET_E_BADARG;
  if (degree < 1 || degree > 4) return PDENET_E_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  const LinClusterPlan pl = lin_cluster_plan(sizeof(T), B, H, W, k, N, sms, 1);
  if (!pl.cs) return PDENET_E_UNSUPPORTED;
  const SnlLayout lay = snl_layout(sizeof(T), B, H, W, k, N, steps, mesh, degree, sms);
  if (ws_bytes < lay.total) return PDENET_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = (char *)workspace;
  const size_t img = (size_t)H * W, stride = (size_t)B * img;
  const int kk = k * k;
  T *lam_all = (T *)(ws + lay.off_lam), *x_all = (T *)(ws + lay.off_x), *field = (T *)(ws + lay.off_field);
  LinClusterParams<T> q;
  memset(&q, 0, sizeof(q));
  q.u0 = u0, q.k0 = k0, q.kernels = kernels, q.coe = coe, q.coe_const = 1, q.states = const_cast<T *>(states);
  q.grad_states = grad_states, q.grad_all = grad_mode == PDENET_GRAD_ALL, q.grad_u0 = grad_u0;
  q.R = (T *)(ws + lay.off_R);
  q.fcoef = fcoef, q.f_lo = (T)lo, q.f_inv_h = (T)((double)mesh / (hi - lo)), q.f_mesh = mesh, q.f_deg = degree;
  q.lam_all = grad_fcoef ? lam_all : nullptr;
  q.B = B, q.H = H, q.W = W, q.N = N, q.steps = steps, q.cs = pl.cs, q.hb_max = pl.hb_max;
  q.dt = (T)dt;
  {
    ProfScope ps(1, st);
    if (int rc = lin_cluster_bwd_launch<T>(q, k, pl, st)) return rc;
  }
  if (grad_k0 || grad_kernels || grad_coe) {
    T *Rsum = q.R + (size_t)B * kk * img;
    ProfScope ps(4, st);
    lin_cluster_dcoe_kernel<T><<<(unsigned)((img + kDcoePts - 1) / kDcoePts), kDcoeThreads, 0, st>>>(q.R, B, kk, img, kernels, N, (T)dt, Rsum,
                                                                              grad_coe ? field : nullptr);
    if (grad_coe) lin_cluster_dcoe_const_kernel<T><<<N, 256, 0, st>>>(field, img, grad_coe);
    if (grad_k0 || grad_kernels)
      lin_cluster_dk_kernel<T><<<dim3(kk, N + 1), 256, 0, st>>>(Rsum, coe, 1, img, kk, (T)dt, grad_k0, grad_kernels);
  }
  if (grad_fcoef) { // node gradient: sum over steps and points of (dt lam_{t+1}) * basis(u_t), one deterministic reduction
    if (cudaMemcpyAsync(x_all, u0, stride * sizeof(T), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return PDENET_E_CUDA;
    if (steps > 1 && cudaMemcpyAsync(x_all + stride, states, (size_t)(steps - 1) * stride * sizeof(T),
                                     cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return PDENET_E_CUDA;
    if (int rc = lagrange1d_bwd<T>(x_all, lam_all, fcoef, nullptr, grad_fcoef, (size_t)steps * stride, lo, hi, mesh,
                                   degree, ws + lay.off_part, lay.part_bytes, stream))
      return rc;
  }
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

// ---- FMA peak microbenchmark -----------------------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T a, T b) {
  T x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (T)(threadIdx.x + i) * (T)1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fma_(x[i], a, b);
  }
  T s = T(0);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == (T)123456789) out[0] = s; // never true; keeps the loop alive
}

// operand-pattern probes: 16 accumulators, 16 + 16 multiplicands kept in registers
template <typename T, int VARIANT> __global__ void __launch_bounds__(256) fma_probe_kernel(T *out, int iters) {
  T x[16], y[16], z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = (T)(threadIdx.x + i) * (T)1e-3;
    y[i] = (T)0.999 + (T)i * (T)1e-6;
    z[i] = (T)1e-3 * (T)(i + 1);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (VARIANT == 1) x[i] = fma_(y[(i + r) & 15], z[(i + 3 * r + 1) & 15], x[i]);
        if (VARIANT == 2) x[i] = fma_(y[(i + r) & 15], z[((i >> 2) + r) & 15], x[i]);
      }
    if (it == iters + 1) { // never true: keeps y, z live as variables
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        y[i] += x[i];
        z[i] += x[i];
      }
    }
  }
  T s = T(0);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == (T)123456789) out[0] = s;
}

// operand pattern of the filter-gradient kernel: 49 accumulators per thread, the g value shared by
// 7 consecutive FMAs, the window row re-read from shared memory per row.  VEC = columns per lane
// (2: every accumulator meets two adjacent window registers; 1: exactly one).
template <typename T, int VEC> __global__ void __launch_bounds__(192, 2) fgrad_probe_kernel(T *out, int iters) {
  constexpr int K = 7, RW = 16 / VEC;
  __shared__ T sm[16 * 48];
  for (int i = threadIdx.x; i < 16 * 48; i += blockDim.x) sm[i] = (T)0.999 + (T)i * (T)1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  T acc[K * K], gc[RW][VEC];
#pragma unroll
  for (int e = 0; e < K * K; ++e) acc[e] = T(0);
#pragma unroll
  for (int j = 0; j < RW; ++j)
#pragma unroll
    for (int i = 0; i < VEC; ++i) gc[j][i] = (T)1e-3 * (T)(threadIdx.x + j + 3 * i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rho = 0; rho < RW + K - 1; ++rho) {
      T row[VEC + K - 1];
#pragma unroll
      for (int x = 0; x < VEC + K - 1; ++x) row[x] = sm[((rho + it) & 15) * 48 + ((lane * VEC + x) & 47)];
#pragma unroll
      for (int i = 0; i < VEC; ++i)
#pragma unroll
        for (int a = 0; a < K; ++a) {
          const int j = rho - a;
          if (j >= 0 && j < RW) {
#pragma unroll
            for (int bb = 0; bb < K; ++bb) acc[a * K + bb] = fma_(gc[j][i], row[i + bb], acc[a * K + bb]);
          }
        }
    }
    if (it == iters + 1) { // never true: keeps gc live as variables
#pragma unroll
      for (int j = 0; j < RW; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) gc[j][i] += acc[j];
    }
  }
  T s = T(0);
#pragma unroll
  for (int e = 0; e < K * K; ++e) s += acc[e];
  if (s == (T)123456789) out[0] = s;
}

// the same 49-accumulator pattern without shared-memory reads and without the ragged first / last
// window rows: every pass touches all 49 accumulators (isolates the register-operand limit)
template <typename T, bool LOADS> __global__ void __launch_bounds__(192, 2) fgrad_probe_dense_kernel(T *out, int iters) {
  constexpr int K = 7;
  __shared__ T sm[16 * 48];
  for (int i = threadIdx.x; i < 16 * 48; i += blockDim.x) sm[i] = (T)0.999 + (T)i * (T)1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  T acc[K * K], gc[16], row[K + 1];
#pragma unroll
  for (int e = 0; e < K * K; ++e) acc[e] = T(0);
#pragma unroll
  for (int j = 0; j < 16; ++j) gc[j] = (T)1e-3 * (T)(threadIdx.x + j);
#pragma unroll
  for (int x = 0; x < K + 1; ++x) row[x] = (T)0.999 + (T)x * (T)1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (LOADS) { // a fresh window row per pass, as the real kernel reads it (one value of the 8 is reused)
#pragma unroll
        for (int x = 0; x < K + 1; ++x) row[x] = sm[((r + it) & 15) * 48 + ((lane * 2 + x) & 47)];
      }
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int bb = 0; bb < K; ++bb) acc[a * K + bb] = fma_(gc[(r + a) & 15], row[(r & 1) + bb], acc[a * K + bb]);
    }
    if (it == iters + 1) { // never true: keeps gc, row live as variables
#pragma unroll
      for (int j = 0; j < 16; ++j) gc[j] += acc[j];
#pragma unroll
      for (int x = 0; x < K + 1; ++x) row[x] += acc[x];
    }
  }
  T s = T(0);
#pragma unroll
  for (int e = 0; e < K * K; ++e) s += acc[e];
  if (s == (T)123456789) out[0] = s;
}

} // namespace pdenet

using namespace pdenet;

extern "C" {

int pdenet_version(void) { return 200; }

#ifndef PDENET_BUILD_ID
#define PDENET_BUILD_ID "unknown"
#endif
const char *pdenet_build_id(void) { return PDENET_BUILD_ID; }

const char *pdenet_error_string(int code) {
  switch (code) {
  case PDENET_OK:
    return "ok";
  case PDENET_E_BADARG:
    return "bad argument (null pointer, non-positive size, even kernel size)";
  case PDENET_E_UNSUPPORTED:
    return "unsupported (kernel_size, filters, channels, layers) combination";
  case PDENET_E_WORKSPACE:
    return "workspace too small";
  case PDENET_E_CUDA:
    return "CUDA error";
  default:
    return "unknown error";
  }
}

int pdenet_supported(int elem_size, int k, int N, int C, int L) {
  if (elem_size != 4 && elem_size != 8) return 0;
  if (k != 3 && k != 5 && k != 7) return 0;
  if (L < 0 || L > PDENET_MAX_LAYERS) return 0;
  if (!nc_supported(N, C)) return 0;
  return pick_vl(N * C, L) != 0;
}

int pdenet_symnet_param_count(int m, int L) { return symnet_param_count(m, L); }

int pdenet_rollout_bwd_needs_features(int elem_size, int k, int N, int C, int L) {
  if (!pdenet_supported(elem_size, k, N, C, L)) return -1;
  const int Psym = symnet_param_count(N * C, L);
  size_t smem = 0;
  if (elem_size == 8) {
    PDENET_DISPATCH_K(k, smem = (step_smem_bytes<double, KK_>(MODE_BWD_A, N, C, Psym)));
  } else {
    PDENET_DISPATCH_K(k, smem = (step_smem_bytes<float, KK_>(MODE_BWD_A, N, C, Psym)));
  }
  return smem > (size_t)227 * 1024 ? 1 : 0;
}

int pdenet_m2k_f64(const double *m, const double *s, double *kout, int N, int k, void *st) {
  return sandwich<double>(SW_M2K, m, s, nullptr, kout, N, k, st);
}
int pdenet_m2k_f32(const float *m, const double *s, float *kout, int N, int k, void *st) {
  return sandwich<float>(SW_M2K, m, s, nullptr, kout, N, k, st);
}
int pdenet_k2m_f64(const double *kin, const double *s, double *m, int N, int k, void *st) {
  return sandwich<double>(SW_K2M, kin, s, nullptr, m, N, k, st);
}
int pdenet_k2m_f32(const float *kin, const double *s, float *m, int N, int k, void *st) {
  return sandwich<float>(SW_K2M, kin, s, nullptr, m, N, k, st);
}
int pdenet_m2k_bwd_f64(const double *gk, const double *s, const unsigned char *mask, double *gm, int N, int k,
                       void *st) {
  return sandwich<double>(SW_M2K_BWD, gk, s, mask, gm, N, k, st);
}
int pdenet_m2k_bwd_f32(const float *gk, const double *s, const unsigned char *mask, float *gm, int N, int k,
                       void *st) {
  return sandwich<float>(SW_M2K_BWD, gk, s, mask, gm, N, k, st);
}

int pdenet_fd2d_fwd_f64(const double *u, const double *kernels, double *feats, int B, int C, int H, int W, int k,
                        int N, void *st) {
  return fd2d_fwd<double>(u, kernels, feats, B, C, H, W, k, N, st);
}
int pdenet_fd2d_fwd_f32(const float *u, const float *kernels, float *feats, int B, int C, int H, int W, int k, int N,
                        void *st) {
  return fd2d_fwd<float>(u, kernels, feats, B, C, H, W, k, N, st);
}

size_t pdenet_fd2d_bwd_workspace_bytes(int k, int N) {
  const int sms = sm_count();
  if (sms <= 0 || k <= 0 || N <= 0) return 0;
  const int kk = k * k, vlb = kk <= 16 ? 16 : (kk <= 32 ? 32 : 64);
  return (size_t)dk_slot_count(sms, N) * N * vlb * sizeof(double);
}
int pdenet_fd2d_bwd_f64(const double *u, const double *grad_feats, const double *kernels, double *grad_u,
                        double *grad_kernels, void *workspace, size_t workspace_bytes, int B, int C, int H, int W,
                        int k, int N, void *st) {
  return fd2d_bwd<double>(u, grad_feats, kernels, grad_u, grad_kernels, workspace, workspace_bytes, B, C, H, W, k, N,
                          st);
}
int pdenet_fd2d_bwd_f32(const float *u, const float *grad_feats, const float *kernels, float *grad_u,
                        float *grad_kernels, void *workspace, size_t workspace_bytes, int B, int C, int H, int W,
                        int k, int N, void *st) {
  return fd2d_bwd<float>(u, grad_feats, kernels, grad_u, grad_kernels, workspace, workspace_bytes, B, C, H, W, k, N,
                         st);
}

int pdenet_rollout_fwd_f64(const double *u0, const double *kernels, const double *symnet, double *states, int B,
                           int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_fwd<double>(u0, kernels, symnet, states, nullptr, B, C, H, W, k, N, L, steps, dt, st);
}
int pdenet_rollout_fwd_f32(const float *u0, const float *kernels, const float *symnet, float *states, int B, int C,
                           int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_fwd<float>(u0, kernels, symnet, states, nullptr, B, C, H, W, k, N, L, steps, dt, st);
}
int pdenet_rollout_fwd_keep_f64(const double *u0, const double *kernels, const double *symnet, double *states,
                                double *feats, int B, int C, int H, int W, int k, int N, int L, int steps, double dt,
                                void *st) {
  if (!feats) return PDENET_E_BADARG;
  return rollout_fwd<double>(u0, kernels, symnet, states, feats, B, C, H, W, k, N, L, steps, dt, st);
}
int pdenet_rollout_fwd_keep_f32(const float *u0, const float *kernels, const float *symnet, float *states,
                                float *feats, int B, int C, int H, int W, int k, int N, int L, int steps, double dt,
                                void *st) {
  if (!feats) return PDENET_E_BADARG;
  return rollout_fwd<float>(u0, kernels, symnet, states, feats, B, C, H, W, k, N, L, steps, dt, st);
}

int pdenet_rhs_fwd_f64(const double *u, const double *kernels, const double *symnet, double *out, int B, int C, int H,
                       int W, int k, int N, int L, void *st) {
  return rollout_fwd<double>(u, kernels, symnet, out, nullptr, B, C, H, W, k, N, L, 1, 1.0, st, 1);
}
int pdenet_rhs_fwd_f32(const float *u, const float *kernels, const float *symnet, float *out, int B, int C, int H,
                       int W, int k, int N, int L, void *st) {
  return rollout_fwd<float>(u, kernels, symnet, out, nullptr, B, C, H, W, k, N, L, 1, 1.0, st, 1);
}
int pdenet_rhs_bwd_f64(const double *u, const double *grad_out, const double *kernels, const double *symnet,
                       double *grad_u, double *grad_kernels, double *grad_symnet, void *workspace,
                       size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                       void *st) {
  return rollout_bwd<double>(u, nullptr, nullptr, grad_out, PDENET_GRAD_LAST, kernels, symnet, grad_u, grad_kernels,
                             grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, 1, 1.0, st,
                             false);
}
int pdenet_rhs_bwd_f32(const float *u, const float *grad_out, const float *kernels, const float *symnet,
                       float *grad_u, float *grad_kernels, float *grad_symnet, void *workspace,
                       size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                       void *st) {
  return rollout_bwd<float>(u, nullptr, nullptr, grad_out, PDENET_GRAD_LAST, kernels, symnet, grad_u, grad_kernels,
                            grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, 1, 1.0, st,
                            false);
}

int pdenet_rollout_fwd_upwind_f64(const double *u0, const double *kernels, const double *symnet, double *states, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_fwd<double>(u0, kernels, symnet, states, nullptr, B, C, H, W, k, N, L, steps, dt, st, 0, true);
}
int pdenet_rollout_fwd_upwind_f32(const float *u0, const float *kernels, const float *symnet, float *states, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_fwd<float>(u0, kernels, symnet, states, nullptr, B, C, H, W, k, N, L, steps, dt, st, 0, true);
}
int pdenet_rhs_upwind_fwd_f64(const double *u, const double *kernels, const double *symnet, double *out, int B, int C,
                              int H, int W, int k, int N, int L, void *st) {
  return rollout_fwd<double>(u, kernels, symnet, out, nullptr, B, C, H, W, k, N, L, 1, 1.0, st, 1, true);
}
int pdenet_rhs_upwind_fwd_f32(const float *u, const float *kernels, const float *symnet, float *out, int B, int C,
                              int H, int W, int k, int N, int L, void *st) {
  return rollout_fwd<float>(u, kernels, symnet, out, nullptr, B, C, H, W, k, N, L, 1, 1.0, st, 1, true);
}
size_t pdenet_rollout_bwd_upwind_workspace_bytes(int elem_size, int B, int C, int H, int W, int k, int N, int L,
                                                 int batch_chunk) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, C, H, W, k, N) || L < 0 || !upwind_supported(N, C)) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return bwd_layout_upwind((size_t)elem_size, B, C, H, W, k, N, L, batch_chunk, sms).total;
}
int pdenet_rollout_bwd_upwind_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                                  const double *kernels, const double *symnet, double *grad_u0, double *grad_kernels,
                                  double *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_bwd<double>(u0, states, nullptr, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                             grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st,
                             true, true);
}
int pdenet_rollout_bwd_upwind_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                                  const float *kernels, const float *symnet, float *grad_u0, float *grad_kernels,
                                  float *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_bwd<float>(u0, states, nullptr, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                            grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st,
                            true, true);
}
int pdenet_rhs_upwind_bwd_f64(const double *u, const double *grad_out, const double *kernels, const double *symnet,
                              double *grad_u, double *grad_kernels, double *grad_symnet, void *workspace,
                              size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                              void *st) {
  return rollout_bwd<double>(u, nullptr, nullptr, grad_out, PDENET_GRAD_LAST, kernels, symnet, grad_u, grad_kernels,
                             grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, 1, 1.0, st,
                             false, true);
}
int pdenet_rhs_upwind_bwd_f32(const float *u, const float *grad_out, const float *kernels, const float *symnet,
                              float *grad_u, float *grad_kernels, float *grad_symnet, void *workspace,
                              size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                              void *st) {
  return rollout_bwd<float>(u, nullptr, nullptr, grad_out, PDENET_GRAD_LAST, kernels, symnet, grad_u, grad_kernels,
                            grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, 1, 1.0, st,
                            false, true);
}
int pdenet_upwind_supported(int elem_size, int k, int N, int C, int L) {
  return pdenet_supported(elem_size, k, N, C, L) && upwind_supported(N, C) ? 1 : 0;
}

size_t pdenet_rollout_bwd_workspace_bytes(int elem_size, int B, int C, int H, int W, int k, int N, int L,
                                          int batch_chunk) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, C, H, W, k, N) || L < 0) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return bwd_layout((size_t)elem_size, B, C, H, W, k, N, L, batch_chunk, sms).total;
}
int pdenet_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                           const double *kernels, const double *symnet, double *grad_u0, double *grad_kernels,
                           double *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                           int C, int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_bwd<double>(u0, states, nullptr, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                             grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt,
                             st);
}
int pdenet_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                           const float *kernels, const float *symnet, float *grad_u0, float *grad_kernels,
                           float *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B, int C,
                           int H, int W, int k, int N, int L, int steps, double dt, void *st) {
  return rollout_bwd<float>(u0, states, nullptr, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                            grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st);
}
int pdenet_rollout_bwd_kept_f64(const double *u0, const double *states, const double *feats,
                                const double *grad_states, int grad_mode, const double *kernels,
                                const double *symnet, double *grad_u0, double *grad_kernels, double *grad_symnet,
                                void *workspace, size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W,
                                int k, int N, int L, int steps, double dt, void *st) {
  if (!feats) return PDENET_E_BADARG;
  return rollout_bwd<double>(u0, states, feats, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                             grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt,
                             st);
}
int pdenet_rollout_bwd_kept_f32(const float *u0, const float *states, const float *feats, const float *grad_states,
                                int grad_mode, const float *kernels, const float *symnet, float *grad_u0,
                                float *grad_kernels, float *grad_symnet, void *workspace, size_t workspace_bytes,
                                int batch_chunk, int B, int C, int H, int W, int k, int N, int L, int steps,
                                double dt, void *st) {
  if (!feats) return PDENET_E_BADARG;
  return rollout_bwd<float>(u0, states, feats, grad_states, grad_mode, kernels, symnet, grad_u0, grad_kernels,
                            grad_symnet, workspace, workspace_bytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st);
}

int pdenet_linear_rollout_fwd_f64(const double *u0, const double *id_kernel, const double *kernels,
                                  const double *coe, double *states, int B, int H, int W, int k, int N, int steps,
                                  double dt, void *st) {
  return linear_fwd<double>(u0, id_kernel, kernels, coe, states, B, H, W, k, N, steps, dt, st);
}
int pdenet_linear_rollout_fwd_f32(const float *u0, const float *id_kernel, const float *kernels, const float *coe,
                                  float *states, int B, int H, int W, int k, int N, int steps, double dt, void *st) {
  return linear_fwd<float>(u0, id_kernel, kernels, coe, states, B, H, W, k, N, steps, dt, st);
}
size_t pdenet_linear_rollout_bwd_workspace_bytes(int elem_size, int B, int H, int W, int k, int N) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, 1, H, W, k, N)) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return lin_layout((size_t)elem_size, B, H, W, k, N, sms).total;
}
int pdenet_linear_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                                  const double *id_kernel, const double *kernels, const double *coe,
                                  double *grad_u0, double *grad_id_kernel, double *grad_kernels, double *grad_coe,
                                  void *workspace, size_t workspace_bytes, int B, int H, int W, int k, int N,
                                  int steps, double dt, void *st) {
  return linear_bwd<double>(u0, states, grad_states, grad_mode, id_kernel, kernels, coe, grad_u0, grad_id_kernel,
                            grad_kernels, grad_coe, workspace, workspace_bytes, B, H, W, k, N, steps, dt, st);
}
int pdenet_linear_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                                  const float *id_kernel, const float *kernels, const float *coe, float *grad_u0,
                                  float *grad_id_kernel, float *grad_kernels, float *grad_coe, void *workspace,
                                  size_t workspace_bytes, int B, int H, int W, int k, int N, int steps, double dt,
                                  void *st) {
  return linear_bwd<float>(u0, states, grad_states, grad_mode, id_kernel, kernels, coe, grad_u0, grad_id_kernel,
                           grad_kernels, grad_coe, workspace, workspace_bytes, B, H, W, k, N, steps, dt, st);
}

int pdenet_slab_step_fwd_f64(const double *u_ext, const double *kernels, const double *symnet, double *out, int B,
                             int C, int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k,
                             int N, int L, double dt, void *st) {
  return slab_step_fwd<double>(u_ext, kernels, symnet, out, nullptr, nullptr, nullptr, 0, B, C, Hs, W, halo,
                               out_rows_per_image, out_row_offset, k, N, L, dt, st);
}
int pdenet_slab_step_fwd_f32(const float *u_ext, const float *kernels, const float *symnet, float *out, int B, int C,
                             int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N,
                             int L, double dt, void *st) {
  return slab_step_fwd<float>(u_ext, kernels, symnet, out, nullptr, nullptr, nullptr, 0, B, C, Hs, W, halo,
                              out_rows_per_image, out_row_offset, k, N, L, dt, st);
}

size_t pdenet_slab_bwd_workspace_bytes(int elem_size, int B, int C, int Hs, int W, int k, int N, int L) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, C, Hs, W, k, N) || L < 0) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return slab_layout((size_t)elem_size, B, C, Hs, W, k, N, L, sms).total;
}
int pdenet_slab_bwd_begin(void *workspace, size_t workspace_bytes, int elem_size, int B, int C, int Hs, int W, int k,
                          int N, int L, void *stream) {
  if (!workspace || (elem_size != 4 && elem_size != 8) || bad_geom(B, C, Hs, W, k, N) || L < 0)
    return PDENET_E_BADARG;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  const SlabLayout lay = slab_layout((size_t)elem_size, B, C, Hs, W, k, N, L, sms);
  if (workspace_bytes < lay.total) return PDENET_E_WORKSPACE;
  char *ws = (char *)workspace;
  return cudaMemsetAsync(ws + lay.off_dk, 0, lay.total - lay.off_dk, (cudaStream_t)stream) == cudaSuccess
             ? PDENET_OK
             : PDENET_E_CUDA;
}
int pdenet_slab_step_bwd_f64(const double *u_ext, const double *lam_ext, const double *kernels,
                             const double *symnet, double *lam_out, int out_rows_per_image, int out_row_offset,
                             void *workspace, size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N,
                             int L, double dt, void *st) {
  return slab_step_bwd<double>(u_ext, lam_ext, nullptr, kernels, symnet, lam_out, nullptr, nullptr, 0, out_rows_per_image,
                               out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L, dt, st);
}
int pdenet_slab_step_bwd_f32(const float *u_ext, const float *lam_ext, const float *kernels, const float *symnet,
                             float *lam_out, int out_rows_per_image, int out_row_offset, void *workspace,
                             size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                             void *st) {
  return slab_step_bwd<float>(u_ext, lam_ext, nullptr, kernels, symnet, lam_out, nullptr, nullptr, 0, out_rows_per_image,
                              out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L, dt, st);
}
int pdenet_slab_step_fwd_p2p_f64(const double *u_ext, const double *kernels, const double *symnet, double *out,
                                 double *peer_up_out, double *peer_down_out, int mirror_rows, int B, int C, int Hs,
                                 int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N, int L,
                                 double dt, void *st) {
  return slab_step_fwd<double>(u_ext, kernels, symnet, out, nullptr, peer_up_out, peer_down_out, mirror_rows, B, C, Hs, W, halo,
                               out_rows_per_image, out_row_offset, k, N, L, dt, st);
}
int pdenet_slab_step_fwd_p2p_f32(const float *u_ext, const float *kernels, const float *symnet, float *out,
                                 float *peer_up_out, float *peer_down_out, int mirror_rows, int B, int C, int Hs,
                                 int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N, int L,
                                 double dt, void *st) {
  return slab_step_fwd<float>(u_ext, kernels, symnet, out, nullptr, peer_up_out, peer_down_out, mirror_rows, B, C, Hs, W, halo,
                              out_rows_per_image, out_row_offset, k, N, L, dt, st);
}
int pdenet_slab_step_bwd_p2p_f64(const double *u_ext, const double *lam_ext, const double *kernels,
                                 const double *symnet, double *lam_out, double *peer_up_out, double *peer_down_out,
                                 int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                 size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                 void *st) {
  return slab_step_bwd<double>(u_ext, lam_ext, nullptr, kernels, symnet, lam_out, peer_up_out, peer_down_out, mirror_rows,
                               out_rows_per_image, out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L,
                               dt, st);
}
int pdenet_slab_step_bwd_p2p_f32(const float *u_ext, const float *lam_ext, const float *kernels,
                                 const float *symnet, float *lam_out, float *peer_up_out, float *peer_down_out,
                                 int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                 size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                 void *st) {
  return slab_step_bwd<float>(u_ext, lam_ext, nullptr, kernels, symnet, lam_out, peer_up_out, peer_down_out, mirror_rows,
                              out_rows_per_image, out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L, dt,
                              st);
}
int pdenet_slab_step_fwd_keep_f64(const double *u_ext, const double *kernels, const double *symnet, double *out,
                                  double *feats, double *peer_up_out, double *peer_down_out, int mirror_rows, int B,
                                  int C, int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k,
                                  int N, int L, double dt, void *st) {
  if (!feats) return PDENET_E_BADARG;
  return slab_step_fwd<double>(u_ext, kernels, symnet, out, feats, peer_up_out, peer_down_out, mirror_rows, B, C, Hs, W,
                               halo, out_rows_per_image, out_row_offset, k, N, L, dt, st);
}
int pdenet_slab_step_fwd_keep_f32(const float *u_ext, const float *kernels, const float *symnet, float *out,
                                  float *feats, float *peer_up_out, float *peer_down_out, int mirror_rows, int B, int C,
                                  int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N,
                                  int L, double dt, void *st) {
  if (!feats) return PDENET_E_BADARG;
  return slab_step_fwd<float>(u_ext, kernels, symnet, out, feats, peer_up_out, peer_down_out, mirror_rows, B, C, Hs, W,
                              halo, out_rows_per_image, out_row_offset, k, N, L, dt, st);
}
int pdenet_slab_step_bwd_kept_f64(const double *u_ext, const double *lam_ext, const double *feats,
                                  const double *kernels, const double *symnet, double *lam_out, double *peer_up_out,
                                  double *peer_down_out, int mirror_rows, int out_rows_per_image, int out_row_offset,
                                  void *workspace, size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N,
                                  int L, double dt, void *st) {
  if (!feats) return PDENET_E_BADARG;
  return slab_step_bwd<double>(u_ext, lam_ext, feats, kernels, symnet, lam_out, peer_up_out, peer_down_out, mirror_rows,
                               out_rows_per_image, out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L,
                               dt, st);
}
int pdenet_slab_step_bwd_kept_f32(const float *u_ext, const float *lam_ext, const float *feats, const float *kernels,
                                  const float *symnet, float *lam_out, float *peer_up_out, float *peer_down_out,
                                  int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                  size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                  void *st) {
  if (!feats) return PDENET_E_BADARG;
  return slab_step_bwd<float>(u_ext, lam_ext, feats, kernels, symnet, lam_out, peer_up_out, peer_down_out, mirror_rows,
                              out_rows_per_image, out_row_offset, workspace, workspace_bytes, B, C, Hs, W, k, N, L, dt,
                              st);
}
int pdenet_slab_bwd_end_f64(void *workspace, size_t workspace_bytes, double *grad_kernels, double *grad_symnet, int B,
                            int C, int Hs, int W, int k, int N, int L, void *st) {
  return slab_bwd_end<double>(workspace, workspace_bytes, grad_kernels, grad_symnet, B, C, Hs, W, k, N, L, st);
}
int pdenet_slab_bwd_end_f32(void *workspace, size_t workspace_bytes, float *grad_kernels, float *grad_symnet, int B,
                            int C, int Hs, int W, int k, int N, int L, void *st) {
  return slab_bwd_end<float>(workspace, workspace_bytes, grad_kernels, grad_symnet, B, C, Hs, W, k, N, L, st);
}

int pdenet_mse_seed_f64(const double *states, const double *observed, const double *weights_host,
                        double *grad_states, double *loss_out, size_t n_per_step, int steps, void *workspace,
                        size_t workspace_bytes, void *st) {
  return mse_seed<double>(states, observed, weights_host, grad_states, loss_out, n_per_step, steps, workspace,
                          workspace_bytes, st);
}
int pdenet_mse_seed_f32(const float *states, const float *observed, const double *weights_host, float *grad_states,
                        double *loss_out, size_t n_per_step, int steps, void *workspace, size_t workspace_bytes,
                        void *st) {
  return mse_seed<float>(states, observed, weights_host, grad_states, loss_out, n_per_step, steps, workspace,
                         workspace_bytes, st);
}

size_t pdenet_lagrange1d_bwd_workspace_bytes(size_t n, int mesh, int degree) {
  const int sms = sm_count();
  if (sms <= 0 || mesh < 1 || degree < 1) return 0;
  return (size_t)lagrange_blocks(n, sms) * (mesh * degree + 1) * sizeof(double);
}
int pdenet_lagrange1d_fwd_f64(const double *x, const double *coef, double *y, size_t n, double lo, double hi, int mesh,
                              int degree, void *st) {
  return lagrange1d_fwd<double>(x, coef, y, n, lo, hi, mesh, degree, st);
}
int pdenet_lagrange1d_fwd_f32(const float *x, const float *coef, float *y, size_t n, double lo, double hi, int mesh,
                              int degree, void *st) {
  return lagrange1d_fwd<float>(x, coef, y, n, lo, hi, mesh, degree, st);
}
int pdenet_lagrange1d_bwd_f64(const double *x, const double *gy, const double *coef, double *gx, double *gcoef,
                              size_t n, double lo, double hi, int mesh, int degree, void *workspace,
                              size_t workspace_bytes, void *st) {
  return lagrange1d_bwd<double>(x, gy, coef, gx, gcoef, n, lo, hi, mesh, degree, workspace, workspace_bytes, st);
}
int pdenet_lagrange1d_bwd_f32(const float *x, const float *gy, const float *coef, float *gx, float *gcoef, size_t n,
                              double lo, double hi, int mesh, int degree, void *workspace, size_t workspace_bytes,
                              void *st) {
  return lagrange1d_bwd<float>(x, gy, coef, gx, gcoef, n, lo, hi, mesh, degree, workspace, workspace_bytes, st);
}

int pdenet_single_nonlinear_fits(int elem_size, int B, int H, int W, int k, int N) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, 1, H, W, k, N)) return 0;
  return lin_cluster_plan((size_t)elem_size, B, H, W, k, N, sm_count()).cs ? 1 : 0;
}
size_t pdenet_single_nonlinear_rollout_bwd_workspace_bytes(int elem_size, int B, int H, int W, int k, int N, int steps,
                                                           int mesh, int degree) {
  if ((elem_size != 4 && elem_size != 8) || bad_geom(B, 1, H, W, k, N) || steps < 1 || mesh < 1 || degree < 1) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return snl_layout((size_t)elem_size, B, H, W, k, N, steps, mesh, degree, sms).total;
}
int pdenet_single_nonlinear_rollout_fwd_f64(const double *u0, const double *id_kernel, const double *kernels,
                                            const double *coe, const double *fcoef, double *states, int B, int H, int W,
                                            int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *st) {
  return snl_fwd<double>(u0, id_kernel, kernels, coe, fcoef, states, B, H, W, k, N, steps, dt, lo, hi, mesh, degree, st);
}
int pdenet_single_nonlinear_rollout_fwd_f32(const float *u0, const float *id_kernel, const float *kernels,
                                            const float *coe, const float *fcoef, float *states, int B, int H, int W,
                                            int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *st) {
  return snl_fwd<float>(u0, id_kernel, kernels, coe, fcoef, states, B, H, W, k, N, steps, dt, lo, hi, mesh, degree, st);
}
int pdenet_single_nonlinear_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states,
                                            int grad_mode, const double *id_kernel, const double *kernels,
                                            const double *coe, const double *fcoef, double *grad_u0,
                                            double *grad_id_kernel, double *grad_kernels, double *grad_coe,
                                            double *grad_fcoef, void *workspace, size_t workspace_bytes, int B, int H,
                                            int W, int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *st) {
  return snl_bwd<double>(u0, states, grad_states, grad_mode, id_kernel, kernels, coe, fcoef, grad_u0, grad_id_kernel,
                         grad_kernels, grad_coe, grad_fcoef, workspace, workspace_bytes, B, H, W, k, N, steps, dt, lo,
                         hi, mesh, degree, st);
}
int pdenet_single_nonlinear_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states,
                                            int grad_mode, const float *id_kernel, const float *kernels,
                                            const float *coe, const float *fcoef, float *grad_u0, float *grad_id_kernel,
                                            float *grad_kernels, float *grad_coe, float *grad_fcoef, void *workspace,
                                            size_t workspace_bytes, int B, int H, int W, int k, int N, int steps,
                                            double dt, double lo, double hi, int mesh, int degree, void *st) {
  return snl_bwd<float>(u0, states, grad_states, grad_mode, id_kernel, kernels, coe, fcoef, grad_u0, grad_id_kernel,
                        grad_kernels, grad_coe, grad_fcoef, workspace, workspace_bytes, B, H, W, k, N, steps, dt, lo, hi,
                        mesh, degree, st);
}

int pdenet_small_tile_factor(int factor) {
  if (factor < 0) return g_small_tile_factor.load();
  return g_small_tile_factor.exchange(factor);
}

int pdenet_filter_grad_tile_level(int level) {
  const int prev = pdenet::g_fgrad_level.load(std::memory_order_relaxed);
  if (level >= -1 && level <= 2) pdenet::g_fgrad_level.store(level, std::memory_order_relaxed);
  return prev;
}
int pdenet_linear_small_grid_path(int on) {
  const int prev = g_linear_cluster_on.exchange(on != 0 ? 1 : 0);
  return prev;
}

int pdenet_profile_enable(int on) {
  g_prof.on.store(on != 0);
  return PDENET_OK;
}
int pdenet_profile_read(double *ms_by_kind, long long *launches_by_kind) {
  double ms[PDENET_PROF_KINDS] = {0, 0, 0, 0, 0};
  int rc = PDENET_OK;
  std::lock_guard<std::mutex> lk(g_prof.m);
  for (auto &r : g_prof.recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess)
      rc = PDENET_E_CUDA;
    ms[r.kind] += t;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.recs.clear();
  for (int i = 0; i < PDENET_PROF_KINDS; ++i) {
    if (ms_by_kind) ms_by_kind[i] = ms[i];
    const long long n = g_prof.launches[i].exchange(0);
    if (launches_by_kind) launches_by_kind[i] = n;
  }
  return rc;
}

int pdenet_fma_peak(int elem_size, int iters, double *flops_out, float *ms_out, void *stream) {
  return pdenet_fma_probe(elem_size, iters, 0, flops_out, ms_out, stream);
}

int pdenet_fma_probe(int elem_size, int iters, int variant, double *flops_out, float *ms_out, void *stream) {
  if ((elem_size != 4 && elem_size != 8) || iters <= 0 || !flops_out || !ms_out || variant < 0 || variant > 6)
    return PDENET_E_BADARG;
  const int sms = sm_count();
  if (sms <= 0) return PDENET_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  void *scratch = nullptr;
  if (cudaMalloc(&scratch, 64) != cudaSuccess) return PDENET_E_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 8, block = variant >= 3 ? 192 : 256;
  for (int rep = 0; rep < 2; ++rep) { // first pass warms up
    cudaEventRecord(e0, st);
    if (elem_size == 8) {
      if (variant == 0) fma_peak_kernel<double><<<grid, block, 0, st>>>((double *)scratch, iters, 0.999, 1e-3);
      if (variant == 1) fma_probe_kernel<double, 1><<<grid, block, 0, st>>>((double *)scratch, iters);
      if (variant == 2) fma_probe_kernel<double, 2><<<grid, block, 0, st>>>((double *)scratch, iters);
      if (variant == 3) fgrad_probe_kernel<double, 2><<<grid, block, 0, st>>>((double *)scratch, iters);
      if (variant == 4) fgrad_probe_kernel<double, 1><<<grid, block, 0, st>>>((double *)scratch, iters);
      if (variant == 5) fgrad_probe_dense_kernel<double, false><<<grid, block, 0, st>>>((double *)scratch, iters);
      if (variant == 6) fgrad_probe_dense_kernel<double, true><<<grid, block, 0, st>>>((double *)scratch, iters);
    } else {
      if (variant == 0) fma_peak_kernel<float><<<grid, block, 0, st>>>((float *)scratch, iters, 0.999f, 1e-3f);
      if (variant == 1) fma_probe_kernel<float, 1><<<grid, block, 0, st>>>((float *)scratch, iters);
      if (variant == 2) fma_probe_kernel<float, 2><<<grid, block, 0, st>>>((float *)scratch, iters);
      if (variant == 3) fgrad_probe_kernel<float, 2><<<grid, block, 0, st>>>((float *)scratch, iters);
      if (variant == 4) fgrad_probe_kernel<float, 1><<<grid, block, 0, st>>>((float *)scratch, iters);
      if (variant == 5) fgrad_probe_dense_kernel<float, false><<<grid, block, 0, st>>>((float *)scratch, iters);
      if (variant == 6) fgrad_probe_dense_kernel<float, true><<<grid, block, 0, st>>>((float *)scratch, iters);
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
  }
  cudaEventElapsedTime(ms_out, e0, e1);
  *flops_out = 2.0 * (variant >= 3 ? 784 : 16 * 8) * (double)iters * (double)grid * block; // variant 5: 16*49 = 784 too
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(scratch);
  return cudaGetLastError() == cudaSuccess ? PDENET_OK : PDENET_E_CUDA;
}

} // extern "C"
