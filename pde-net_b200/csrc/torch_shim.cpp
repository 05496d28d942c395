// torch C++ extension over the C-ABI (include/pdenet_b200.h): registers the hot-path entries as
// dispatcher ops `torch.ops.pdenet_b200.*` (TORCH_LIBRARY), CUDA backend only.  A thin shim: every op
// allocates its outputs / workspace as torch tensors, takes the current CUDA stream and forwards raw
// device pointers to the same extern "C" entry the ctypes binding calls -- no kernels, no logic here.
// Compiled with g++ against the torch headers and linked to libpdenet_b200.so (pde-net_b200/build.py).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/pdenet_b200.h"

namespace {

using at::Tensor;

inline void check(int rc, const char *what) {
  // the Python side maps "(code N)" back to pdenet_b200._lib.PdenetError
  TORCH_CHECK(rc == PDENET_OK, what, " failed: ", pdenet_error_string(rc), " (code ", rc, ")");
}
inline void *stream_of(const Tensor &t) { return (void *)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
inline bool f64(const Tensor &t) {
  TORCH_CHECK(t.scalar_type() == at::kDouble || t.scalar_type() == at::kFloat, "pdenet_b200 supports float64 and float32");
  return t.scalar_type() == at::kDouble;
}
inline Tensor req(const Tensor &t, const char *name, const Tensor *like = nullptr) {
  TORCH_CHECK(t.is_cuda(), "pdenet_b200: `", name, "` must be a CUDA tensor (the hot path has no CPU fallback)");
  if (like) TORCH_CHECK(t.scalar_type() == like->scalar_type(), "pdenet_b200: `", name, "` has the wrong dtype");
  return t.contiguous();
}
template <typename T> inline T *ptr(const Tensor &t) { return t.defined() ? t.data_ptr<T>() : nullptr; }
inline Tensor bytes(size_t n, const Tensor &like) {
  return at::empty({(int64_t)(n > 8 ? n : 8)}, like.options().dtype(at::kByte));
}

constexpr int64_t FLAG_KEEP = 1, FLAG_UPWIND = 2, FLAG_LAST_ONLY = 4;

// ---- rows a7-a10: fused rollout --------------------------------------------------------------
std::tuple<Tensor, Tensor> rollout_fwd(const Tensor &u0_, const Tensor &kernels_, const Tensor &symnet_, int64_t L,
                                       double dt, int64_t steps, int64_t flags) {
  const Tensor u0 = req(u0_, "u0"), kernels = req(kernels_, "kernels", &u0), symnet = req(symnet_, "symnet", &u0);
  TORCH_CHECK(u0.dim() == 4 && kernels.dim() == 3, "u0 must be [B,C,H,W], kernels [N,k,k]");
  c10::cuda::CUDAGuard guard(u0.device());
  const int B = u0.size(0), C = u0.size(1), H = u0.size(2), W = u0.size(3), N = kernels.size(0), k = kernels.size(2);
  Tensor states = at::empty({steps, B, C, H, W}, u0.options());
  Tensor feats;
  void *st = stream_of(u0);
  if (flags & FLAG_UPWIND) {
    check(f64(u0) ? pdenet_rollout_fwd_upwind_f64(ptr<double>(u0), ptr<double>(kernels), ptr<double>(symnet),
                                                  ptr<double>(states), B, C, H, W, k, N, L, steps, dt, st)
                  : pdenet_rollout_fwd_upwind_f32(ptr<float>(u0), ptr<float>(kernels), ptr<float>(symnet),
                                                  ptr<float>(states), B, C, H, W, k, N, L, steps, dt, st),
          "pdenet_rollout_fwd_upwind");
  } else if (flags & FLAG_KEEP) {
    feats = at::empty({steps, B, (int64_t)C * N, H, W}, u0.options());
    check(f64(u0) ? pdenet_rollout_fwd_keep_f64(ptr<double>(u0), ptr<double>(kernels), ptr<double>(symnet),
                                                ptr<double>(states), ptr<double>(feats), B, C, H, W, k, N, L, steps,
                                                dt, st)
                  : pdenet_rollout_fwd_keep_f32(ptr<float>(u0), ptr<float>(kernels), ptr<float>(symnet),
                                                ptr<float>(states), ptr<float>(feats), B, C, H, W, k, N, L, steps, dt,
                                                st),
          "pdenet_rollout_fwd_keep");
  } else {
    check(f64(u0) ? pdenet_rollout_fwd_f64(ptr<double>(u0), ptr<double>(kernels), ptr<double>(symnet),
                                           ptr<double>(states), B, C, H, W, k, N, L, steps, dt, st)
                  : pdenet_rollout_fwd_f32(ptr<float>(u0), ptr<float>(kernels), ptr<float>(symnet), ptr<float>(states),
                                           B, C, H, W, k, N, L, steps, dt, st),
          "pdenet_rollout_fwd");
  }
  return {states, feats.defined() ? feats : at::empty({0}, u0.options())};
}

// ---- row a11: reverse rollout ----------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor> rollout_bwd(const Tensor &u0_, const Tensor &states_, const Tensor &feats_,
                                               const Tensor &grad_, const Tensor &kernels_, const Tensor &symnet_,
                                               int64_t L, double dt, int64_t steps, int64_t batch_chunk,
                                               bool need_u, bool need_k, bool need_s, int64_t flags) {
  const Tensor u0 = req(u0_, "u0"), kernels = req(kernels_, "kernels", &u0), symnet = req(symnet_, "symnet", &u0);
  const Tensor grad = req(grad_, "grad_states", &u0), states = req(states_, "states", &u0);
  c10::cuda::CUDAGuard guard(u0.device());
  const int B = u0.size(0), C = u0.size(1), H = u0.size(2), W = u0.size(3), N = kernels.size(0), k = kernels.size(2);
  const bool upwind = flags & FLAG_UPWIND, kept = (flags & FLAG_KEEP) && feats_.numel() > 0;
  const int mode = (flags & FLAG_LAST_ONLY) ? PDENET_GRAD_LAST : PDENET_GRAD_ALL;
  const int es = (int)u0.element_size();
  const size_t nbytes = upwind ? pdenet_rollout_bwd_upwind_workspace_bytes(es, B, C, H, W, k, N, L, batch_chunk)
                               : pdenet_rollout_bwd_workspace_bytes(es, B, C, H, W, k, N, L, batch_chunk);
  Tensor ws = bytes(nbytes, u0);
  Tensor gu = need_u ? at::empty_like(u0) : Tensor(), gk = need_k ? at::empty_like(kernels) : Tensor(),
         gs = need_s ? at::empty_like(symnet) : Tensor();
  void *st = stream_of(u0);
  int rc;
  if (f64(u0)) {
    using T = double;
    if (upwind)
      rc = pdenet_rollout_bwd_upwind_f64(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(kernels), ptr<T>(symnet),
                                         ptr<T>(gu), ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C,
                                         H, W, k, N, L, steps, dt, st);
    else if (kept)
      rc = pdenet_rollout_bwd_kept_f64(ptr<T>(u0), ptr<T>(states), ptr<T>(req(feats_, "feats", &u0)), ptr<T>(grad), mode,
                                       ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu), ptr<T>(gk), ptr<T>(gs),
                                       ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st);
    else
      rc = pdenet_rollout_bwd_f64(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(kernels), ptr<T>(symnet),
                                  ptr<T>(gu), ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k,
                                  N, L, steps, dt, st);
  } else {
    using T = float;
    if (upwind)
      rc = pdenet_rollout_bwd_upwind_f32(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(kernels), ptr<T>(symnet),
                                         ptr<T>(gu), ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C,
                                         H, W, k, N, L, steps, dt, st);
    else if (kept)
      rc = pdenet_rollout_bwd_kept_f32(ptr<T>(u0), ptr<T>(states), ptr<T>(req(feats_, "feats", &u0)), ptr<T>(grad), mode,
                                       ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu), ptr<T>(gk), ptr<T>(gs),
                                       ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N, L, steps, dt, st);
    else
      rc = pdenet_rollout_bwd_f32(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(kernels), ptr<T>(symnet),
                                  ptr<T>(gu), ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k,
                                  N, L, steps, dt, st);
  }
  check(rc, upwind ? "pdenet_rollout_bwd_upwind" : (kept ? "pdenet_rollout_bwd_kept" : "pdenet_rollout_bwd"));
  auto none = [&](const Tensor &t) { return t.defined() ? t : at::empty({0}, u0.options()); };
  return {none(gu), none(gk), none(gs)};
}

// ---- row a8: F(u) ---------------------------------------------------------------------------
Tensor rhs_fwd(const Tensor &u_, const Tensor &kernels_, const Tensor &symnet_, int64_t L, int64_t flags) {
  const Tensor u = req(u_, "u"), kernels = req(kernels_, "kernels", &u), symnet = req(symnet_, "symnet", &u);
  c10::cuda::CUDAGuard guard(u.device());
  const int B = u.size(0), C = u.size(1), H = u.size(2), W = u.size(3), N = kernels.size(0), k = kernels.size(2);
  Tensor out = at::empty_like(u);
  void *st = stream_of(u);
  const bool upwind = flags & FLAG_UPWIND;
  int rc;
  if (f64(u))
    rc = upwind ? pdenet_rhs_upwind_fwd_f64(ptr<double>(u), ptr<double>(kernels), ptr<double>(symnet), ptr<double>(out),
                                            B, C, H, W, k, N, L, st)
                : pdenet_rhs_fwd_f64(ptr<double>(u), ptr<double>(kernels), ptr<double>(symnet), ptr<double>(out), B, C,
                                     H, W, k, N, L, st);
  else
    rc = upwind ? pdenet_rhs_upwind_fwd_f32(ptr<float>(u), ptr<float>(kernels), ptr<float>(symnet), ptr<float>(out), B,
                                            C, H, W, k, N, L, st)
                : pdenet_rhs_fwd_f32(ptr<float>(u), ptr<float>(kernels), ptr<float>(symnet), ptr<float>(out), B, C, H, W,
                                     k, N, L, st);
  check(rc, upwind ? "pdenet_rhs_upwind_fwd" : "pdenet_rhs_fwd");
  return out;
}

std::tuple<Tensor, Tensor, Tensor> rhs_bwd(const Tensor &u_, const Tensor &grad_, const Tensor &kernels_,
                                           const Tensor &symnet_, int64_t L, int64_t batch_chunk, bool need_u,
                                           bool need_k, bool need_s, int64_t flags) {
  const Tensor u = req(u_, "u"), kernels = req(kernels_, "kernels", &u), symnet = req(symnet_, "symnet", &u);
  const Tensor grad = req(grad_, "grad_out", &u);
  c10::cuda::CUDAGuard guard(u.device());
  const int B = u.size(0), C = u.size(1), H = u.size(2), W = u.size(3), N = kernels.size(0), k = kernels.size(2);
  const bool upwind = flags & FLAG_UPWIND;
  const int es = (int)u.element_size();
  const size_t nbytes = upwind ? pdenet_rollout_bwd_upwind_workspace_bytes(es, B, C, H, W, k, N, L, batch_chunk)
                               : pdenet_rollout_bwd_workspace_bytes(es, B, C, H, W, k, N, L, batch_chunk);
  Tensor ws = bytes(nbytes, u);
  Tensor gu = need_u ? at::empty_like(u) : Tensor(), gk = need_k ? at::empty_like(kernels) : Tensor(),
         gs = need_s ? at::empty_like(symnet) : Tensor();
  void *st = stream_of(u);
  int rc;
  if (f64(u)) {
    using T = double;
    rc = upwind ? pdenet_rhs_upwind_bwd_f64(ptr<T>(u), ptr<T>(grad), ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu),
                                            ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N,
                                            L, st)
                : pdenet_rhs_bwd_f64(ptr<T>(u), ptr<T>(grad), ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu), ptr<T>(gk),
                                     ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N, L, st);
  } else {
    using T = float;
    rc = upwind ? pdenet_rhs_upwind_bwd_f32(ptr<T>(u), ptr<T>(grad), ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu),
                                            ptr<T>(gk), ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N,
                                            L, st)
                : pdenet_rhs_bwd_f32(ptr<T>(u), ptr<T>(grad), ptr<T>(kernels), ptr<T>(symnet), ptr<T>(gu), ptr<T>(gk),
                                     ptr<T>(gs), ws.data_ptr(), nbytes, batch_chunk, B, C, H, W, k, N, L, st);
  }
  check(rc, upwind ? "pdenet_rhs_upwind_bwd" : "pdenet_rhs_bwd");
  auto none = [&](const Tensor &t) { return t.defined() ? t : at::empty({0}, u.options()); };
  return {none(gu), none(gk), none(gs)};
}

// ---- rows a1-a4: moment -> kernel map and its reverse ------------------------------------------
Tensor m2k(const Tensor &moments_, c10::ArrayRef<double> scale, const c10::optional<Tensor> &mask) {
  const Tensor m = req(moments_, "moments");
  TORCH_CHECK(m.dim() == 3, "moments must be [N,k,k]");
  c10::cuda::CUDAGuard guard(m.device());
  const int N = m.size(0), k = m.size(2);
  TORCH_CHECK(scale.empty() || (int)scale.size() == N, "scale must have one entry per filter");
  Tensor out = at::empty_like(m);
  const double *sc = scale.empty() ? nullptr : scale.data();
  check(f64(m) ? pdenet_m2k_f64(ptr<double>(m), sc, ptr<double>(out), N, k, stream_of(m))
               : pdenet_m2k_f32(ptr<float>(m), sc, ptr<float>(out), N, k, stream_of(m)),
        "pdenet_m2k");
  (void)mask;
  return out;
}
Tensor m2k_bwd(const Tensor &gk_, c10::ArrayRef<double> scale, const c10::optional<Tensor> &mask) {
  const Tensor gk = req(gk_, "grad_kernels");
  c10::cuda::CUDAGuard guard(gk.device());
  const int N = gk.size(0), k = gk.size(2);
  Tensor gm = at::empty_like(gk);
  const double *sc = scale.empty() ? nullptr : scale.data();
  const unsigned char *mk = nullptr;
  Tensor mt;
  if (mask.has_value() && mask->defined()) {
    mt = req(*mask, "pinned_mask").to(at::kByte).contiguous();
    mk = mt.data_ptr<unsigned char>();
  }
  check(f64(gk) ? pdenet_m2k_bwd_f64(ptr<double>(gk), sc, mk, ptr<double>(gm), N, k, stream_of(gk))
                : pdenet_m2k_bwd_f32(ptr<float>(gk), sc, mk, ptr<float>(gm), N, k, stream_of(gk)),
        "pdenet_m2k_bwd");
  return gm;
}

// ---- row a12: PDE-Net 1 linear rollout --------------------------------------------------------
Tensor linear_rollout_fwd(const Tensor &u0_, const Tensor &k0_, const Tensor &kernels_, const Tensor &coe_, double dt,
                          int64_t steps) {
  const Tensor u0 = req(u0_, "u0"), k0 = req(k0_, "id_kernel", &u0), kernels = req(kernels_, "kernels", &u0),
               coe = req(coe_, "coe", &u0);
  c10::cuda::CUDAGuard guard(u0.device());
  const int B = u0.size(0), H = u0.size(1), W = u0.size(2), N = kernels.size(0), k = kernels.size(2);
  Tensor states = at::empty({steps, B, H, W}, u0.options());
  check(f64(u0) ? pdenet_linear_rollout_fwd_f64(ptr<double>(u0), ptr<double>(k0), ptr<double>(kernels), ptr<double>(coe),
                                                ptr<double>(states), B, H, W, k, N, steps, dt, stream_of(u0))
                : pdenet_linear_rollout_fwd_f32(ptr<float>(u0), ptr<float>(k0), ptr<float>(kernels), ptr<float>(coe),
                                                ptr<float>(states), B, H, W, k, N, steps, dt, stream_of(u0)),
        "pdenet_linear_rollout_fwd");
  return states;
}
std::tuple<Tensor, Tensor, Tensor, Tensor> linear_rollout_bwd(const Tensor &u0_, const Tensor &states_, const Tensor &grad_,
                                                              const Tensor &k0_, const Tensor &kernels_,
                                                              const Tensor &coe_, double dt, int64_t steps,
                                                              bool last_only, bool need_u, bool need_k0, bool need_k,
                                                              bool need_c) {
  const Tensor u0 = req(u0_, "u0"), k0 = req(k0_, "id_kernel", &u0), kernels = req(kernels_, "kernels", &u0),
               coe = req(coe_, "coe", &u0), grad = req(grad_, "grad_states", &u0), states = req(states_, "states", &u0);
  c10::cuda::CUDAGuard guard(u0.device());
  const int B = u0.size(0), H = u0.size(1), W = u0.size(2), N = kernels.size(0), k = kernels.size(2);
  const size_t nbytes = pdenet_linear_rollout_bwd_workspace_bytes((int)u0.element_size(), B, H, W, k, N);
  Tensor ws = bytes(nbytes, u0);
  Tensor gu = need_u ? at::empty_like(u0) : Tensor(), g0 = need_k0 ? at::empty_like(k0) : Tensor(),
         gk = need_k ? at::empty_like(kernels) : Tensor(), gc = need_c ? at::empty_like(coe) : Tensor();
  const int mode = last_only ? PDENET_GRAD_LAST : PDENET_GRAD_ALL;
  int rc;
  if (f64(u0)) {
    using T = double;
    rc = pdenet_linear_rollout_bwd_f64(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(k0), ptr<T>(kernels),
                                       ptr<T>(coe), ptr<T>(gu), ptr<T>(g0), ptr<T>(gk), ptr<T>(gc), ws.data_ptr(), nbytes,
                                       B, H, W, k, N, steps, dt, stream_of(u0));
  } else {
    using T = float;
    rc = pdenet_linear_rollout_bwd_f32(ptr<T>(u0), ptr<T>(states), ptr<T>(grad), mode, ptr<T>(k0), ptr<T>(kernels),
                                       ptr<T>(coe), ptr<T>(gu), ptr<T>(g0), ptr<T>(gk), ptr<T>(gc), ws.data_ptr(), nbytes,
                                       B, H, W, k, N, steps, dt, stream_of(u0));
  }
  check(rc, "pdenet_linear_rollout_bwd");
  auto none = [&](const Tensor &t) { return t.defined() ? t : at::empty({0}, u0.options()); };
  return {none(gu), none(g0), none(gk), none(gc)};
}

} // namespace

TORCH_LIBRARY(pdenet_b200, m) {
  m.def("rollout_fwd(Tensor u0, Tensor kernels, Tensor symnet, int L, float dt, int steps, int flags) -> (Tensor, Tensor)");
  m.def("rollout_bwd(Tensor u0, Tensor states, Tensor feats, Tensor grad, Tensor kernels, Tensor symnet, int L, float dt, "
        "int steps, int batch_chunk, bool need_u, bool need_k, bool need_s, int flags) -> (Tensor, Tensor, Tensor)");
  m.def("rhs_fwd(Tensor u, Tensor kernels, Tensor symnet, int L, int flags) -> Tensor");
  m.def("rhs_bwd(Tensor u, Tensor grad, Tensor kernels, Tensor symnet, int L, int batch_chunk, bool need_u, bool need_k, "
        "bool need_s, int flags) -> (Tensor, Tensor, Tensor)");
  m.def("m2k(Tensor moments, float[] scale, Tensor? mask) -> Tensor");
  m.def("m2k_bwd(Tensor grad_kernels, float[] scale, Tensor? mask) -> Tensor");
  m.def("linear_rollout_fwd(Tensor u0, Tensor id_kernel, Tensor kernels, Tensor coe, float dt, int steps) -> Tensor");
  m.def("linear_rollout_bwd(Tensor u0, Tensor states, Tensor grad, Tensor id_kernel, Tensor kernels, Tensor coe, float dt, "
        "int steps, bool last_only, bool need_u, bool need_k0, bool need_k, bool need_c) -> (Tensor, Tensor, Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(pdenet_b200, CUDA, m) {
  m.impl("rollout_fwd", &rollout_fwd);
  m.impl("rollout_bwd", &rollout_bwd);
  m.impl("rhs_fwd", &rhs_fwd);
  m.impl("rhs_bwd", &rhs_bwd);
  m.impl("m2k", &m2k);
  m.impl("m2k_bwd", &m2k_bwd);
  m.impl("linear_rollout_fwd", &linear_rollout_fwd);
  m.impl("linear_rollout_bwd", &linear_rollout_bwd);
}
