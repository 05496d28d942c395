/*
 * pdenet_b200.h -- C-ABI of the B200-native PDE-Net / PDE-Net 2.0 time-stepping hot path.
 *
 * Reference interfaces replaced.  The mounted reference (/root/reference) is the upstream
 * `master` branch and contains no source (README.md:8 points at the absent `PDE-Net` /
 * `PDE-Net-2.0` branches; the numerical library aTEAM is un-vendored, .gitignore:1), so no
 * file:line anchor exists for the model code.  Each entry below therefore names the SURVEY.md
 * section-8(a) row (and the upstream symbol recalled there) it stands in for; the only in-mount
 * anchors are README.md:5-8 and figures/pdenet.jpg (the delta-t block chain).
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer to contiguous row-major memory unless the
 *     parameter name ends in `_host`;
 *   - state u: [B, C, H, W]; axis 0 of a filter acts along H, axis 1 along W;
 *   - filter bank: [N, k, k] cross-correlation kernels (PyTorch conv2d semantics, no flip),
 *     already multiplied by 1/dx^|order| ("scaled kernels");
 *   - feature index fed to SymNet: f = c*N + n;
 *   - SymNet flat layout per output channel (Psym scalars): for i < L: W_i[2, m+i] row-major,
 *     b_i[2]; then w_f[m+L], b_f[1];  m = C*N;
 *   - the callee allocates nothing (outputs and workspaces are the caller's); all work is ordered
 *     on `stream` (a cudaStream_t passed as void*); reductions are deterministic (fixed order);
 *   - the only cross-call state is a ring of 8 constant-memory slots per device that hold a
 *     call's filter taps: a call LEASES a free slot for the whole time it is enqueuing work (a
 *     ninth concurrent caller blocks until one is returned), makes `stream` wait (event) for the
 *     slot's previous user and records the slot's release event after its last kernel, so calls
 *     may be issued concurrently from several host threads / streams; the launch counters of the
 *     measurement hooks are atomic.  Inside a CUDA-graph capture the events are not used:
 *     a captured graph is self-consistent, but do not replay it concurrently with other calls
 *     into this library on other streams of the same device;
 *   - return value 0 = success, negative = PDENET_E_* below.  No exceptions cross the boundary.
 */
#ifndef PDENET_B200_H_
#define PDENET_B200_H_

#include <stddef.h>

#if defined(__GNUC__)
#define PDENET_API __attribute__((visibility("default")))
#else
#define PDENET_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PDENET_OK 0
#define PDENET_E_BADARG (-1)      /* null pointer / non-positive size / even k            */
#define PDENET_E_UNSUPPORTED (-2) /* (k, N, C, L) outside the compiled set                 */
#define PDENET_E_WORKSPACE (-3)   /* workspace too small                                   */
#define PDENET_E_CUDA (-4)        /* a CUDA runtime call / launch failed (see last error)  */

#define PDENET_MAX_K 7
#define PDENET_MAX_FILTERS 16
#define PDENET_MAX_LAYERS 8

/* gradient seed layout for the reverse rollout */
#define PDENET_GRAD_LAST 0 /* grad_states is [B,C,H,W]: dLoss/d(state after the last step) */
#define PDENET_GRAD_ALL 1  /* grad_states is [steps,B,C,H,W]: one seed per step            */

/* C-ABI revision (bumped with every signature change; the bindings refuse a mismatching binary) and
 * the content hash of the sources the binary was built from (pde-net_b200/build.py). */
PDENET_API int pdenet_version(void);
PDENET_API const char *pdenet_build_id(void);
PDENET_API const char *pdenet_error_string(int code);
/* 1 if rollout kernels were compiled for (elem_size 4|8, k, N, C); L only needs L<=MAX_LAYERS */
PDENET_API int pdenet_supported(int elem_size, int k, int N, int C, int L);
/* number of SymNet scalars per output channel: sum_{i<L}(2(m+i)+2) + (m+L) + 1 */
PDENET_API int pdenet_symnet_param_count(int m, int L);
/* 1 if the recomputing reverse rollout (pdenet_rollout_bwd_*) does not fit this bank in shared
 * memory (3-channel banks with deep SymNets in fp64) and returns PDENET_E_UNSUPPORTED, so that the
 * feature-keeping pair (pdenet_rollout_fwd_keep_* / pdenet_rollout_bwd_kept_*) must be used;
 * 0 if both work; -1 if the combination is not supported at all. */
PDENET_API int pdenet_rollout_bwd_needs_features(int elem_size, int k, int N, int C, int L);

/* ---- rows a1-a3: moment <-> kernel map (upstream aTEAM MK.py::_MK/M2K/K2M) -----------------
 * kernels[n] = scale[n] * A^-1 . moments[n] . A^-T,  A[i,j] = (j-(k-1)/2)^i / i!.
 * scale_host: N host scalars (NULL = all ones).  k2m is the exact inverse (divides by scale).
 * m2k_bwd maps dLoss/dkernels to dLoss/dmoments: scale * A^-T . dK . A^-1.
 * pinned_mask (device, N*k*k bytes, may be NULL): entries with mask!=0 get gradient 0
 * (row a4 `grad_proj`). */
PDENET_API int pdenet_m2k_f64(const double *moments, const double *scale_host, double *kernels, int N, int k, void *stream);
PDENET_API int pdenet_m2k_f32(const float *moments, const double *scale_host, float *kernels, int N, int k, void *stream);
PDENET_API int pdenet_k2m_f64(const double *kernels, const double *scale_host, double *moments, int N, int k, void *stream);
PDENET_API int pdenet_k2m_f32(const float *kernels, const double *scale_host, float *moments, int N, int k, void *stream);
PDENET_API int pdenet_m2k_bwd_f64(const double *grad_kernels, const double *scale_host, const unsigned char *pinned_mask,
                       double *grad_moments, int N, int k, void *stream);
PDENET_API int pdenet_m2k_bwd_f32(const float *grad_kernels, const double *scale_host, const unsigned char *pinned_mask,
                       float *grad_moments, int N, int k, void *stream);

/* ---- rows a5-a6: periodic FD2d bank (upstream aTEAM FD.py::FD2d.forward + periodicpad) -----
 * feats[b, c*N+n, y, x] = sum_{a,b'} kernels[n,a,b'] * u[b,c,(y+a-r)%H,(x+b'-r)%W]. */
PDENET_API int pdenet_fd2d_fwd_f64(const double *u, const double *kernels, double *feats, int B, int C, int H, int W, int k,
                        int N, void *stream);
PDENET_API int pdenet_fd2d_fwd_f32(const float *u, const float *kernels, float *feats, int B, int C, int H, int W, int k,
                        int N, void *stream);

/* Reverse of the bank: grad_u[b,c] = sum_n corr^T(kernels[n], grad_feats[b,c*N+n]) (may be NULL),
 * grad_kernels[n,a,b'] = sum_{b,c,y,x} grad_feats[b,c*N+n,y,x] * u[b,c,y+a-r,x+b'-r] (may be NULL).
 * Workspace: pdenet_fd2d_bwd_workspace_bytes(k, N) bytes. */
PDENET_API size_t pdenet_fd2d_bwd_workspace_bytes(int k, int N);
PDENET_API int pdenet_fd2d_bwd_f64(const double *u, const double *grad_feats, const double *kernels, double *grad_u,
                        double *grad_kernels, void *workspace, size_t workspace_bytes, int B, int C, int H, int W,
                        int k, int N, void *stream);
PDENET_API int pdenet_fd2d_bwd_f32(const float *u, const float *grad_feats, const float *kernels, float *grad_u,
                        float *grad_kernels, void *workspace, size_t workspace_bytes, int B, int C, int H, int W,
                        int k, int N, void *stream);

/* ---- rows a7-a10: fused rollout (upstream polypde.py::POLYPDE2D.multistep) -----------------
 * states[t] = state after step t+1, t = 0..steps-1;  u_{t+1} = u_t + dt * SymNet(FD2d(u_t)).
 * One fused kernel per step: stencils + SymNet + Euler, tile+halo staged in shared memory. */
PDENET_API int pdenet_rollout_fwd_f64(const double *u0, const double *kernels, const double *symnet, double *states, int B,
                           int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
PDENET_API int pdenet_rollout_fwd_f32(const float *u0, const float *kernels, const float *symnet, float *states, int B, int C,
                           int H, int W, int k, int N, int L, int steps, double dt, void *stream);

/* ---- row a8: the right-hand side itself (upstream polypde.py::POLYPDE2D.RightHandItems) -----
 * out[b,c] = SymNet_c(FD2d(u)[b]) = F(u), no Euler update (the same fused kernel, one launch).
 * pdenet_rhs_bwd_*: given dLoss/dF returns dLoss/du (may be NULL), dLoss/dkernels (may be NULL) and
 * dLoss/dsymnet (may be NULL); workspace as for pdenet_rollout_bwd_* (same size function). */
PDENET_API int pdenet_rhs_fwd_f64(const double *u, const double *kernels, const double *symnet, double *out, int B, int C,
                       int H, int W, int k, int N, int L, void *stream);
PDENET_API int pdenet_rhs_fwd_f32(const float *u, const float *kernels, const float *symnet, float *out, int B, int C, int H,
                       int W, int k, int N, int L, void *stream);
PDENET_API int pdenet_rhs_bwd_f64(const double *u, const double *grad_out, const double *kernels, const double *symnet,
                       double *grad_u, double *grad_kernels, double *grad_symnet, void *workspace,
                       size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                       void *stream);
PDENET_API int pdenet_rhs_bwd_f32(const float *u, const float *grad_out, const float *kernels, const float *symnet,
                       float *grad_u, float *grad_kernels, float *grad_symnet, void *workspace, size_t workspace_bytes,
                       int batch_chunk, int B, int C, int H, int W, int k, int N, int L, void *stream);

/* ---- row a11: reverse rollout -------------------------------------------------------------
 * Recomputes stencils + SymNet from the stored states; returns dLoss/du0 [B,C,H,W] (may be
 * NULL), dLoss/dkernels [N,k,k] (w.r.t. the SCALED kernels; may be NULL = frozen filters) and
 * dLoss/dsymnet [C,Psym] (may be NULL).  Outputs are overwritten, not accumulated.
 * Workspace: at least pdenet_rollout_bwd_workspace_bytes(...) bytes of device memory
 * (`batch_chunk` = how many trajectories share the derivative-adjoint scratch; 0 = all B). */
PDENET_API size_t pdenet_rollout_bwd_workspace_bytes(int elem_size, int B, int C, int H, int W, int k, int N, int L,
                                          int batch_chunk);
PDENET_API int pdenet_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                           const double *kernels, const double *symnet, double *grad_u0, double *grad_kernels,
                           double *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                           int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
PDENET_API int pdenet_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                           const float *kernels, const float *symnet, float *grad_u0, float *grad_kernels,
                           float *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                           int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);

/* ---- row f-1: pseudo-upwind scheme (upstream polypde.py::POLYPDE2D(scheme='upwind'); restated from
 * the description in arXiv 1812.04426 -- the upstream source is not in the mount, so UNPINNED) ------
 * Same arguments and layouts as the central entries above (`kernels` is still the N-filter bank, in bank
 * order: index 1 = order (0,1) along W, index 2 = order (1,0) along H).  Per output channel co and grid
 * point, each first-order input (c, dir) of SymNet_co is taken from the filter as given when the local
 * velocity a = d SymNet_co / d feature (evaluated at the central features) is >= 0, and from its flipped
 * twin  K-[a,b] = -K[a,k-1-b]  (dir W)  /  -K[k-1-a,b]  (dir H)  when a < 0; the selection is treated
 * as constant by the reverse pass.  Twins are evaluated from the already staged tile; the reverse
 * pass returns the gradient folded back onto the N filters.  Supported: N in {3, 6}, C in {1, 2}. */
PDENET_API int pdenet_upwind_supported(int elem_size, int k, int N, int C, int L);
PDENET_API int pdenet_rollout_fwd_upwind_f64(const double *u0, const double *kernels, const double *symnet, double *states,
                                  int B, int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
PDENET_API int pdenet_rollout_fwd_upwind_f32(const float *u0, const float *kernels, const float *symnet, float *states, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
PDENET_API size_t pdenet_rollout_bwd_upwind_workspace_bytes(int elem_size, int B, int C, int H, int W, int k, int N, int L,
                                                 int batch_chunk);
PDENET_API int pdenet_rollout_bwd_upwind_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                                  const double *kernels, const double *symnet, double *grad_u0, double *grad_kernels,
                                  double *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
PDENET_API int pdenet_rollout_bwd_upwind_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                                  const float *kernels, const float *symnet, float *grad_u0, float *grad_kernels,
                                  float *grad_symnet, void *workspace, size_t workspace_bytes, int batch_chunk, int B,
                                  int C, int H, int W, int k, int N, int L, int steps, double dt, void *stream);
/* F(u) itself under the upwind scheme (row a8 with scheme='upwind'); workspace as pdenet_rollout_bwd_upwind_*. */
PDENET_API int pdenet_rhs_upwind_fwd_f64(const double *u, const double *kernels, const double *symnet, double *out, int B,
                              int C, int H, int W, int k, int N, int L, void *stream);
PDENET_API int pdenet_rhs_upwind_fwd_f32(const float *u, const float *kernels, const float *symnet, float *out, int B, int C,
                              int H, int W, int k, int N, int L, void *stream);
PDENET_API int pdenet_rhs_upwind_bwd_f64(const double *u, const double *grad_out, const double *kernels, const double *symnet,
                              double *grad_u, double *grad_kernels, double *grad_symnet, void *workspace,
                              size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                              void *stream);
PDENET_API int pdenet_rhs_upwind_bwd_f32(const float *u, const float *grad_out, const float *kernels, const float *symnet,
                              float *grad_u, float *grad_kernels, float *grad_symnet, void *workspace,
                              size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W, int k, int N, int L,
                              void *stream);

/* Feature-keeping variant of the rollout (same states, input and filter gradients, bit for bit; the SymNet gradient
 * equal up to the order of its sum over the points, ~1e-15 relative in fp64; trades HBM for FP64 work).
 * pdenet_rollout_fwd_keep_* additionally stores the C*N derivative features of every step's
 * INPUT state: feats[t] = FD bank applied to u_t, t = 0..steps-1, layout [steps][B][C*N][H][W]
 * (what the reference's autograd keeps alive as conv2d outputs, polypde.py::UInputs [RECALL]).
 * pdenet_rollout_bwd_kept_* reads them, so its first reverse kernel runs only the SymNet
 * reverse sweep instead of recomputing the stencils.  Everything else as the plain pair above. */
PDENET_API int pdenet_rollout_fwd_keep_f64(const double *u0, const double *kernels, const double *symnet, double *states,
                                double *feats, int B, int C, int H, int W, int k, int N, int L, int steps, double dt,
                                void *stream);
PDENET_API int pdenet_rollout_fwd_keep_f32(const float *u0, const float *kernels, const float *symnet, float *states,
                                float *feats, int B, int C, int H, int W, int k, int N, int L, int steps, double dt,
                                void *stream);
PDENET_API int pdenet_rollout_bwd_kept_f64(const double *u0, const double *states, const double *feats,
                                const double *grad_states, int grad_mode, const double *kernels,
                                const double *symnet, double *grad_u0, double *grad_kernels, double *grad_symnet,
                                void *workspace, size_t workspace_bytes, int batch_chunk, int B, int C, int H, int W,
                                int k, int N, int L, int steps, double dt, void *stream);
PDENET_API int pdenet_rollout_bwd_kept_f32(const float *u0, const float *states, const float *feats, const float *grad_states,
                                int grad_mode, const float *kernels, const float *symnet, float *grad_u0,
                                float *grad_kernels, float *grad_symnet, void *workspace, size_t workspace_bytes,
                                int batch_chunk, int B, int C, int H, int W, int k, int N, int L, int steps,
                                double dt, void *stream);

/* ---- row a12: PDE-Net 1 linear variable-coefficient rollout (VariantCoeLinear2d.forward) ----
 * u_{t+1} = K0 (*) u_t + dt * sum_n coe[n,y,x] * (kernels[n] (*) u_t);  u: [B,H,W],
 * coe: [N,H,W], id_kernel K0: [k,k].  grad_coe is [N,H,W]. */
PDENET_API int pdenet_linear_rollout_fwd_f64(const double *u0, const double *id_kernel, const double *kernels,
                                  const double *coe, double *states, int B, int H, int W, int k, int N, int steps,
                                  double dt, void *stream);
PDENET_API int pdenet_linear_rollout_fwd_f32(const float *u0, const float *id_kernel, const float *kernels, const float *coe,
                                  float *states, int B, int H, int W, int k, int N, int steps, double dt,
                                  void *stream);
/* Small grids (a band of the grid per CTA fits in shared memory; BASELINE config 2: 50 x 50) take a dedicated
 * path: ONE launch per direction, one trajectory per thread-block cluster (1-8 CTAs), state / adjoint resident
 * in shared memory over all steps, halo rows stored straight into the neighbour CTAs' shared memory, one cluster
 * barrier per step, cluster size chosen by what the device can hold at once (cudaOccupancyMaxActiveClusters); the N+1 filters and coefficient fields are folded into one variable-coefficient stencil per point
 * (exact, they are constant over a rollout).  Same results up to re-association.
 * pdenet_linear_small_grid_path(0) switches it off (process-global; returns the previous setting): testing /
 * measurement aid, the tiled kernels then serve every size. */
PDENET_API int pdenet_linear_small_grid_path(int on);
PDENET_API size_t pdenet_linear_rollout_bwd_workspace_bytes(int elem_size, int B, int H, int W, int k, int N);
PDENET_API int pdenet_linear_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states, int grad_mode,
                                  const double *id_kernel, const double *kernels, const double *coe,
                                  double *grad_u0, double *grad_id_kernel, double *grad_kernels, double *grad_coe,
                                  void *workspace, size_t workspace_bytes, int B, int H, int W, int k, int N,
                                  int steps, double dt, void *stream);
PDENET_API int pdenet_linear_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states, int grad_mode,
                                  const float *id_kernel, const float *kernels, const float *coe, float *grad_u0,
                                  float *grad_id_kernel, float *grad_kernels, float *grad_coe, void *workspace,
                                  size_t workspace_bytes, int B, int H, int W, int k, int N, int steps, double dt,
                                  void *stream);

/* ---- slab domain decomposition helpers (config 5) --------------------------------
 * Same step as the rollout, but on a slab of rows that is NOT periodic along H: the caller
 * supplies the slab with `halo` ghost rows above and below already filled (by the NVLink halo
 * exchange): u_ext is [B, C, Hs + 2*halo, W]; out is [B, C, Hs, W] written at out_row_stride
 * rows per image so it can land inside another extended buffer. */
PDENET_API int pdenet_slab_step_fwd_f64(const double *u_ext, const double *kernels, const double *symnet, double *out, int B,
                             int C, int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k,
                             int N, int L, double dt, void *stream);
PDENET_API int pdenet_slab_step_fwd_f32(const float *u_ext, const float *kernels, const float *symnet, float *out, int B,
                             int C, int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k,
                             int N, int L, double dt, void *stream);

/* Same forward slab step with the halo exchange fused in: besides its own rows the kernel stores
 * its first / last `mirror_rows` output rows straight into the ring neighbours' output buffers
 * (peer_up_out / peer_down_out: device pointers to the neighbours' buffers of the SAME layout,
 * mapped into this process -- CUDA IPC / symmetric memory over NVLink), where they are the
 * neighbours' bottom / top ghost rows.  The caller synchronises the ranks between steps. */
PDENET_API int pdenet_slab_step_fwd_p2p_f64(const double *u_ext, const double *kernels, const double *symnet, double *out,
                                 double *peer_up_out, double *peer_down_out, int mirror_rows, int B, int C, int Hs,
                                 int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N, int L,
                                 double dt, void *stream);
PDENET_API int pdenet_slab_step_fwd_p2p_f32(const float *u_ext, const float *kernels, const float *symnet, float *out,
                                 float *peer_up_out, float *peer_down_out, int mirror_rows, int B, int C, int Hs,
                                 int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N, int L,
                                 double dt, void *stream);

/* Reverse of one slab step.  u_ext: state at the step's input, [B,C,Hs+4r,W] with 2r ghost rows
 * on each side (r = (k-1)/2); lam_ext: adjoint of the step's output INCLUDING this step's loss
 * seed, [B,C,Hs+2r,W] with r ghost rows on each side.  Writes the adjoint of the step's input for
 * the Hs owned rows into lam_out (rows [out_row_offset, out_row_offset+Hs) of a buffer with
 * out_rows_per_image rows per image) and accumulates this rank's filter / SymNet gradient partials
 * (owned rows only) in the workspace: call _begin once, _step_bwd per step (last to first), _end once,
 * then sum grad_kernels / grad_symnet over the ranks. */
PDENET_API size_t pdenet_slab_bwd_workspace_bytes(int elem_size, int B, int C, int Hs, int W, int k, int N, int L);
PDENET_API int pdenet_slab_bwd_begin(void *workspace, size_t workspace_bytes, int elem_size, int B, int C, int Hs, int W,
                          int k, int N, int L, void *stream);
PDENET_API int pdenet_slab_step_bwd_f64(const double *u_ext, const double *lam_ext, const double *kernels,
                             const double *symnet, double *lam_out, int out_rows_per_image, int out_row_offset,
                             void *workspace, size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N,
                             int L, double dt, void *stream);
PDENET_API int pdenet_slab_step_bwd_f32(const float *u_ext, const float *lam_ext, const float *kernels,
                             const float *symnet, float *lam_out, int out_rows_per_image, int out_row_offset,
                             void *workspace, size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N,
                             int L, double dt, void *stream);
/* pdenet_slab_step_bwd with the adjoint halo exchange fused in (see pdenet_slab_step_fwd_p2p). */
PDENET_API int pdenet_slab_step_bwd_p2p_f64(const double *u_ext, const double *lam_ext, const double *kernels,
                                 const double *symnet, double *lam_out, double *peer_up_out, double *peer_down_out,
                                 int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                 size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                 void *stream);
PDENET_API int pdenet_slab_step_bwd_p2p_f32(const float *u_ext, const float *lam_ext, const float *kernels,
                                 const float *symnet, float *lam_out, float *peer_up_out, float *peer_down_out,
                                 int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                 size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                 void *stream);
/* Feature-keeping slab steps (see pdenet_rollout_fwd_keep): the forward step also stores the C*N
 * derivative features of its OWNED rows, feats [B,C*N,Hs,W]; the reverse step reads them for the
 * owned rows and recomputes only the ghost rows.  Peers may be NULL with mirror_rows = 0. */
PDENET_API int pdenet_slab_step_fwd_keep_f64(const double *u_ext, const double *kernels, const double *symnet, double *out,
                                  double *feats, double *peer_up_out, double *peer_down_out, int mirror_rows, int B,
                                  int C, int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k,
                                  int N, int L, double dt, void *stream);
PDENET_API int pdenet_slab_step_fwd_keep_f32(const float *u_ext, const float *kernels, const float *symnet, float *out,
                                  float *feats, float *peer_up_out, float *peer_down_out, int mirror_rows, int B, int C,
                                  int Hs, int W, int halo, int out_rows_per_image, int out_row_offset, int k, int N,
                                  int L, double dt, void *stream);
PDENET_API int pdenet_slab_step_bwd_kept_f64(const double *u_ext, const double *lam_ext, const double *feats,
                                  const double *kernels, const double *symnet, double *lam_out, double *peer_up_out,
                                  double *peer_down_out, int mirror_rows, int out_rows_per_image, int out_row_offset,
                                  void *workspace, size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N,
                                  int L, double dt, void *stream);
PDENET_API int pdenet_slab_step_bwd_kept_f32(const float *u_ext, const float *lam_ext, const float *feats, const float *kernels,
                                  const float *symnet, float *lam_out, float *peer_up_out, float *peer_down_out,
                                  int mirror_rows, int out_rows_per_image, int out_row_offset, void *workspace,
                                  size_t workspace_bytes, int B, int C, int Hs, int W, int k, int N, int L, double dt,
                                  void *stream);
PDENET_API int pdenet_slab_bwd_end_f64(void *workspace, size_t workspace_bytes, double *grad_kernels, double *grad_symnet,
                            int B, int C, int Hs, int W, int k, int N, int L, void *stream);
PDENET_API int pdenet_slab_bwd_end_f32(void *workspace, size_t workspace_bytes, float *grad_kernels, float *grad_symnet,
                            int B, int C, int Hs, int W, int k, int N, int L, void *stream);

/* ---- training-loss seed fused into one pass (upstream: the MSE data term of the
 * loss closure in train.py) -------------------------------------------------------------------
 * loss = sum_t w_t * mean((states[t] - observed[t])^2) over t < steps, each step n_per_step scalars;
 * grad_states[t] = 2 w_t (states[t] - observed[t]) / n_per_step (the seed pdenet_rollout_bwd takes
 * with PDENET_GRAD_ALL).  weights_host: `steps` host scalars (NULL = all 1).  loss_out: one device
 * double.  Deterministic (fixed-order two-stage reduction).  Workspace: 8 KiB. */
PDENET_API int pdenet_mse_seed_f64(const double *states, const double *observed, const double *weights_host,
                        double *grad_states, double *loss_out, size_t n_per_step, int steps, void *workspace,
                        size_t workspace_bytes, void *stream);
PDENET_API int pdenet_mse_seed_f32(const float *states, const float *observed, const double *weights_host,
                        float *grad_states, double *loss_out, size_t n_per_step, int steps, void *workspace,
                        size_t workspace_bytes, void *stream);

/* ---- row f-2: 1-D piecewise Lagrange interpolation (source term f(u) of PDE-Net 1's
 * SingleNonLinear2d; upstream aTEAM/nn/modules/Interpolation.py::LagrangeInterp with m = 1
 * [recollection]) -------------------------------------------------------------------------------
 * `mesh` cells on [lo, hi], `degree` (1..4) + 1 equispaced nodes per cell with shared end nodes;
 * coef[mesh*degree + 1] = the values at the nodes.  y[i] = f(x[i]); points outside [lo, hi] use the
 * first / last cell's polynomial.  Reverse: gx[i] = gy[i] f'(x[i]) (gx may be NULL),
 * gcoef[q] = sum_i gy[i] * basis_q(x[i]) (overwritten; deterministic, no atomics). */
PDENET_API int pdenet_lagrange1d_fwd_f64(const double *x, const double *coef, double *y, size_t n, double lo, double hi,
                              int mesh, int degree, void *stream);
PDENET_API int pdenet_lagrange1d_fwd_f32(const float *x, const float *coef, float *y, size_t n, double lo, double hi, int mesh,
                              int degree, void *stream);
PDENET_API size_t pdenet_lagrange1d_bwd_workspace_bytes(size_t n, int mesh, int degree);
PDENET_API int pdenet_lagrange1d_bwd_f64(const double *x, const double *gy, const double *coef, double *gx, double *gcoef,
                              size_t n, double lo, double hi, int mesh, int degree, void *workspace,
                              size_t workspace_bytes, void *stream);
PDENET_API int pdenet_lagrange1d_bwd_f32(const float *x, const float *gy, const float *coef, float *gx, float *gcoef, size_t n,
                              double lo, double hi, int mesh, int degree, void *workspace, size_t workspace_bytes,
                              void *stream);

/* ---- row f-2: PDE-Net 1 SingleNonLinear2d, the whole rollout fused (upstream branch `PDE-Net`,
 * SingleNonLinear2d.forward [recollection]) -----------------------------------------------------------
 * u_{t+1} = K0 (*) u_t + dt * ( sum_n coe[n] * (kernels[n] (*) u_t) + f(u_t) ),  coe: N device scalars,
 * f the piecewise Lagrange interpolant of pdenet_lagrange1d_* (fcoef: mesh*degree+1 node values on [lo, hi]).
 * One launch per direction on the cluster-resident small-grid path (see pdenet_linear_small_grid_path): it is
 * available when pdenet_single_nonlinear_fits(...) returns 1 (the 50 x 50 grids the model is used on); for
 * larger grids the entries return PDENET_E_UNSUPPORTED and the caller steps through
 * pdenet_linear_rollout_* + pdenet_lagrange1d_* (what pdenet_b200.SingleNonLinear2d does).
 * Reverse: grad_coe [N], grad_fcoef [mesh*degree+1] (any gradient pointer may be NULL). */
PDENET_API int pdenet_single_nonlinear_fits(int elem_size, int B, int H, int W, int k, int N);
PDENET_API int pdenet_single_nonlinear_rollout_fwd_f64(const double *u0, const double *id_kernel, const double *kernels,
                                            const double *coe, const double *fcoef, double *states, int B, int H, int W,
                                            int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *stream);
PDENET_API int pdenet_single_nonlinear_rollout_fwd_f32(const float *u0, const float *id_kernel, const float *kernels,
                                            const float *coe, const float *fcoef, float *states, int B, int H, int W,
                                            int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *stream);
PDENET_API size_t pdenet_single_nonlinear_rollout_bwd_workspace_bytes(int elem_size, int B, int H, int W, int k, int N,
                                                           int steps, int mesh, int degree);
PDENET_API int pdenet_single_nonlinear_rollout_bwd_f64(const double *u0, const double *states, const double *grad_states,
                                            int grad_mode, const double *id_kernel, const double *kernels,
                                            const double *coe, const double *fcoef, double *grad_u0,
                                            double *grad_id_kernel, double *grad_kernels, double *grad_coe,
                                            double *grad_fcoef, void *workspace, size_t workspace_bytes, int B, int H,
                                            int W, int k, int N, int steps, double dt, double lo, double hi, int mesh,
                                            int degree, void *stream);
PDENET_API int pdenet_single_nonlinear_rollout_bwd_f32(const float *u0, const float *states, const float *grad_states,
                                            int grad_mode, const float *id_kernel, const float *kernels,
                                            const float *coe, const float *fcoef, float *grad_u0, float *grad_id_kernel,
                                            float *grad_kernels, float *grad_coe, float *grad_fcoef, void *workspace,
                                            size_t workspace_bytes, int B, int H, int W, int k, int N, int steps,
                                            double dt, double lo, double hi, int mesh, int degree, void *stream);

/* ---- tile policy ---------------------------------------------------------------------------------
 * The step / data-gradient / filter-gradient kernels exist with large tiles (64 x 32..64 points: the headline
 * grids) and with small tiles (32 x 16 / 64 x 16: grids whose large-tile count would leave SMs idle or badly
 * balanced -- BASELINE configs 1 and 3).  A launch uses the small tiles when it has fewer than `factor` large
 * tiles per resident CTA slot (default 4; environment PDENET_SMALL_TILE_FACTOR; 0 = always large tiles).
 * Results are identical bit for bit (same per-point arithmetic; the partial sums of the parameter gradients are
 * grouped differently, fixed order either way).  Returns the previous factor; a negative argument only queries. */
PDENET_API int pdenet_small_tile_factor(int factor);
/* The filter-gradient kernel also has a MEDIUM tile (64 x 32 points), taken when a launch has too few large tiles
 * but at least `factor` medium ones per CTA slot (BASELINE config 3).  pdenet_filter_grad_tile_level(level) forces
 * its tile for periodic launches: 0 large, 1 small, 2 medium, -1 back to the policy (process-global; returns the
 * previous setting; other values only query): testing / measurement aid. */
PDENET_API int pdenet_filter_grad_tile_level(int level);

/* ---- measurement hooks (bench.py; not part of the reference interface) ------------------------
 * Kernel kinds: 0 = fused forward step, 1 = reverse kernel A (stencils + SymNet reverse),
 * 2 = reverse kernel B1 (data gradient), 3 = reverse kernel B2 (filter gradient),
 * 4 = small kernels (m2k, reductions).
 * The library always counts its launches; with profiling enabled it also brackets every
 * launch with CUDA events on the launching stream.  pdenet_profile_read synchronises those
 * events, returns the summed milliseconds and launch counts per kind since the last read,
 * and resets both.  Process-global, not thread-safe: a measurement aid only. */
#define PDENET_PROF_KINDS 5
PDENET_API int pdenet_profile_enable(int on);
PDENET_API int pdenet_profile_read(double *ms_by_kind, long long *launches_by_kind);

/* ---- measurement helper: register-only FMA loop, returns elapsed ms for `flops` flops -------
 * (used by bench.py for the CUDA-core FLOP roofline denominator). */
PDENET_API int pdenet_fma_peak(int elem_size, int iters, double *flops_out, float *ms_out, void *stream);
/* Register-operand probes of the FMA pipe: variant 0 = x = fma(x, a, b) with
 * loop-invariant a, b (what pdenet_fma_peak runs); 1 = three distinct register operands per FMA,
 * none shared with the neighbouring FMAs; 2 = one operand shared by 4 consecutive FMAs;
 * 3 / 4 = the filter-gradient kernel's register pattern (49 accumulators per thread, one
 * multiplicand shared by 7 consecutive FMAs, window rows re-read from shared memory) with two / one
 * grid columns per lane; 5 = the same 49 accumulators with every operand already in registers and no
 * ragged first / last window rows; 6 = variant 5 plus the per-row window loads from shared memory. */
PDENET_API int pdenet_fma_probe(int elem_size, int iters, int variant, double *flops_out, float *ms_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDENET_B200_H_ */
