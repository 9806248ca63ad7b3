/*
 * nhmc.h -- C ABI of the B200 (sm_100a) noise-space HMC kernels.
 *
 * The reference (Sunsett5/Noise-space-HMC) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b); every entry point below replaces a group of implicit ATen launches
 * on the `--algo hmc` / `--algo hmc_latent` path, cited per function as file:line of the
 * upstream tree.  INTEGRATION.md shows the ctypes stub a maintainer of the reference adds.
 *
 * Conventions
 *  - every tensor argument is a raw DEVICE pointer plus explicit sizes; `stream` is a
 *    cudaStream_t (pass torch's current stream); nothing synchronises the device;
 *  - the caller owns every buffer; the library allocates no device memory, keeps no mutable state between calls
 *    (no setters, no environment switches: execution-path choices travel in a per-call `flags` word) and is
 *    re-entrant: any number of host threads may call it concurrently as long as calls that run at the same time
 *    on one device use different workspaces;
 *  - return value: 0 = OK, otherwise an NHMC_ERR_* code; nhmc_last_error() gives the text
 *    (thread-local); no C++ exception crosses the boundary;
 *  - per-chain scalars live in a device array of `nhmc_chain` (one entry per chain) so that
 *    chains of one batch can sit at different epochs / step sizes / noise levels;
 *  - images are fp32, contiguous [B, C, H, W]; N = C*H*W elements per chain;
 *  - "workspace": device scratch, size from the matching *_workspace_bytes(); it must be
 *    zero-filled ONCE after allocation (reduction tickets live there and reset themselves).
 *  - sm_100a only; there is no CPU or other-arch fallback.
 */
#ifndef NHMC_H
#define NHMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NHMC_OK 0
#define NHMC_ERR_ARG 1
#define NHMC_ERR_CUDA 2
#define NHMC_ERR_UNSUPPORTED 3

typedef void* nhmc_stream_t; /* cudaStream_t */

/* element types of the score-network output / its cotangent */
#define NHMC_F32 0
#define NHMC_BF16 1
#define NHMC_F16 2

int nhmc_version(void);
const char* nhmc_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Per-chain sampler state (device array, one per chain).  Mirrors the python locals of
 * hmc() (main_sampling.py:660-689) and hmc_latent() (main_sampling_latent.py:623-646).
 * Doubles first, then 4-byte fields: identical layout for C, ctypes and numpy.
 * ---------------------------------------------------------------------------------------- */
typedef struct nhmc_chain {
  double epsilon;      /* current step size (python float in the reference)            */
  double tau;          /* current trajectory length                                     */
  double sigma_y;      /* noise level used by the running iteration                     */
  double sigma_0;      /* measurement noise, already doubled (main_sampling.py:348)     */
  double sigma_y_init; /* latent path: opt.sigma_y (main_sampling_latent.py:625)        */
  double sum_x2[2];    /* sum x^2 : [0] current state, [1] proposal (fp64 accumulate)   */
  double sum_p2[2];    /* sum p^2 : [0] at the momentum draw, [1] end of trajectory     */
  double loss[2];      /* sum r^2 : [0] at the current state, [1] at the proposal       */
  float H[2];          /* fp32 totals, reference order prior + data + kinetic           */
  float prob;          /* min(1, exp(-(H[1]-H[0])))                                     */
  float u;             /* uniform draw consumed by the last decision                    */
  int32_t epoch;
  int32_t accepted;
  int32_t rejected;
  int32_t n_collected; /* samples stored so far                                         */
  int32_t iters;       /* HMC iterations run                                            */
  int32_t accept;      /* decision of the last iteration (0/1)                          */
  int32_t collect;     /* 1 if the last iteration stored a sample                       */
  int32_t slot;        /* ring-buffer slot written by the last iteration, or -1         */
  int32_t active;      /* 0 once the chain's schedule is complete                       */
  int32_t reserved;    /* leapfrog step index l inside the running trajectory (diagonal-mass variant)   */
} nhmc_chain;

size_t nhmc_sizeof_chain(void);

#define NHMC_SCHED_PIXEL 0  /* main_sampling.py:681-749        */
#define NHMC_SCHED_LATENT 1 /* main_sampling_latent.py:646-733 */
#define NHMC_SCHED_COND 2   /* hmc_test_conditioning, main_sampling.py:776-894 (adaptive diagonal mass) */

/* Top of an iteration: sigma_y / tau / epsilon for this epoch.
 * Replaces main_sampling.py:683-689 (pixel).  The latent schedule sets sigma_y on accept
 * (done by nhmc_accept), so this is a no-op for NHMC_SCHED_LATENT except iters bookkeeping. */
int nhmc_begin_iteration(nhmc_chain* chains, int n_chains, int sched, nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1/K6 fused leapfrog (main_sampling.py:697,702,706-707,713,715,717)
 *   force F = x + g/(2 sigma_y^2),  g = g_a (+ g_b if non-NULL)
 *   NHMC_LF_FIRST : energies of (x_in, p) -> sum_x2[0], sum_p2[0], loss[0] = loss_in;
 *                   p -= (eps/2) F ; x_out = x_in + eps/m * p
 *   NHMC_LF_MIDDLE: p -= eps F     ; x_out = x_in + eps/m * p        (x_out may alias x_in)
 *   NHMC_LF_LAST  : p -= eps F ; p += (eps/2) F ; energies of (x_in, p) -> [1], loss[1] = loss_in
 * fp32 arithmetic in the reference's operation order with no FMA contraction; energy sums
 * are accumulated in fp64 in-register and reduced warp-shuffle + one block reduce + a
 * deterministic last-block pass.  loss_in: device double[B] (may be NULL for MIDDLE).
 * ---------------------------------------------------------------------------------------- */
#define NHMC_LF_FIRST 0
#define NHMC_LF_MIDDLE 1
#define NHMC_LF_LAST 2
size_t nhmc_leapfrog_workspace_bytes(int n_chains);
int nhmc_leapfrog(int mode, const float* x_in, float* x_out, float* p, const float* g_a,
                  const float* g_b, const double* loss_in, nhmc_chain* chains, double mass,
                  int n_chains, int64_t n_per_chain, void* workspace, nhmc_stream_t stream);

/* Diagonal-mass variant of the fused leapfrog step (hmc_test_conditioning, main_sampling.py:776-894): inv_mass [B, N]
 * fp32 is the per-element inverse mass; the kinetic sums become sum(inv_mass * p^2) (use mass = 1 in nhmc_accept), the
 * drift x + (eps * p) * inv_mass (:835).  wf_mean / wf_m2 ([B, N] fp32, zero-filled by the caller at the top of every
 * iteration; both NULL: no running variance): in the MIDDLE / LAST modes the Welford mean and M2 of the trajectory's
 * positions (:845-849) are updated in the same pass for chains with epoch - 5 > 13, with step count
 * chains[b].reserved + 1 (reset by nhmc_begin_iteration, advanced here after every MIDDLE / LAST step). */
int nhmc_leapfrog_diag(int mode, const float* x_in, float* x_out, float* p, const float* g_a, const float* g_b,
                       const double* loss_in, nhmc_chain* chains, const float* inv_mass, float* wf_mean, float* wf_m2,
                       int n_chains, int64_t n_per_chain, void* workspace, nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2 DDIM epilogue forward (algos/unconditional.py:17-28; formula of functions/denoising.py:62)
 *   x0raw = (xt - c1*eps)/sa ; mask = (-1 <= x0raw <= 1) ; xt' = ss*clamp(x0raw,-1,1) + c2*eps
 * with c1 = sqrt(1-abar_t), sa = sqrt(abar_t), c2 = sqrt(1-abar_next), ss = sqrt(abar_next),
 * computed by the caller in fp32 exactly as the reference does.  eps is the network output
 * [B, eps_channels, H, W] of which the first C channels are used (unconditional.py:17-18):
 * eps_chain_stride = eps_channels*H*W elements.  mask_out: 1 bit per element, packed per
 * 128-element group as 4 uint32 (word j, bit l <-> element 4*l + j); (N+127)/128*16 bytes
 * per chain; may be NULL.
 * ---------------------------------------------------------------------------------------- */
size_t nhmc_ddim_mask_bytes(int64_t n_per_chain);
int nhmc_ddim_fwd(const float* xt, const void* eps, int eps_dtype, int64_t eps_chain_stride,
                  float* xt_next, uint32_t* mask_out, float c1, float sa, float c2, float ss,
                  int n_chains, int64_t n_per_chain, nhmc_stream_t stream);

/* K3 DDIM epilogue backward (autograd of the above), g' = g_a (+ g_b):
 *   g_xt = ss*g'*mask/sa ; g_eps = c2*g' - c1*g_xt
 * g_eps_out: [B, eps_channels, H, W] cotangent for the network VJP (channels >= C are never
 * written: zero them once); may be NULL (latent path: eps carries no gradient). */
int nhmc_ddim_bwd(const float* g_a, const float* g_b, const uint32_t* mask, float* g_xt,
                  void* g_eps_out, int eps_dtype, int64_t eps_chain_stride, float c1, float sa,
                  float c2, float ss, int n_chains, int64_t n_per_chain, nhmc_stream_t stream);

/* Latent path: the clip of the DDIM output that feeds the decoder, zc = clamp(z, -1, 1) (main_sampling_latent.py:651;
 * zc_copy nullable: a second destination, the sample the sampler keeps), and its gradient mask behind the decoder's VJP,
 * out = g * 1[-1 <= z <= 1] with g in the decoder's dtype and out fp32 (autograd of clamp is inclusive). */
int nhmc_clip_fwd(const float* z, float* zc, float* zc_copy, int64_t n, nhmc_stream_t stream);
int nhmc_clip_bwd(const void* g, int g_dtype, const float* z, float* out, int64_t n, nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4/K5 measurement operators (obs_functions/Hfuncs.py).  All table pointers are DEVICE
 * pointers owned by the caller.
 * ---------------------------------------------------------------------------------------- */
#define NHMC_OP_INPAINT 0  /* Hfuncs.py:119-154  */
#define NHMC_OP_SR 1       /* Hfuncs.py:180-234  */
#define NHMC_OP_HDR 2      /* Hfuncs.py:406-445  */
#define NHMC_OP_DEBLUR2D 3 /* Hfuncs.py:448-523  */
#define NHMC_OP_PHASE 4    /* Hfuncs.py:318-366 + fastmri_utils.py:67-114 */
#define NHMC_OP_DENOISE 5  /* Hfuncs.py:157-177  (H = identity)                */
#define NHMC_OP_COLOR 6    /* Hfuncs.py:655-695  (per-pixel weighted grey value) */
#define NHMC_OP_CS 8       /* Hfuncs.py:611-651  (Walsh-Hadamard compressive sensing: kept_pix = pixel permutation, n_kept = D*D/ratio) */
#define NHMC_OP_SEPMAT 7   /* Hfuncs.py:527-607  (SRConv: Y_c = L X_c L^T, L = [D/ratio, D] in U1, P = [D, D/ratio] in V1) */

typedef struct nhmc_op {
  int32_t kind;
  int32_t channels;
  int32_t img_dim;
  int32_t ratio;              /* SR: block size                                              */
  int32_t n_kept;             /* inpaint: kept pixels                                        */
  int32_t pad;                /* phase: zero padding per side (64)                           */
  int32_t no_clamp;           /* 1: loss_grad skips clamp(x_hat,-1,1) and its mask (latent   */
                              /*    path: the decoded image is not clipped,                  */
                              /*    main_sampling_latent.py:651)                             */
  int32_t reserved;
  const int32_t* kept_pix;    /* inpaint: [n_kept] ascending pixel indices                   */
  const int32_t* pix_to_kept; /* inpaint: [D*D] rank among kept pixels, -1 if dropped        */
  const float* U1;            /* deblur2d: [D,D] row-major, as Hfuncs.Deblurring2D.U_small1  */
                              /* (Gaussian `Deblurring`, Hfuncs.py:236-316: U1 = U2, V1 = V2) */
  const float* V1;
  const float* U2;
  const float* V2;
  const float* gain;          /* deblur2d: [C,D,D] spectral gain D_c (SURVEY.md headline 6)  */
  const float* gain_pinv;     /* deblur2d: [C,D,D] 1/D_c where non-zero else 0               */
  float coef[12];             /* color: [0..2] forward weights w_c (y = sum_c w_c x_c),      */
                              /*        [3..5] adjoint weights, [6..8] pseudo-inverse weights */
  const float* tc_tables;     /* optional: the derived tables nhmc_op_prepare_tables wrote -- deblur2d / sepmat: tf32  */
                              /* hi/lo splits of U1,V1,U2,V2 and their transposes for the tensor-core GEMMs; phase:     */
                              /* the FFT twiddle table W_L^j (float2[L]).  NULL: every call rebuilds them               */
} nhmc_op;

size_t nhmc_sizeof_op(void);

/* measurement elements per chain (M) for this operator */
int64_t nhmc_op_meas_size(const nhmc_op* op);
size_t nhmc_op_workspace_bytes(const nhmc_op* op, int n_chains);
/* Derived operator tables (caller-owned, written once): bytes needed (0: the operator has none) and the kernel launches
 * that fill them; afterwards set op->tc_tables = tables. */
size_t nhmc_op_tables_bytes(const nhmc_op* op);
int nhmc_op_prepare_tables(const nhmc_op* op, void* tables, nhmc_stream_t stream);

/* y = H(x) in the reference's layout (no clamp). Replaces H_functions.H (Hfuncs.py:65-71). */
int nhmc_op_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* workspace,
                    nhmc_stream_t stream);
/* out = H^T(r) for the linear operators (Hfuncs.py:73-79). */
int nhmc_op_adjoint(const nhmc_op* op, const float* r, float* out, int n_chains, void* workspace,
                    nhmc_stream_t stream);
/* out = H^+(y) (Hfuncs.py:81-91; hdr: identity; phase: |ifft2c| cropped). */
int nhmc_op_pinv(const nhmc_op* op, const float* y, float* out, int n_chains, void* workspace,
                 nhmc_stream_t stream);
/* Fused loss + gradient of main_sampling.py:709-711 w.r.t. x_hat (the DDIM output):
 *   xc = clamp(x_hat,-1,1); r = y - H(xc); loss[b] = sum r^2 (fp64);
 *   g = dloss/dx_hat = -2 J_H(xc)^T r * 1[-1 <= x_hat <= 1].
 * xc_out (nullable) receives the clamped image (the sample the reference keeps). */
int nhmc_op_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g,
                      float* xc_out, double* loss, int n_chains, void* workspace,
                      nhmc_stream_t stream);

/* Deblurring operators (kind NHMC_OP_DEBLUR2D on the tensor-core path), for callers that evaluate many times against ONE
 * measurement (a sampler run): with the orthogonal left factors U1, U2, |Y - U1 Z U2^T|^2 = |U1^T Y U2 - Z|^2, so the
 * residual never has to return to the image domain.  prepare_measurement writes Y^ = U1^T Y U2 (two products, same
 * shape as y) once; loss_grad_prepared then needs four products per evaluation instead of the eight of
 * nhmc_op_loss_grad (which does the same internally with six).  Results equal nhmc_op_loss_grad's up to the fp32
 * orthogonality defect of the SVD factors.  NHMC_ERR_UNSUPPORTED for operators without a prepared form. */
int nhmc_op_prepare_measurement(const nhmc_op* op, const float* y, float* y_prepared, int n_chains, void* workspace,
                                nhmc_stream_t stream);
int nhmc_op_loss_grad_prepared(const nhmc_op* op, const float* x_hat, const float* y_prepared, float* g, float* xc_out,
                               double* loss, int n_chains, void* workspace, nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7 Metropolis decision + schedule update, on device (main_sampling.py:697,717-731,743-749;
 * latent main_sampling_latent.py:676-733).  u: device float[B] uniform draws.
 * Writes H[], prob, accept, collect, slot and advances epoch/rejected/epsilon/tau/sigma_y.
 * n_slots = capacity of the per-chain sample ring buffer (pixel 20, latent 10).
 * ---------------------------------------------------------------------------------------- */
int nhmc_accept(nhmc_chain* chains, const float* u, double mass, int sched, int n_slots,
                int n_chains, nhmc_stream_t stream);
/* Apply the decision: x <- x_prop where accepted; store the sample.
 * pixel : accepted & collect -> samples[b][slot] = x_hat[b]
 * latent: collect -> samples[b][slot] = x_keep[b] (previous accepted x_hat, one accept
 *         behind, main_sampling_latent.py:705-713); accepted -> x_keep[b] = x_hat[b].
 * samples: [B, n_slots, N]; x_keep may be NULL for the pixel schedule. */
int nhmc_commit(const nhmc_chain* chains, float* x, const float* x_prop, const float* x_hat,
                float* x_keep, float* samples, int sched, int n_slots, int n_chains,
                int64_t n_per_chain, nhmc_stream_t stream);

/* Optional: keep (loss, gradient) of the CURRENT state per chain so that the
 * evaluation at the top of each iteration (main_sampling.py:693-695, identical to the last evaluation
 * of the previous accepted trajectory) is not repeated: accepted chains take the proposal's
 * g = g_a (+ g_b) and loss, rejected chains keep theirs.  Call after nhmc_accept. */
int nhmc_commit_grad(const nhmc_chain* chains, float* g_state, const float* g_a, const float* g_b,
                     double* loss_state, const double* loss_prop, int n_chains, int64_t n_per_chain,
                     nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K9 U-Net epilogues: GroupNorm(32 groups, eps 1e-5, fp32 statistics) fused with its
 * follow-up (nn.py:17-19; unet_ffhq.py:253-257,277-280,319-323,356,683-685).
 *   NHMC_GN_PLAIN : y = gn(x)                          (attention norm)
 *   NHMC_GN_SILU  : y = silu(gn(x))                    (in_layers / out)
 *   NHMC_GN_FILM_SILU : y = silu(gn(x)*(1+scale)+shift) (out_layers, use_scale_shift_norm)
 * x, y: [B, C, HW] contiguous (NCHW) or channels-last [B, HW, C] (layout flag), dtype
 * f32/bf16/f16; gamma/beta fp32 [C]; scale/shift fp32 rows of C values, sample b at
 * scale + b*film_stride (FiLM only; film_stride 0 means C: this lets both point into one
 * [B, 2C] embedding projection, unet_ffhq.py:319-321, without a split copy); in_bias fp32 [C] or
 * NULL: the tensor that is normalised is x + in_bias[b*in_bias_stride + c] -- the bias of the
 * convolution that produced x (in_bias_stride 0: one vector for all samples) and, for ResBlocks
 * without scale-shift norm (ldm openaimodel.py:267-276), the per-sample embedding projection
 * (in_bias_stride >= C), folded in here instead of a separate broadcast-add pass.
 * stats_out: fp32 [B, G, 2] (mean, rstd) saved for the backward.
 * ---------------------------------------------------------------------------------------- */
#define NHMC_GN_PLAIN 0
#define NHMC_GN_SILU 1
#define NHMC_GN_FILM_SILU 2
#define NHMC_LAYOUT_NCHW 0
#define NHMC_LAYOUT_NHWC 1
/* GroupNorm scratch (fp64 accumulators, reduction tickets, barrier counters, the timeout flag): CALLER-OWNED device
 * memory of nhmc_gn_workspace_bytes(max_batch) bytes, 16-byte aligned, zero-filled ONCE after allocation (everything in
 * it resets itself).  Calls that may run concurrently (other streams / lanes / host threads) need separate workspaces;
 * calls ordered on one stream share one.  `flags`: 0 = defaults, or a combination of */
#define NHMC_F_NO_PDL (1u << 6)              /* plain launches instead of programmatic dependent launch (also pool2 / bias_add) */
#define NHMC_GN_F_NO_RESIDENT (1u << 0)      /* channels-last: statistics + apply kernels only                              */
#define NHMC_GN_F_NO_CLUSTER (1u << 1)       /* no thread-block-cluster variant of the shared-memory-resident kernel          */
#define NHMC_GN_F_SLICE_BWD (1u << 2)        /* resident-by-channel-slice kernel for large backward tensors (off by default)   */
#define NHMC_GN_F_SLICE_ALL (1u << 3)        /* ... and for large forward tensors                                              */
#define NHMC_GN_F_NO_SOLO (1u << 4)          /* no one-CTA-per-unit kernel on the small maps                                   */
#define NHMC_GN_F_NO_SOLO_CLUSTER (1u << 5)  /* no cluster-per-unit kernel on the 64x64 maps                                   */
#define NHMC_GN_F_LANES_SHIFT 8              /* bits 8-10: (number of lanes - 1) this call shares the chip with (sizing only)  */
size_t nhmc_gn_workspace_bytes(int max_batch);
int nhmc_gn_fwd(int mode, int layout, int dtype, const void* x, void* y, const float* gamma,
                const float* beta, const float* scale, const float* shift, int64_t film_stride,
                const float* in_bias, int64_t in_bias_stride, float* stats_out,
                int batch, int channels, int64_t hw, int groups, float eps,
                void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream);
/* dx (and optionally dscale/dshift [B,C] fp32) given dy; weights are frozen on this path, so
 * no dgamma/dbeta.  dx_add (nullable, same dtype/layout as dx, channels-last only): a tensor added to
 * dx in the same pass -- the gradient arriving through the ResBlock's skip branch
 * (unet_ffhq.py:327), which autograd would otherwise add with a separate read-read-write pass.
 * dx_add2 (channels-last, needs dx_add): a second tensor added in the same pass -- the gradient that reaches a block
 * input through the U-Net's long skip connection (unet_ffhq.py:729-731).
 *
 * Channels-last execution: when one sample's slab fits in the shared memory of the chip (every
 * FFHQ-256 site except the backward of the two 256-channel 256x256 sites), ONE kernel reads the
 * tensor once into shared memory (TMA bulk copies), reduces, synchronises the CTAs of a sample
 * (cluster barrier, or a per-sample arrival counter under a cooperative launch) and normalises out of shared
 * memory: HBM traffic equals the algorithmic bytes.  Otherwise a statistics kernel + an apply kernel (two reads). */
int nhmc_gn_bwd(int mode, int layout, int dtype, const void* x, const void* dy, void* dx,
                const float* gamma, const float* beta, const float* scale, const float* shift,
                int64_t film_stride, const float* in_bias, int64_t in_bias_stride, const float* stats,
                float* dscale, float* dshift,
                const void* dx_add, const void* dx_add2, int batch, int channels, int64_t hw, int groups,
                void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream);
/* GroupNorm over the VIRTUAL concatenation cat([x0, x1], channel) (the skip connections of the output
 * blocks, unet_ffhq.py:729-731): x0 [B, HW, c0] and x1 [B, HW, c1] are separate dense channels-last
 * tensors, y is the dense [B, HW, c0+c1] result the next convolution reads; gamma/beta/scale/shift index
 * the concatenated channel.  The backward writes the two input gradients separately and can add a tensor
 * to each in the same pass (add0/add1, independently nullable).  Saves the concat pass forward and the
 * split pass backward. */
int nhmc_gn_cat_fwd(int mode, int dtype, const void* x0, const void* x1, int c0, int c1, void* y,
                    const float* gamma, const float* beta, const float* scale, const float* shift,
                    int64_t film_stride, float* stats_out, int batch, int64_t hw, int groups, float eps,
                    void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream);
int nhmc_gn_cat_bwd(int mode, int dtype, const void* x0, const void* x1, int c0, int c1, const void* dy,
                    void* dx0, void* dx1, const float* gamma, const float* beta, const float* scale,
                    const float* shift, int64_t film_stride, const float* stats, const void* add0,
                    const void* add1, int batch, int64_t hw, int groups,
                    void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream);
/* Copies the workspace's timeout flag to *flag_out (device or pinned-host int32) on `stream`, without synchronising:
 * non-zero = a barrier wait of a shared-memory-resident GroupNorm kernel timed out since the workspace was last
 * zero-filled (the results of that call are invalid; zero-fill the workspace again before reuse).  The global-barrier
 * kernel is launched cooperatively, so this is a backstop, not an expected event. */
int nhmc_gn_check(const void* workspace, size_t workspace_bytes, int32_t* flag_out, nhmc_stream_t stream);

/* K10 attention softmax epilogue (unet_ffhq.py:418-425): in place over the last dim of
 * logits [rows, T] (dtype f32/bf16/f16), computed in fp32; logits are multiplied by
 * `scale` first (the reference pre-scales q and k by ch^-1/4 each). */
int nhmc_softmax_fwd(int dtype, const void* logits, void* probs, float scale, int64_t rows,
                     int cols, nhmc_stream_t stream);
/* dlogits = scale * probs * (dprobs - sum(dprobs*probs)) */
int nhmc_softmax_bwd(int dtype, const void* probs, const void* dprobs, void* dlogits, float scale,
                     int64_t rows, int cols, nhmc_stream_t stream);

/* K10b fused attention forward for passes that keep no gradient state (the latent U-Net of hmc_latent:
 * openaimodel.py:545-575 QKVAttentionLegacy == unet_ffhq.py:408-425): out = softmax(scale * q k^T) v per (sample, head),
 * logits and probabilities never written to memory (tcgen05: S and O in tensor memory, P through shared memory).
 *   qk  [batch, tokens, heads, 2, head_dim]  token-major, q then k of one head adjacent
 *   v   [batch, heads*head_dim, tokens]      channel-major
 *   out [batch, tokens, heads*head_dim]      token-major
 * dtype NHMC_BF16 / NHMC_F16, head_dim 32 or 64, tokens % 8 == 0; fp32 accumulation, fp32 softmax statistics.
 * flags: NHMC_F_NO_PDL.  Anything else returns NHMC_ERR_UNSUPPORTED (callers keep the three-kernel path). */
int nhmc_attn_fwd(int dtype, const void* qk, const void* v, void* out, int batch, int tokens, int heads,
                  int head_dim, float scale, unsigned int flags, nhmc_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * U-Net resampling / skip-concat epilogues, channels-last [B, H, W, C] (unet_ffhq.py:153-211 as used
 * by the up/down ResBlocks :307-316; skip concat :729-731).
 *   gather : out_small[b,i,j,:] = scale * sum of the 2x2 block of in_big   (avg_pool2d fwd, scale 1/4;
 *            nearest-upsample backward, scale 1)
 *   scatter: out_big[b,2i+di,2j+dj,:] = scale * in_small[b,i,j,:]         (nearest-upsample fwd, scale 1;
 *            avg_pool2d backward, scale 1/4)
 *   split2 : backward of cat([a,b], channel): in [rows, c1+c2] -> out1 [rows, c1], out2 [rows, c2]
 * ---------------------------------------------------------------------------------------- */
int nhmc_pool2_gather(int dtype, const void* in_big, void* out_small, int batch, int h_small,
                      int w_small, int channels, float scale, unsigned int flags, nhmc_stream_t stream);
int nhmc_pool2_scatter(int dtype, const void* in_small, void* out_big, int batch, int h_small,
                       int w_small, int channels, float scale, unsigned int flags, nhmc_stream_t stream);
int nhmc_split2(int dtype, const void* in, void* out1, void* out2, int64_t rows, int c1, int c2,
                nhmc_stream_t stream);
/* out[r, c] = a[r, c] (+ b[r, c] (+ c3[r, c])) (+ bias[c]): residual add of the ResBlock (unet_ffhq.py:327) with
 * the biases of the convolutions that produced the addends folded in; the third addend is the second half of a
 * skip 1x1 convolution applied block-wise to a virtual concatenation.  channels-last rows, out may alias a. */
int nhmc_bias_add(int dtype, const void* a, const void* b, const void* c3, const float* bias, void* out,
                  int64_t rows, int channels, unsigned int flags, nhmc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Latent path, first stage: nearest-codebook lookup of the VQ-f4 autoencoder
 * (VQModelInterface.decode -> taming VectorQuantizer2.forward, ldm/models/autoencoder.py:274-282):
 * zq[b,:,p] = z + (e - z) with e = codebook[argmin_k |z[b,:,p] - e_k|^2] (the reference's straight-through
 * form, equal to e up to fp32 rounding), first minimum wins.  z, zq: fp32 [B, dim, pixels]
 * (NCHW); codebook fp32 [n_codes, dim]; idx_out int32 [B, pixels] or NULL.  The gradient is
 * straight-through (identity), so no backward entry point exists.
 * ---------------------------------------------------------------------------------------- */
int nhmc_vq_nearest(const float* z, const float* codebook, float* zq, int32_t* idx_out, int batch,
                    int dim, int64_t pixels, int n_codes, nhmc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NHMC_H */
