// Anisotropic deblurring operator in its spectral form (Hfuncs.py:448-523):
//   Y_c = U1 (G_c .* (V1^T X_c V2)) U2^T ,  G_c[j] = s_sorted[(3*rank(j)+c) mod D^2]  (SURVEY.md headline 6)
// and the fused loss/gradient of main_sampling.py:709-711 through it:
//   R = y - Y(clamp(x_hat)) ; loss = sum R^2 ; g = -2 V1 (G_c .* (U1^T R U2)) V2^T .* 1[-1<=x_hat<=1]
//
// Eight D x D x D products per (chain, channel) (D = 256: 0.27 GFLOP), true fp32 FFMA (max-abs 1e-3
// parity rules out single-pass TF32/bf16), each a 64x64-tile SIMT GEMM whose epilogue carries the
// elementwise work (spectral gain, residual + loss reduction, final clip mask) so no extra pass over
// the images is made.  The SVD factor tables are inputs: torch.svd signs / tie order are
// implementation-defined, so they are built once by the caller and shared with the checker.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace nhmc {

constexpr int TM = 64, TN = 64, TK = 16;

enum { EPI_NONE = 0, EPI_GAIN = 1, EPI_RESID = 2, EPI_GRAD = 3 };

struct GemmArgs {
  const float* A; int64_t sA; int lda_m, lda_k;   // A(m,k) = A[mat*sA + m*lda_m + k*lda_k]
  const float* B; int64_t sB; int ldb_k, ldb_n;   // B(k,n) = B[mat*sB + k*ldb_k + n*ldb_n]
  float* C; int64_t sC;                           // C(m,n) = C[mat*sC + m*N + n]
  int M, N, K, channels;                          // C is M x N, the contraction runs over K
  const float* gain;   // [channels][M][N]             (EPI_GAIN)
  const float* y;      // [mats][M][N] measurement      (EPI_RESID)
  const float* xhat;   // [mats][M][N] unclamped image  (EPI_GRAD)
  double* loss;        // [chains]                      (EPI_RESID)
  int no_clamp;        // skip the final clip and its mask (latent path)
  RedWs ws;
};

template <int EPI, bool CLAMP_B>
__global__ void __launch_bounds__(256)
gemm_kernel(GemmArgs a) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int M = a.M, N = a.N, K = a.K;
  const int mat = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* Ab = a.A + (int64_t)mat * a.sA;
  const float* Bb = a.B + (int64_t)mat * a.sB;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (a.lda_k == 1) { k = idx & (TK - 1); m = idx >> 4; } else { m = idx & (TM - 1); k = idx >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? __ldg(Ab + (int64_t)gm * a.lda_m + (int64_t)gk * a.lda_k) : 0.f;
      int n, kb;
      if (a.ldb_n == 1) { n = idx & (TN - 1); kb = idx >> 6; } else { kb = idx & (TK - 1); n = idx >> 4; }
      const int gn = n0 + n, gkb = k0 + kb;
      float v = (gn < N && gkb < K) ? __ldg(Bb + (int64_t)gkb * a.ldb_k + (int64_t)gn * a.ldb_n) : 0.f;
      if (CLAMP_B && !a.no_clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
      Bs[kb][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int c = mat % a.channels, chain = mat / a.channels;
  double lsum = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        const int64_t off = (int64_t)gm * N + gn;
        float v = acc[i][j];
        if (EPI == EPI_GAIN) v *= __ldg(a.gain + (int64_t)c * M * N + off);
        if (EPI == EPI_RESID) {
          v = __fsub_rn(__ldg(a.y + (int64_t)mat * M * N + off), v);
          lsum += (double)v * (double)v;
        }
        if (EPI == EPI_GRAD) {
          const float xh = __ldg(a.xhat + (int64_t)mat * M * N + off);
          v = -(2.0f * v) * ((a.no_clamp || (xh >= -1.0f && xh <= 1.0f)) ? 1.0f : 0.0f);
        }
        a.C[(int64_t)mat * a.sC + off] = v;
      }
    }
  }
  if (EPI == EPI_RESID) {
    __shared__ double red[32];
    double vsum[1] = {lsum};
    block_sum<double, 1>(vsum, red);
    const int tiles = gridDim.x * gridDim.y;
    const int blk = (c * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (grid_finish<1>(vsum, a.ws.partials + (int64_t)chain * kMaxBlocksPerChain, a.ws.tickets + chain, blk,
                       tiles * a.channels))
      a.loss[chain] = vsum[0];
  }
}

struct Mats { const float *U1, *V1, *U2, *V2; };

// out = L' * X * R' with elementwise epilogue; L', R' given by (pointer, transposed?)
template <int EPI, bool CLAMP>
static void launch(GemmArgs a, int mats, cudaStream_t st) {
  gemm_kernel<EPI, CLAMP><<<dim3((a.N + TN - 1) / TN, (a.M + TM - 1) / TM, mats), 256, 0, st>>>(a);
}

static GemmArgs left(const float* M, bool trans, const float* X, float* out, int D, int C) {
  GemmArgs a{};
  a.A = M; a.sA = 0; a.lda_m = trans ? 1 : D; a.lda_k = trans ? D : 1;
  a.B = X; a.sB = (int64_t)D * D; a.ldb_k = D; a.ldb_n = 1;
  a.C = out; a.sC = (int64_t)D * D; a.M = a.N = a.K = D; a.channels = C;
  return a;
}
static GemmArgs right(const float* X, const float* M, bool trans, float* out, int D, int C) {
  GemmArgs a{};
  a.A = X; a.sA = (int64_t)D * D; a.lda_m = D; a.lda_k = 1;
  a.B = M; a.sB = 0; a.ldb_k = trans ? 1 : D; a.ldb_n = trans ? D : 1;
  a.C = out; a.sC = (int64_t)D * D; a.M = a.N = a.K = D; a.channels = C;
  return a;
}

// =====================================================================================
// Tensor-core path (gemm_tc.cu): every product of the chains below is C = A * B^T with BOTH operands K-major, so
// the chains are ordered right-factor first and alternate transposed / row-major results:
//   S^T = (X R)^T ; W^T = (gain .* (L S))^T ; T = L' W ; out = T R'^T
// (matrix products are associative; the result differs from the reference's left-to-right evaluation only by fp32
// rounding).  Operands travel as (hi, lo) tf32 plane pairs; the factor tables are split (and transposed where the
// chain wants them) by one small kernel per table at the head of every call -- the caller owns all memory, so the
// split tables live in the call's workspace.
// =====================================================================================
namespace {

struct TcWs {                    // floats
  float* tab[8];                 // plane pairs of V2^T, V1^T, U1, U2, U2^T, U1^T, V1, V2     (each 2 * D * D)
  float *xs, *t0, *t1, *rr;      // plane pairs of [mats][D][D]
  int64_t P, DD;
};

static inline size_t tc_ws_floats(int64_t mats, int64_t D) { return (size_t)(16 * D * D + 8 * mats * D * D); }

static TcWs tc_ws(const nhmc_op* op, float* base, int64_t mats, int64_t D) {
  TcWs w; w.DD = D * D; w.P = mats * w.DD;
  float* tb = op->tc_tables ? const_cast<float*>(op->tc_tables) : base;     // prepared once by the caller, or rebuilt per call
  for (int i = 0; i < 8; ++i) w.tab[i] = tb + (int64_t)i * 2 * w.DD;
  float* d = base + 16 * w.DD;
  w.xs = d; w.t0 = d + 2 * w.P; w.t1 = d + 4 * w.P; w.rr = d + 6 * w.P;
  return w;
}

static TcOperand tab_op(const float* p, int rows, int cols) { TcOperand o{p, rows, cols, 0, (int64_t)rows * cols}; return o; }
static TcOperand dat_op(const float* p, int rows, int cols, int64_t mats) {
  TcOperand o{p, rows, cols, (int64_t)rows * cols, mats * rows * cols}; return o;
}

static TcGemm tc_base(TcOperand A, TcOperand B, int mats, int C, float* out, bool planes, bool transposed) {
  TcGemm g{};
  g.A = A; g.B = B; g.mats = mats; g.channels = C; g.epi = TC_EPI_NONE; g.out_transposed = transposed ? 1 : 0;
  g.out_hi = out; g.out_lo = planes ? out + (int64_t)mats * A.rows * B.rows : nullptr;
  g.out_mat_stride = (int64_t)A.rows * B.rows;
  return g;
}

// tables of the forward chain (first = true) or of the adjoint chain
static int tc_tables(const nhmc_op* op, const TcWs& w, bool fwd, cudaStream_t st, bool force = false) {
  const int D = op->img_dim;
  if (op->tc_tables && !force) return NHMC_OK;
  struct { const float* src; int idx; int tr; } f[4] = {{op->V2, 0, 1}, {op->V1, 1, 1}, {op->U1, 2, 0}, {op->U2, 3, 0}},
                                                 b[4] = {{op->U2, 4, 1}, {op->U1, 5, 1}, {op->V1, 6, 0}, {op->V2, 7, 0}};
  for (int i = 0; i < 4; ++i) {
    const auto& t = fwd ? f[i] : b[i];
    if (int e = tc_split_table_launch(t.src, w.tab[t.idx], w.tab[t.idx] + w.DD, D, D, t.tr, st)) return e;
  }
  return NHMC_OK;
}

// chain on split input `in` (plane pair [mats][D][D]):  fwd: U1 (gain .* (V1^T in V2)) U2^T ;  adjoint: V1 (gain .* (U1^T in U2)) V2^T.
// The last product is configured by `last` (epilogue, outputs) and launched here.
static int tc_chain(const nhmc_op* op, const TcWs& w, const float* in, const float* gain, bool fwd, TcGemm last, int mats,
                    cudaStream_t st) {
  const int D = op->img_dim, C = op->channels;
  const int o = fwd ? 0 : 4;
  TcGemm g1 = tc_base(dat_op(in, D, D, mats), tab_op(w.tab[o + 0], D, D), mats, C, w.t0, true, true);       // S^T = (in R)^T
  if (int e = tc_gemm_launch(g1, st)) return e;
  TcGemm g2 = tc_base(tab_op(w.tab[o + 1], D, D), dat_op(w.t0, D, D, mats), mats, C, w.t1, true, true);     // (gain .* (L S))^T
  g2.epi = TC_EPI_GAIN; g2.gain = gain;
  if (int e = tc_gemm_launch(g2, st)) return e;
  TcGemm g3 = tc_base(tab_op(w.tab[o + 2], D, D), dat_op(w.t1, D, D, mats), mats, C, w.t0, true, false);    // T = L' W
  if (int e = tc_gemm_launch(g3, st)) return e;
  last.A = dat_op(w.t0, D, D, mats); last.B = tab_op(w.tab[o + 3], D, D);                                   // out = T R'^T
  last.mats = mats; last.channels = C; last.out_mat_stride = (int64_t)D * D;
  return tc_gemm_launch(last, st);
}

}  // namespace

static bool use_tc(const nhmc_op* op) {
  const int D = op->img_dim, d = op->kind == NHMC_OP_SEPMAT ? D / op->ratio : D;
  return tc_gemm_supported(D, D, D) && d > 0 && d % 4 == 0;
}

size_t deblur_workspace_bytes(const nhmc_op* op, int n_chains) {
  const size_t mats = (size_t)n_chains * op->channels, D = op->img_dim;
  const size_t simt = (size_t)3 * mats * D * D, tc = tc_ws_floats(mats, D);
  return (simt > tc ? simt : tc) * sizeof(float) + 256;
}

// y = U1 (gain .* (V1^T x V2)) U2^T       (Hfuncs.py:65-71 with Vt :493-499, U :501-509)
int deblur_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* ws, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  if (use_tc(op)) {
    const TcWs w = tc_ws(op, reinterpret_cast<float*>(ws), mats, D);
    if (int e = tc_tables(op, w, true, st)) return e;
    if (int e = tc_split_launch(x, w.xs, w.xs + w.P, w.P, 0, st)) return e;
    TcGemm last{}; last.out_hi = y;
    return tc_chain(op, w, w.xs, op->gain, true, last, mats, st);
  }
  float* t0 = reinterpret_cast<float*>(ws);
  float* t1 = t0 + (size_t)mats * D * D;
  launch<EPI_NONE, false>(left(op->V1, true, x, t0, D, C), mats, st);
  GemmArgs g2 = right(t0, op->V2, false, t1, D, C); g2.gain = op->gain;
  launch<EPI_GAIN, false>(g2, mats, st);
  launch<EPI_NONE, false>(left(op->U1, false, t1, t0, D, C), mats, st);
  launch<EPI_NONE, false>(right(t0, op->U2, true, y, D, C), mats, st);
  return NHMC_LAUNCH_CHECK();
}

// out = V1 (gain' .* (U1^T r U2)) V2^T   with gain' = gain (H^T, Hfuncs.py:73-79) or 1/gain (H^+, :81-91)
int deblur_adjoint(const nhmc_op* op, const float* r, float* out, int n_chains, void* ws, cudaStream_t st, bool pinv) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  if (use_tc(op)) {
    const TcWs w = tc_ws(op, reinterpret_cast<float*>(ws), mats, D);
    if (int e = tc_tables(op, w, false, st)) return e;
    if (int e = tc_split_launch(r, w.xs, w.xs + w.P, w.P, 0, st)) return e;
    TcGemm last{}; last.out_hi = out;
    return tc_chain(op, w, w.xs, pinv ? op->gain_pinv : op->gain, false, last, mats, st);
  }
  float* t0 = reinterpret_cast<float*>(ws);
  float* t1 = t0 + (size_t)mats * D * D;
  launch<EPI_NONE, false>(left(op->U1, true, r, t0, D, C), mats, st);
  GemmArgs g2 = right(t0, op->U2, false, t1, D, C); g2.gain = pinv ? op->gain_pinv : op->gain;
  launch<EPI_GAIN, false>(g2, mats, st);
  launch<EPI_NONE, false>(left(op->V1, false, t1, t0, D, C), mats, st);
  launch<EPI_NONE, false>(right(t0, op->V2, true, out, D, C), mats, st);
  return NHMC_LAUNCH_CHECK();
}

__global__ void clamp_copy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int nc) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = nc ? x[e] : fminf(fmaxf(x[e], -1.0f), 1.0f);
}

// Spectral-domain form of the loss.  U1 and U2 are the (square, orthogonal) left singular vectors of the two blur
// matrices, so with Y^ = U1^T Y U2 (constant for a fixed measurement) and Z = gain .* (V1^T Xc V2):
//   loss = |Y - U1 Z U2^T|^2 = |Y^ - Z|^2 ,      g = -2 H^T (Y - H Xc) = -2 V1 (gain .* (Y^ - Z)) V2^T
// -- four products per evaluation instead of eight (the residual never returns to the image domain).  Equal to the
// image-domain evaluation up to the fp32 orthogonality defect of torch.svd's factors (~1e-6 relative).
int deblur_spectral_measurement(const nhmc_op* op, const float* y, float* y_spec, int n_chains, void* ws, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  NHMC_CHECK_ARG(use_tc(op), "spectral measurement needs the tensor-core path (img_dim % 4 == 0)");
  const TcWs w = tc_ws(op, reinterpret_cast<float*>(ws), mats, D);
  if (int e = tc_tables(op, w, false, st)) return e;
  if (int e = tc_split_launch(y, w.xs, w.xs + w.P, w.P, 0, st)) return e;
  TcGemm g1 = tc_base(dat_op(w.xs, D, D, mats), tab_op(w.tab[4], D, D), mats, C, w.t0, true, true);         // S^T = (Y U2)^T
  if (int e = tc_gemm_launch(g1, st)) return e;
  TcGemm g2 = tc_base(tab_op(w.tab[5], D, D), dat_op(w.t0, D, D, mats), mats, C, y_spec, false, false);     // Y^ = U1^T S
  return tc_gemm_launch(g2, st);
}

int deblur_loss_grad_spec(const nhmc_op* op, const float* x_hat, const float* y_spec, float* g, float* xc_out, double* loss,
                          int n_chains, void* workspace, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  NHMC_CHECK_ARG(use_tc(op), "the spectral loss needs the tensor-core path (img_dim % 4 == 0)");
  const RedWs rw = red_ws(workspace, n_chains);
  float* scratch = reinterpret_cast<float*>((char*)workspace + red_ws_aligned(n_chains));
  const TcWs w = tc_ws(op, scratch, mats, D);
  if (int e = tc_tables(op, w, true, st)) return e;
  if (int e = tc_tables(op, w, false, st)) return e;
  if (int e = tc_split_launch(x_hat, w.xs, w.xs + w.P, w.P, op->no_clamp ? 0 : 1, st)) return e;            // clamp(x_hat), split
  TcGemm g1 = tc_base(dat_op(w.xs, D, D, mats), tab_op(w.tab[0], D, D), mats, C, w.t0, true, true);         // S^T = (Xc V2)^T
  if (int e = tc_gemm_launch(g1, st)) return e;
  TcGemm g2 = tc_base(tab_op(w.tab[1], D, D), dat_op(w.t0, D, D, mats), mats, C, w.t1, true, true);         // (gain .* (Y^ - gain .* V1^T S))^T, loss
  g2.epi = TC_EPI_SPEC; g2.gain = op->gain; g2.y = y_spec; g2.loss = loss; g2.ws = rw;
  if (int e = tc_gemm_launch(g2, st)) return e;
  TcGemm g3 = tc_base(tab_op(w.tab[6], D, D), dat_op(w.t1, D, D, mats), mats, C, w.t0, true, false);        // T = V1 W
  if (int e = tc_gemm_launch(g3, st)) return e;
  TcGemm g4 = tc_base(dat_op(w.t0, D, D, mats), tab_op(w.tab[7], D, D), mats, C, g, false, false);          // g = -2 (T V2^T) .* mask
  g4.epi = TC_EPI_GRAD; g4.xhat = x_hat; g4.no_clamp = op->no_clamp;
  if (int e = tc_gemm_launch(g4, st)) return e;
  if (xc_out) {
    const int64_t total = (int64_t)mats * D * D;
    clamp_copy_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x_hat, xc_out, total, op->no_clamp);
  }
  return NHMC_LAUNCH_CHECK();
}

int deblur_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss,
                     int n_chains, void* workspace, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  const RedWs rw = red_ws(workspace, n_chains);
  float* scratch = reinterpret_cast<float*>((char*)workspace + red_ws_aligned(n_chains));
  if (use_tc(op)) {
    // measurement in the image domain: its spectral form is built here (two products), then the four-product evaluation;
    // callers that evaluate many times against one measurement prepare it once (nhmc_op_prepare_measurement)
    const TcWs w = tc_ws(op, scratch, mats, D);
    if (int e = deblur_spectral_measurement(op, y, w.rr, n_chains, scratch, st)) return e;
    return deblur_loss_grad_spec(op, x_hat, w.rr, g, xc_out, loss, n_chains, workspace, st);
  }
  {
    NHMC_CHECK_ARG(((D + TM - 1) / TM) * ((D + TN - 1) / TN) * C <= kMaxBlocksPerChain, "image too large for the deblur loss reduction");
    float* t0 = scratch;
    float* t1 = t0 + (size_t)mats * D * D;
    float* rr = t1 + (size_t)mats * D * D;
    // forward through clamp(x_hat)
    GemmArgs g1 = left(op->V1, true, x_hat, t0, D, C); g1.no_clamp = op->no_clamp;
    launch<EPI_NONE, true>(g1, mats, st);
    GemmArgs g2 = right(t0, op->V2, false, t1, D, C); g2.gain = op->gain;
    launch<EPI_GAIN, false>(g2, mats, st);
    launch<EPI_NONE, false>(left(op->U1, false, t1, t0, D, C), mats, st);
    GemmArgs g4 = right(t0, op->U2, true, rr, D, C); g4.y = y; g4.loss = loss; g4.ws = rw;
    launch<EPI_RESID, false>(g4, mats, st);
    // adjoint of the residual
    launch<EPI_NONE, false>(left(op->U1, true, rr, t0, D, C), mats, st);
    GemmArgs g6 = right(t0, op->U2, false, t1, D, C); g6.gain = op->gain;
    launch<EPI_GAIN, false>(g6, mats, st);
    launch<EPI_NONE, false>(left(op->V1, false, t1, t0, D, C), mats, st);
    GemmArgs g8 = right(t0, op->V2, true, g, D, C); g8.xhat = x_hat; g8.no_clamp = op->no_clamp;
    launch<EPI_GRAD, false>(g8, mats, st);
  }
  if (xc_out) {
    const int64_t total = (int64_t)mats * D * D;
    clamp_copy_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x_hat, xc_out, total, op->no_clamp);
  }
  return NHMC_LAUNCH_CHECK();
}

// =====================================================================================
// Separable matrix operator Y_c = L X_c L^T (bicubic "SRConv", Hfuncs.py:527-607): in the reference's SVD form
// H = U S V^T with the permutation of :558-559 this is L = U_small diag(s) V_small[:, :d]^T, a d x D matrix
// (d = D / stride) applied on both sides; H^T R = L^T R L, H^+ Y = P Y P^T with P = V_small[:, :d] diag(1/s) U_small^T.
// L and P are built by the caller from the reference's own SVD factors.  Tensor-core path: two products per
// application, S^T = (X Rt^T)^T then out = Lt S (Rt, Lt row-major tables whose K dimension is contiguous).
// =====================================================================================
namespace {

// out[mats][m2 x n1] = Lt (m2 x k2) * in[mats] (k2 x k1) * Rt^T (k1 x n1);  Lt given K-major [m2][k2], Rt K-major [n1][k1].
// `in` is a plane pair; st_t = plane-pair scratch for S^T [mats][n1][k2]; `last` carries the epilogue / outputs.
static int tc_sandwich(const float* in, int k2, int k1, const float* Rt, int n1, const float* Lt, int m2, float* st_t,
                       TcGemm last, int mats, int C, cudaStream_t st) {
  TcGemm g1 = tc_base(dat_op(in, k2, k1, mats), tab_op(Rt, n1, k1), mats, C, st_t, true, true);            // S^T [n1][k2]
  if (int e = tc_gemm_launch(g1, st)) return e;
  last.A = tab_op(Lt, m2, k2); last.B = dat_op(st_t, n1, k2, mats);
  last.mats = mats; last.channels = C; last.out_mat_stride = (int64_t)m2 * n1;
  return tc_gemm_launch(last, st);
}

struct SepWs { float *L, *Lt, *P, *in, *s, *rr; int64_t dD; };
static inline size_t sep_ws_floats(int64_t mats, int64_t D, int64_t d) { return (size_t)(6 * d * D + 2 * mats * D * D + 2 * mats * d * D + 2 * mats * d * d); }
static SepWs sep_ws(const nhmc_op* op, float* base, int64_t mats, int64_t D, int64_t d) {
  SepWs w; w.dD = d * D;
  float* tb = op->tc_tables ? const_cast<float*>(op->tc_tables) : base;
  w.L = tb; w.Lt = tb + 2 * w.dD; w.P = tb + 4 * w.dD;
  w.in = base + 6 * w.dD; w.s = w.in + 2 * mats * D * D; w.rr = w.s + 2 * mats * d * D;
  return w;
}

}  // namespace

// out[mats][M x N] = Lm(M x K) * X[mats][K x N]          (Lm given row-major [rows x cols], optionally transposed)
static GemmArgs left_rect(const float* Lm, int rows, int cols, bool trans, const float* X, float* out, int N, int C) {
  GemmArgs a{};
  const int M = trans ? cols : rows, K = trans ? rows : cols;
  a.A = Lm; a.sA = 0; a.lda_m = trans ? 1 : cols; a.lda_k = trans ? cols : 1;
  a.B = X; a.sB = (int64_t)K * N; a.ldb_k = N; a.ldb_n = 1;
  a.C = out; a.sC = (int64_t)M * N; a.M = M; a.N = N; a.K = K; a.channels = C;
  return a;
}
// out[mats][M x N] = X[mats][M x K] * Rm(K x N)
static GemmArgs right_rect(const float* X, int M, const float* Rm, int rows, int cols, bool trans, float* out, int C) {
  GemmArgs a{};
  const int K = trans ? cols : rows, N = trans ? rows : cols;
  a.A = X; a.sA = (int64_t)M * K; a.lda_m = K; a.lda_k = 1;
  a.B = Rm; a.sB = 0; a.ldb_k = trans ? 1 : cols; a.ldb_n = trans ? cols : 1;
  a.C = out; a.sC = (int64_t)M * N; a.M = M; a.N = N; a.K = K; a.channels = C;
  return a;
}

size_t sepmat_workspace_bytes(const nhmc_op* op, int n_chains) {
  const size_t D = op->img_dim, d = op->img_dim / op->ratio, mats = (size_t)n_chains * op->channels;
  const size_t simt = 2 * mats * d * D + mats * d * d, tc = sep_ws_floats(mats, D, d);
  return (simt > tc ? simt : tc) * sizeof(float) + 256;
}

int sepmat_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* ws, cudaStream_t st) {
  const int D = op->img_dim, d = D / op->ratio, C = op->channels, mats = n_chains * C;
  if (use_tc(op)) {
    const SepWs w = sep_ws(op, reinterpret_cast<float*>(ws), mats, D, d);
    if (!op->tc_tables) if (int e = tc_split_table_launch(op->U1, w.L, w.L + w.dD, d, D, 0, st)) return e;
    if (int e = tc_split_launch(x, w.in, w.in + (int64_t)mats * D * D, (int64_t)mats * D * D, 0, st)) return e;
    TcGemm last{}; last.out_hi = y;
    return tc_sandwich(w.in, D, D, w.L, d, w.L, d, w.s, last, mats, C, st);                                   // L (X L^T)
  }
  float* t0 = reinterpret_cast<float*>(ws);                                       // [mats][d x D]
  launch<EPI_NONE, false>(left_rect(op->U1, d, D, false, x, t0, D, C), mats, st);             // L X
  launch<EPI_NONE, false>(right_rect(t0, d, op->U1, d, D, true, y, C), mats, st);             // (L X) L^T
  return NHMC_LAUNCH_CHECK();
}

int sepmat_adjoint(const nhmc_op* op, const float* r, float* out, int n_chains, void* ws, cudaStream_t st, bool pinv) {
  const int D = op->img_dim, d = D / op->ratio, C = op->channels, mats = n_chains * C;
  if (use_tc(op)) {
    const SepWs w = sep_ws(op, reinterpret_cast<float*>(ws), mats, D, d);
    // H^T R = L^T R L: tables K-major = L^T [D][d];  H^+ Y = P Y P^T: tables K-major = P [D][d]
    if (!op->tc_tables) {
      if (pinv) { if (int e = tc_split_table_launch(op->V1, w.P, w.P + w.dD, D, d, 0, st)) return e; }
      else      { if (int e = tc_split_table_launch(op->U1, w.Lt, w.Lt + w.dD, d, D, 1, st)) return e; }
    }
    if (int e = tc_split_launch(r, w.rr, w.rr + (int64_t)mats * d * d, (int64_t)mats * d * d, 0, st)) return e;
    const float* T = pinv ? w.P : w.Lt;
    TcGemm last{}; last.out_hi = out;
    return tc_sandwich(w.rr, d, d, T, D, T, D, w.s, last, mats, C, st);
  }
  float* t0 = reinterpret_cast<float*>(ws);                                       // [mats][D x d]
  if (!pinv) {
    launch<EPI_NONE, false>(left_rect(op->U1, d, D, true, r, t0, d, C), mats, st);             // L^T R
    launch<EPI_NONE, false>(right_rect(t0, D, op->U1, d, D, false, out, C), mats, st);         // (L^T R) L
  } else {
    launch<EPI_NONE, false>(left_rect(op->V1, D, d, false, r, t0, d, C), mats, st);            // P Y
    launch<EPI_NONE, false>(right_rect(t0, D, op->V1, D, d, true, out, C), mats, st);          // (P Y) P^T
  }
  return NHMC_LAUNCH_CHECK();
}

int sepmat_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss,
                     int n_chains, void* workspace, cudaStream_t st) {
  const int D = op->img_dim, d = D / op->ratio, C = op->channels, mats = n_chains * C;
  const RedWs rw = red_ws(workspace, n_chains);
  float* scratch = reinterpret_cast<float*>((char*)workspace + red_ws_aligned(n_chains));
  if (use_tc(op)) {
    const SepWs w = sep_ws(op, scratch, mats, D, d);
    if (!op->tc_tables) if (int e = tc_split_table_launch(op->U1, w.L, w.L + w.dD, d, D, 0, st)) return e;
    if (!op->tc_tables) if (int e = tc_split_table_launch(op->U1, w.Lt, w.Lt + w.dD, d, D, 1, st)) return e;
    if (int e = tc_split_launch(x_hat, w.in, w.in + (int64_t)mats * D * D, (int64_t)mats * D * D, op->no_clamp ? 0 : 1, st)) return e;
    TcGemm f{}; f.epi = TC_EPI_RESID; f.y = y; f.loss = loss; f.ws = rw; f.out_hi = w.rr; f.out_lo = w.rr + (int64_t)mats * d * d;
    if (int e = tc_sandwich(w.in, D, D, w.L, d, w.L, d, w.s, f, mats, C, st)) return e;                     // R = y - L xc L^T, loss
    TcGemm b{}; b.epi = TC_EPI_GRAD; b.xhat = x_hat; b.no_clamp = op->no_clamp; b.out_hi = g;
    if (int e = tc_sandwich(w.rr, d, d, w.Lt, D, w.Lt, D, w.s, b, mats, C, st)) return e;                   // -2 (L^T R L) .* mask
  } else {
    NHMC_CHECK_ARG(((d + TM - 1) / TM) * ((d + TN - 1) / TN) * C <= kMaxBlocksPerChain, "measurement too large for the loss reduction");
    float* t0 = scratch;                                                                 // [mats][d x D]
    float* t1 = t0 + (size_t)mats * d * D;                                               // [mats][D x d]
    float* rr = t1 + (size_t)mats * d * D;                                               // [mats][d x d]
    GemmArgs g1 = left_rect(op->U1, d, D, false, x_hat, t0, D, C); g1.no_clamp = op->no_clamp;
    launch<EPI_NONE, true>(g1, mats, st);                                                // L clamp(X)
    GemmArgs g2 = right_rect(t0, d, op->U1, d, D, true, rr, C); g2.y = y; g2.loss = loss; g2.ws = rw;
    launch<EPI_RESID, false>(g2, mats, st);                                              // R = y - (L X) L^T, loss
    launch<EPI_NONE, false>(left_rect(op->U1, d, D, true, rr, t1, d, C), mats, st);      // L^T R
    GemmArgs g4 = right_rect(t1, D, op->U1, d, D, false, g, C); g4.xhat = x_hat; g4.no_clamp = op->no_clamp;
    launch<EPI_GRAD, false>(g4, mats, st);                                               // -2 (L^T R L) .* mask
  }
  if (xc_out) {
    const int64_t total = (int64_t)mats * D * D;
    clamp_copy_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x_hat, xc_out, total, op->no_clamp);
  }
  return NHMC_LAUNCH_CHECK();
}

// Derived tables, written once into caller-owned memory (include/nhmc.h: nhmc_op_tables_bytes / nhmc_op_prepare_tables)
size_t spectral_tables_bytes(const nhmc_op* op) {
  if (!use_tc(op)) return 0;
  const size_t D = op->img_dim;
  if (op->kind == NHMC_OP_DEBLUR2D) return 16 * D * D * sizeof(float);
  if (op->kind == NHMC_OP_SEPMAT) return 6 * (D / op->ratio) * D * sizeof(float);
  return 0;
}

int spectral_prepare_tables(const nhmc_op* op, float* tables, cudaStream_t st) {
  const int D = op->img_dim;
  nhmc_op tmp = *op;
  tmp.tc_tables = nullptr;
  if (op->kind == NHMC_OP_DEBLUR2D) {
    const TcWs w = tc_ws(&tmp, tables, 1, D);
    if (int e = tc_tables(&tmp, w, true, st, true)) return e;
    return tc_tables(&tmp, w, false, st, true);
  }
  if (op->kind == NHMC_OP_SEPMAT) {
    const int d = D / op->ratio;
    const SepWs w = sep_ws(&tmp, tables, 1, D, d);
    if (int e = tc_split_table_launch(op->U1, w.L, w.L + w.dD, d, D, 0, st)) return e;
    if (int e = tc_split_table_launch(op->U1, w.Lt, w.Lt + w.dD, d, D, 1, st)) return e;
    return tc_split_table_launch(op->V1, w.P, w.P + w.dD, D, d, 0, st);
  }
  return NHMC_OK;
}

}  // namespace nhmc
