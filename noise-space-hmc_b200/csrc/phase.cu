// Phase-retrieval operator (Hfuncs.py:318-366, fft2_m :8-12, fastmri_utils.py:67-114):
//   y = | fftshift( fft2( ifftshift( pad(x, 64) ), norm="ortho" ) ) |          (384 x 384 per channel)
// and the fused loss / gradient of main_sampling.py:709-711 through it:
//   X = F pad(clamp(x_hat)) ; res = |X| - y ; loss = sum res^2 ; g = 2 Re crop F^H( res * X/|X| ) * clipmask
//
// The even-length shifts only multiply X by (-1)^(u+v) (irrelevant under |.| and cancelling between F and
// F^H), so the kernels run plain orthonormal DFTs and index y at (u + L/2, v + L/2) mod L.
//
// FFT: L = r * 2^k (384 = 3*128, 192 = 3*64, 160 = 5*32): one r-point DFT + twiddle stage followed by an
// in-place radix-2 DIT on the 2^k sub-sequences, NC = 4 lines per CTA in shared memory (fp32, twiddles
// from sincospi once per CTA).  The zero padding is exploited: only the D non-zero rows are transformed in
// the row pass and only D output rows are produced by the inverse column pass.  Three kernels per
// loss/gradient: rows fwd -> columns (fwd, |.|-residual, inverse, all in shared memory) -> rows inverse.
#include "common.cuh"

namespace nhmc {

constexpr int NC = 4;          // lines per CTA
constexpr int kFftThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

struct FftGeom { int L, r, N2, logN2; };

static bool fft_geom(int L, FftGeom& g) {
  int n2 = 1, lg = 0;
  while (L % (n2 * 2) == 0) { n2 *= 2; ++lg; }
  g.L = L; g.N2 = n2; g.logN2 = lg; g.r = L / n2;
  return g.r <= 15 && L >= 8 && L % NC == 0;
}

__device__ __forceinline__ void make_twiddles(float2* tw, int L) {
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)L, &s, &c);
    tw[j] = make_float2(c, -s);  // W_L^j = exp(-2 pi i j / L)
  }
}

// In-place (buf -> buf) DFT of NC lines of length L held in shared memory. inverse: conjugate twiddles.
__device__ void fft_lines(float2* buf, float2* tmp, const float2* tw, FftGeom g, bool inverse) {
  const int L = g.L, r = g.r, N2 = g.N2;
  const float sgn = inverse ? -1.0f : 1.0f;
  for (int idx = threadIdx.x; idx < NC * N2; idx += blockDim.x) {
    const int line = idx / N2, n2 = idx - line * N2;
    const float2* src = buf + line * L;
    float2* dst = tmp + line * L;
    const int rev = (int)(__brev((unsigned)n2) >> (32 - g.logN2)) & (N2 - 1);
    for (int k1 = 0; k1 < r; ++k1) {
      float2 acc = make_float2(0.f, 0.f);
      for (int n1 = 0; n1 < r; ++n1) {
        float2 w = tw[((n1 * k1) % r) * N2];
        w.y *= sgn;
        acc = cadd(acc, cmul(src[N2 * n1 + n2], w));
      }
      float2 t = tw[n2 * k1];
      t.y *= sgn;
      dst[k1 * N2 + (g.logN2 ? rev : 0)] = cmul(acc, t);
    }
  }
  __syncthreads();
  for (int s = 1; s <= g.logN2; ++s) {
    const int half = 1 << (s - 1), m = half << 1, tstep = L / m;
    for (int idx = threadIdx.x; idx < NC * r * (N2 / 2); idx += blockDim.x) {
      const int sub = idx / (N2 / 2), j = idx - sub * (N2 / 2);
      const int grp = j / half, pos = j - grp * half;
      const int i0 = sub * N2 + grp * m + pos, i1 = i0 + half;
      float2 w = tw[pos * tstep];
      w.y *= sgn;
      const float2 t = cmul(w, tmp[i1]), u = tmp[i0];
      tmp[i0] = cadd(u, t);
      tmp[i1] = csub(u, t);
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
    const int line = idx / L, k = idx - line * L;
    const int k1 = k % r, k2 = k / r;
    buf[line * L + k] = tmp[line * L + k1 * N2 + k2];
  }
  __syncthreads();
}

struct PhaseArgs {
  const float* x;      // [mats][D][D] image (x_hat or x)
  const float* y;      // [mats][L][L] measurement
  float* out;          // mode dependent
  float* xc_out;       // clamped image or null
  float2* rw;          // [mats][D][L] row-transformed
  float2* cw;          // [mats][D][L] column-inverse-transformed (cropped rows)
  double* loss; double* partials; unsigned int* tickets;
  int D, pad, channels, no_clamp;
  FftGeom g;
};

// rows forward: image rows (zero-padded to L) -> rw
__global__ void __launch_bounds__(kFftThreads)
phase_rows_fwd_kernel(PhaseArgs a) {
  extern __shared__ float2 sm2[];
  const int L = a.g.L, D = a.D;
  float2* buf = sm2; float2* tmp = buf + NC * L; float2* tw = tmp + NC * L;
  const int mat = blockIdx.y, i0 = blockIdx.x * NC;
  make_twiddles(tw, L);
  for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
    const int line = idx / L, c = idx - line * L, i = i0 + line, j = c - a.pad;
    float v = 0.f;
    if (i < D && j >= 0 && j < D) {
      v = a.x[((int64_t)mat * D + i) * D + j];
      if (!a.no_clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
      if (a.xc_out) a.xc_out[((int64_t)mat * D + i) * D + j] = v;
    }
    buf[idx] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft_lines(buf, tmp, tw, a.g, false);
  const float sc = rsqrtf((float)L);
  for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
    const int line = idx / L, v = idx - line * L, i = i0 + line;
    if (i < D) a.rw[((int64_t)mat * D + i) * L + v] = make_float2(buf[idx].x * sc, buf[idx].y * sc);
  }
}

// columns. MODE 0: forward, write |X| to out (y layout). MODE 1: forward, residual, loss, inverse -> cw.
// MODE 2: pinv input: signed y as real input, inverse -> cw.
template <int MODE>
__global__ void __launch_bounds__(kFftThreads)
phase_cols_kernel(PhaseArgs a) {
  extern __shared__ float2 sm2[];
  const int L = a.g.L, D = a.D, h = L / 2;
  float2* buf = sm2; float2* tmp = buf + NC * L; float2* tw = tmp + NC * L;
  const int mat = blockIdx.y, v0 = blockIdx.x * NC;
  make_twiddles(tw, L);
  const float sc = rsqrtf((float)L);
  if (MODE != 2) {
    for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
      const int u = idx / NC, c = idx - u * NC, i = u - a.pad;
      buf[c * L + u] = (i >= 0 && i < D) ? a.rw[((int64_t)mat * D + i) * L + v0 + c] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_lines(buf, tmp, tw, a.g, false);
  }
  double acc = 0.0;
  for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
    const int u = idx / NC, c = idx - u * NC, v = v0 + c;
    const int64_t yoff = ((int64_t)mat * L + (u + h) % L) * L + (v + h) % L;
    if (MODE == 0) {
      const float2 X = buf[c * L + u];
      a.out[yoff] = sqrtf(X.x * X.x + X.y * X.y) * sc;
    } else if (MODE == 1) {
      float2 X = buf[c * L + u];
      X.x *= sc; X.y *= sc;
      const float mag = sqrtf(X.x * X.x + X.y * X.y);
      const float res = mag - a.y[yoff];
      acc += (double)res * (double)res;
      const float f = mag > 0.f ? res / mag : 0.f;      // d|z|/dz at 0 := 0 (torch's sgn convention)
      buf[c * L + u] = make_float2(X.x * f, X.y * f);
    } else {
      const float sg = ((u + v) & 1) ? -1.0f : 1.0f;
      buf[c * L + u] = make_float2(a.y[yoff] * sg, 0.f);
    }
  }
  if (MODE == 0) return;
  __syncthreads();
  fft_lines(buf, tmp, tw, a.g, true);
  for (int idx = threadIdx.x; idx < NC * D; idx += blockDim.x) {
    const int i = idx / NC, c = idx - i * NC;
    const float2 t = buf[c * L + a.pad + i];
    a.cw[((int64_t)mat * D + i) * L + v0 + c] = make_float2(t.x * sc, t.y * sc);
  }
  if (MODE == 1) {
    __shared__ double red[32];
    double vsum[1] = {acc};
    block_sum<double, 1>(vsum, red);
    const int chain = mat / a.channels, ch = mat - chain * a.channels;
    const int per_chain = a.channels * gridDim.x;
    const int blk = ch * gridDim.x + blockIdx.x;
    if (grid_finish<1>(vsum, a.partials + (int64_t)chain * per_chain, a.tickets + chain, blk, per_chain)) a.loss[chain] = vsum[0];
  }
}

// rows inverse. MODE 0: g = 2 Re(.) * clip mask. MODE 1: out = |.| (pinv).
template <int MODE>
__global__ void __launch_bounds__(kFftThreads)
phase_rows_inv_kernel(PhaseArgs a) {
  extern __shared__ float2 sm2[];
  const int L = a.g.L, D = a.D;
  float2* buf = sm2; float2* tmp = buf + NC * L; float2* tw = tmp + NC * L;
  const int mat = blockIdx.y, i0 = blockIdx.x * NC;
  make_twiddles(tw, L);
  for (int idx = threadIdx.x; idx < NC * L; idx += blockDim.x) {
    const int line = idx / L, v = idx - line * L, i = i0 + line;
    buf[idx] = i < D ? a.cw[((int64_t)mat * D + i) * L + v] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  fft_lines(buf, tmp, tw, a.g, true);
  const float sc = rsqrtf((float)L);
  for (int idx = threadIdx.x; idx < NC * D; idx += blockDim.x) {
    const int line = idx / D, j = idx - line * D, i = i0 + line;
    if (i >= D) continue;
    const float2 t = buf[line * L + a.pad + j];
    const int64_t off = ((int64_t)mat * D + i) * D + j;
    if (MODE == 0) {
      const float xh = a.x[off];
      const float m = (a.no_clamp || (xh >= -1.0f && xh <= 1.0f)) ? 1.0f : 0.0f;
      a.out[off] = 2.0f * t.x * sc * m;
    } else {
      a.out[off] = sqrtf(t.x * t.x + t.y * t.y) * sc;
    }
  }
}

static size_t fft_smem(int L) { return (size_t)(2 * NC + 1) * L * sizeof(float2); }

struct PhaseWs { double* partials; unsigned int* tickets; float2* rw; float2* cw; };
static size_t phase_red_bytes(const nhmc_op* op, int n_chains) {
  const int L = op->img_dim + 2 * op->pad;
  size_t b = (size_t)n_chains * op->channels * (L / NC + 1) * sizeof(double) + (size_t)n_chains * sizeof(unsigned int);
  return (b + 255) & ~(size_t)255;
}
size_t phase_workspace_bytes(const nhmc_op* op, int n_chains) {
  const int L = op->img_dim + 2 * op->pad;
  return phase_red_bytes(op, n_chains) + (size_t)2 * n_chains * op->channels * op->img_dim * L * sizeof(float2) + 256;
}
// `extra` = operator scratch behind the common reduction area (see operators.cu)
static PhaseWs phase_ws(const nhmc_op* op, int n_chains, void* extra) {
  const int L = op->img_dim + 2 * op->pad;
  PhaseWs w;
  w.partials = reinterpret_cast<double*>(extra);
  w.tickets = reinterpret_cast<unsigned int*>(w.partials + (size_t)n_chains * op->channels * (L / NC + 1));
  w.rw = reinterpret_cast<float2*>((char*)extra + phase_red_bytes(op, n_chains));
  w.cw = w.rw + (size_t)n_chains * op->channels * op->img_dim * L;
  return w;
}

static int phase_setup(const nhmc_op* op, PhaseArgs& a, size_t& smem) {
  const int L = op->img_dim + 2 * op->pad;
  if (!fft_geom(L, a.g)) {
    set_error("phase: padded size %d is not r*2^k with odd r <= 15", L);
    return NHMC_ERR_UNSUPPORTED;
  }
  a.D = op->img_dim; a.pad = op->pad; a.channels = op->channels; a.no_clamp = op->no_clamp;
  smem = fft_smem(L);
  if (smem > 200 * 1024) { set_error("phase: padded size %d too large for shared memory", L); return NHMC_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(phase_rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phase_cols_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phase_cols_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phase_cols_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phase_rows_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phase_rows_inv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  return NHMC_OK;
}

int phase_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* extra, cudaStream_t st) {
  PhaseArgs a{}; size_t smem;
  if (int e = phase_setup(op, a, smem)) return e;
  const PhaseWs w = phase_ws(op, n_chains, extra);
  const int mats = n_chains * op->channels, L = a.g.L;
  a.x = x; a.out = y; a.rw = w.rw; a.no_clamp = 1;   // H() itself never clamps
  phase_rows_fwd_kernel<<<dim3((a.D + NC - 1) / NC, mats), kFftThreads, smem, st>>>(a);
  phase_cols_kernel<0><<<dim3(L / NC, mats), kFftThreads, smem, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

int phase_pinv(const nhmc_op* op, const float* y, float* out, int n_chains, void* extra, cudaStream_t st) {
  PhaseArgs a{}; size_t smem;
  if (int e = phase_setup(op, a, smem)) return e;
  const PhaseWs w = phase_ws(op, n_chains, extra);
  const int mats = n_chains * op->channels, L = a.g.L;
  a.y = y; a.out = out; a.cw = w.cw;
  phase_cols_kernel<2><<<dim3(L / NC, mats), kFftThreads, smem, st>>>(a);
  phase_rows_inv_kernel<1><<<dim3((a.D + NC - 1) / NC, mats), kFftThreads, smem, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

int phase_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss,
                    int n_chains, void* workspace, cudaStream_t st) {
  PhaseArgs a{}; size_t smem;
  if (int e = phase_setup(op, a, smem)) return e;
  const PhaseWs w = phase_ws(op, n_chains, (char*)workspace + red_ws_aligned(n_chains));
  const int mats = n_chains * op->channels, L = a.g.L;
  a.x = x_hat; a.y = y; a.out = g; a.xc_out = xc_out; a.rw = w.rw; a.cw = w.cw;
  a.loss = loss; a.partials = w.partials; a.tickets = w.tickets;
  phase_rows_fwd_kernel<<<dim3((a.D + NC - 1) / NC, mats), kFftThreads, smem, st>>>(a);
  phase_cols_kernel<1><<<dim3(L / NC, mats), kFftThreads, smem, st>>>(a);
  phase_rows_inv_kernel<0><<<dim3((a.D + NC - 1) / NC, mats), kFftThreads, smem, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

}  // namespace nhmc
