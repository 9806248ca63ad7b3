// Phase-retrieval operator (Hfuncs.py:318-366, fft2_m :8-12, fastmri_utils.py:67-114):
//   y = | fftshift( fft2( ifftshift( pad(x, 64) ), norm="ortho" ) ) |          (384 x 384 per channel)
// and the fused loss / gradient of main_sampling.py:709-711 through it:
//   X = F pad(clamp(x_hat)) ; res = |X| - y ; loss = sum res^2 ; g = 2 Re crop F^H( res * X/|X| ) * clipmask
//
// The even-length shifts only multiply X by (-1)^(u+v) (irrelevant under |.| and cancelling between F and
// F^H), so the kernels run plain orthonormal DFTs and index y at (u + L/2, v + L/2) mod L.
//
// FFT: L = r * 2^k (384 = 3*128, 192 = 3*64, 160 = 5*32): one r-point DFT + twiddle stage, then a Stockham
// autosort FFT of the r sub-sequences of length 2^k in radix-8 / radix-4 passes (2^7 = 8*4*4: three passes instead of the
// seven radix-2 passes of round 1, each butterfly held in registers, no bit reversal; the last pass stores straight into
// the interleaved output order), NC lines per CTA ping-ponging between two shared-memory buffers (fp32, twiddles from
// sincospi once per CTA).  Rows: 4 lines per CTA, columns: 8 (swept: fewer lines = more resident CTAs hiding the pass barriers wins over
// wider global segments).  The zero padding is exploited: only the D non-zero rows are transformed in the row
// pass and only D output rows are produced by the inverse column pass.  Three kernels per loss/gradient: rows fwd ->
// columns (fwd, |.|-residual, inverse, all in shared memory) -> rows inverse.
#include "common.cuh"

namespace nhmc {

#ifndef NHMC_PHASE_NCR
#define NHMC_PHASE_NCR 4
#endif
constexpr int NCR = NHMC_PHASE_NCR;   // lines per CTA, row kernels
#ifndef NHMC_PHASE_NCC
#define NHMC_PHASE_NCC 8
#endif
constexpr int NCC = NHMC_PHASE_NCC;   // lines per CTA, column kernels (measured: 8 columns / 4 rows per CTA are the fastest, profiles/r2_stream_variants.md)
constexpr int kFftThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

struct FftGeom { int L, r, N2, logN2; };

static bool fft_geom(int L, FftGeom& g) {
  int n2 = 1, lg = 0;
  while (L % (n2 * 2) == 0) { n2 *= 2; ++lg; }
  g.L = L; g.N2 = n2; g.logN2 = lg; g.r = L / n2;
  return (g.r == 1 || g.r == 3 || g.r == 5) && L >= 16 && L % 16 == 0 && lg >= 2;
}

__device__ __forceinline__ void make_twiddles(float2* tw, int L, const float2* table = nullptr) {
  if (table) {   // prepared once by nhmc_op_prepare_tables: L2-resident 8 L bytes instead of L sincospi evaluations per CTA
    for (int j = threadIdx.x; j < L; j += blockDim.x) tw[j] = __ldg(table + j);
    return;
  }
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)L, &s, &c);
    tw[j] = make_float2(c, -s);  // W_L^j = exp(-2 pi i j / L)
  }
}

// R-point DFT of v[0..R) in registers; sgn = +1 forward (e^{-i}), -1 inverse.
template <int R>
__device__ __forceinline__ void dft_r(float2 (&v)[R], float sgn) {
  if (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  } else if (R == 4) {
    const float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]), s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
    const float2 jd = make_float2(sgn * d13.y, -sgn * d13.x);           // -i * d13 (forward), +i * d13 (inverse)
    v[0] = cadd(s02, s13); v[2] = csub(s02, s13);
    v[1] = cadd(d02, jd);  v[3] = csub(d02, jd);
  } else {  // R == 8: two radix-4 on even / odd inputs + twiddles W8^k
    float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft_r<4>(e, sgn); dft_r<4>(o, sgn);
    const float h = 0.70710678118654752f;
    const float2 w1 = make_float2(h, -sgn * h), w2 = make_float2(0.f, -sgn), w3 = make_float2(-h, -sgn * h);
    o[1] = cmul(o[1], w1); o[2] = cmul(o[2], w2); o[3] = cmul(o[3], w3);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = cadd(e[k], o[k]); v[k + 4] = csub(e[k], o[k]); }
  }
}

// One Stockham pass of radix R over the sequences q = line * RT + k1 (length N, based at line * LS + k1 * N in src and dst):
// butterfly j reads src[j + t * N/R], twiddles by W_{Ns R}^{t (j mod Ns)}, and writes dst[(j / Ns) Ns R + (j mod Ns) + t Ns].
// LAST: the store goes to the interleaved final order line * LS + k1 + RT * k2.
template <int R, bool LAST, int RT>
__device__ __forceinline__ void stockham_pass(const float2* src, float2* dst, const float2* tw, int L, int N, int Ns, int nseq, float sgn) {
  // N, Ns and R are powers of two: shifts and masks instead of integer divisions
  const int LS = L + 2;
  const int per = N / R, tstep = L / (Ns * R), lper = 31 - __clz(per), lNs = 31 - __clz(Ns);
  for (int idx = threadIdx.x; idx < nseq * per; idx += blockDim.x) {
    const int q = idx >> lper, j = idx & (per - 1);
    const int line = q / RT, k1 = q - line * RT;
    const int base = line * LS + k1 * N;
    const int jm = j & (Ns - 1);
    const float2* in = src + base + j;
    float2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      float2 w = tw[(t * jm) * tstep];
      w.y *= sgn;
      v[t] = t ? cmul(in[t * per], w) : in[0];
    }
    dft_r<R>(v, sgn);
    const int j0 = ((j >> lNs) << lNs) * R + jm;
    if (LAST) {
#pragma unroll
      for (int t = 0; t < R; ++t) dst[line * LS + k1 + RT * (j0 + t * Ns)] = v[t];
    } else if (Ns == 1 && R >= 4) {
      // first pass: a thread's R outputs are contiguous (j0 = j R): 16-byte stores, no bank conflicts from the R-strided scatter
      float4* d4 = reinterpret_cast<float4*>(dst + base + j0);
#pragma unroll
      for (int t = 0; t < R; t += 2) d4[t / 2] = make_float4(v[t].x, v[t].y, v[t + 1].x, v[t + 1].y);
    } else {
#pragma unroll
      for (int t = 0; t < R; ++t) dst[base + j0 + t * Ns] = v[t];
    }
  }
  __syncthreads();
}

// Shared-memory line stride (float2 units): L + 2 keeps 16-byte alignment and spreads the lines over the banks, so that the
// transposing accesses of the column kernels (16 lines x one element) are 2-way instead of 16-way conflicted.
__host__ __device__ __forceinline__ int line_stride(int L) { return L + 2; }

// Stage 0 of the r * 2^k decomposition, r = RT at compile time (1, 3, 5): r-point DFT over the stride-N2 samples + twiddle
// W_L^{n2 k1}, giving r sequences z_{k1}[n2] per line with X[k1 + r k2] = DFT_N2(z_k1)[k2].  The r-th roots of unity are read
// once per thread; index products are reduced with compile-time moduli.
template <int NCL, int RT>
__device__ __forceinline__ void stage0(const float2* buf, float2* tmp, const float2* tw, int L, int N2, float sgn) {
  const int LS = line_stride(L);
  float2 root[RT];
#pragma unroll
  for (int k = 0; k < RT; ++k) { root[k] = tw[k * N2]; root[k].y *= sgn; }
  const int lN2 = 31 - __clz(N2);
  for (int idx = threadIdx.x; idx < NCL * N2; idx += blockDim.x) {
    const int line = idx >> lN2, n2 = idx & (N2 - 1);
    const float2* src = buf + line * LS;
    float2* dst = tmp + line * LS;
    float2 x[RT];
#pragma unroll
    for (int n1 = 0; n1 < RT; ++n1) x[n1] = src[N2 * n1 + n2];
#pragma unroll
    for (int k1 = 0; k1 < RT; ++k1) {
      float2 acc = x[0];
#pragma unroll
      for (int n1 = 1; n1 < RT; ++n1) acc = cadd(acc, cmul(x[n1], root[(n1 * k1) % RT]));
      float2 t = tw[n2 * k1];
      t.y *= sgn;
      dst[k1 * N2 + n2] = k1 ? cmul(acc, t) : acc;
    }
  }
  __syncthreads();
}

// DFT of NCL lines of length L (line stride line_stride(L)) held in shared memory: input in buf, result in buf (natural
// order).  inverse: conjugate twiddles.  Uses tmp as the second ping-pong buffer.
template <int NCL, int RT>
__device__ __forceinline__ void fft_lines_r(float2* buf, float2* tmp, const float2* tw, FftGeom g, float sgn) {
  const int L = g.L, N2 = g.N2;
  stage0<NCL, RT>(buf, tmp, tw, L, N2, sgn);
  // Stockham passes on the NCL * r sequences of length N2 = 2^k: radix 8 first when k is odd, radix 4 for the rest;
  // ping-pong tmp -> buf -> tmp ...; the last pass writes the interleaved natural order.
  // k = 7: 8,4,4 (tmp->buf->tmp->buf); k = 6: 4,4,4; k = 5: 8,4 then a copy; k = 4: 4,4 + copy.
  const int nseq = NCL * RT, k = g.logN2;
  float2* a = tmp; float2* b = buf;
  int Ns = 1, left = k;
  auto swap = [&]() { float2* t_ = a; a = b; b = t_; };
  if (left & 1) {
    if (left == 3) { stockham_pass<8, true, RT>(a, b, tw, L, N2, Ns, nseq, sgn); left = 0; }
    else { stockham_pass<8, false, RT>(a, b, tw, L, N2, Ns, nseq, sgn); Ns *= 8; left -= 3; }
    swap();
  }
  while (left > 0) {
    if (left == 2) stockham_pass<4, true, RT>(a, b, tw, L, N2, Ns, nseq, sgn);
    else stockham_pass<4, false, RT>(a, b, tw, L, N2, Ns, nseq, sgn);
    Ns *= 4; left -= 2;
    swap();
  }
  if (a != buf) {   // after the swaps `a` holds the result
    const int LS = line_stride(L);
    for (int idx = threadIdx.x; idx < NCL * LS; idx += blockDim.x) buf[idx] = a[idx];
    __syncthreads();
  }
}

template <int NCL>
__device__ void fft_lines(float2* buf, float2* tmp, const float2* tw, FftGeom g, bool inverse) {
  const float sgn = inverse ? -1.0f : 1.0f;
  if (g.r == 3) fft_lines_r<NCL, 3>(buf, tmp, tw, g, sgn);
  else if (g.r == 5) fft_lines_r<NCL, 5>(buf, tmp, tw, g, sgn);
  else fft_lines_r<NCL, 1>(buf, tmp, tw, g, sgn);
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
  const float2* twg;   // prepared twiddle table W_L^j (or nullptr)
  FftGeom g;
};

// rows forward: image rows (zero-padded to L) -> rw
__global__ void __launch_bounds__(kFftThreads)
phase_rows_fwd_kernel(PhaseArgs a) {
  extern __shared__ float2 sm2[];
  const int L = a.g.L, D = a.D;
  const int LS = line_stride(L);
  float2* buf = sm2; float2* tmp = buf + NCR * LS; float2* tw = tmp + NCR * LS;
  const int mat = blockIdx.y, i0 = blockIdx.x * NCR;
  make_twiddles(tw, L, a.twg);
  for (int idx = threadIdx.x; idx < NCR * L; idx += blockDim.x) {
    const int line = idx / L, c = idx - line * L, i = i0 + line, j = c - a.pad;
    float v = 0.f;
    if (i < D && j >= 0 && j < D) {
      v = a.x[((int64_t)mat * D + i) * D + j];
      if (!a.no_clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
      if (a.xc_out) a.xc_out[((int64_t)mat * D + i) * D + j] = v;
    }
    buf[line * LS + c] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft_lines<NCR>(buf, tmp, tw, a.g, false);
  const float sc = rsqrtf((float)L);
  for (int idx = threadIdx.x; idx < NCR * L; idx += blockDim.x) {
    const int line = idx / L, v = idx - line * L, i = i0 + line;
    if (i < D) { const float2 t = buf[line * LS + v]; a.rw[((int64_t)mat * D + i) * L + v] = make_float2(t.x * sc, t.y * sc); }
  }
}

// columns. MODE 0: forward, write |X| to out (y layout). MODE 1: forward, residual, loss, inverse -> cw.
// MODE 2: pinv input: signed y as real input, inverse -> cw.
template <int MODE>
__global__ void __launch_bounds__(kFftThreads)
phase_cols_kernel(PhaseArgs a) {
  extern __shared__ float2 sm2[];
  const int L = a.g.L, D = a.D, h = L / 2;
  const int LS = line_stride(L);
  float2* buf = sm2; float2* tmp = buf + NCC * LS; float2* tw = tmp + NCC * LS;
  const int mat = blockIdx.y, v0 = blockIdx.x * NCC;
  make_twiddles(tw, L, a.twg);
  const float sc = rsqrtf((float)L);
  if (MODE != 2) {
    for (int idx = threadIdx.x; idx < NCC * L; idx += blockDim.x) {
      const int u = idx / NCC, c = idx - u * NCC, i = u - a.pad;
      buf[c * LS + u] = (i >= 0 && i < D) ? a.rw[((int64_t)mat * D + i) * L + v0 + c] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_lines<NCC>(buf, tmp, tw, a.g, false);
  }
  double acc = 0.0;
  for (int idx = threadIdx.x; idx < NCC * L; idx += blockDim.x) {
    const int u = idx / NCC, c = idx - u * NCC, v = v0 + c;
    const int64_t yoff = ((int64_t)mat * L + (u + h) % L) * L + (v + h) % L;
    if (MODE == 0) {
      const float2 X = buf[c * LS + u];
      a.out[yoff] = sqrtf(X.x * X.x + X.y * X.y) * sc;
    } else if (MODE == 1) {
      float2 X = buf[c * LS + u];
      X.x *= sc; X.y *= sc;
      const float mag = sqrtf(X.x * X.x + X.y * X.y);
      const float res = mag - a.y[yoff];
      acc += (double)res * (double)res;
      const float f = mag > 0.f ? res / mag : 0.f;      // d|z|/dz at 0 := 0 (torch's sgn convention)
      buf[c * LS + u] = make_float2(X.x * f, X.y * f);
    } else {
      const float sg = ((u + v) & 1) ? -1.0f : 1.0f;
      buf[c * LS + u] = make_float2(a.y[yoff] * sg, 0.f);
    }
  }
  if (MODE == 0) return;
  __syncthreads();
  fft_lines<NCC>(buf, tmp, tw, a.g, true);
  for (int idx = threadIdx.x; idx < NCC * D; idx += blockDim.x) {
    const int i = idx / NCC, c = idx - i * NCC;
    const float2 t = buf[c * LS + a.pad + i];
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
  const int LS = line_stride(L);
  float2* buf = sm2; float2* tmp = buf + NCR * LS; float2* tw = tmp + NCR * LS;
  const int mat = blockIdx.y, i0 = blockIdx.x * NCR;
  make_twiddles(tw, L, a.twg);
  for (int idx = threadIdx.x; idx < NCR * L; idx += blockDim.x) {
    const int line = idx / L, v = idx - line * L, i = i0 + line;
    buf[line * LS + v] = i < D ? a.cw[((int64_t)mat * D + i) * L + v] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  fft_lines<NCR>(buf, tmp, tw, a.g, true);
  const float sc = rsqrtf((float)L);
  for (int idx = threadIdx.x; idx < NCR * D; idx += blockDim.x) {
    const int line = idx / D, j = idx - line * D, i = i0 + line;
    if (i >= D) continue;
    const float2 t = buf[line * LS + a.pad + j];
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

static size_t fft_smem(int L, int nc) { return ((size_t)2 * nc * line_stride(L) + L) * sizeof(float2); }

struct PhaseWs { double* partials; unsigned int* tickets; float2* rw; float2* cw; };
static size_t phase_red_bytes(const nhmc_op* op, int n_chains) {
  const int L = op->img_dim + 2 * op->pad;
  size_t b = (size_t)n_chains * op->channels * (L / NCC + 1) * sizeof(double) + (size_t)n_chains * sizeof(unsigned int);
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
  w.tickets = reinterpret_cast<unsigned int*>(w.partials + (size_t)n_chains * op->channels * (L / NCC + 1));
  w.rw = reinterpret_cast<float2*>((char*)extra + phase_red_bytes(op, n_chains));
  w.cw = w.rw + (size_t)n_chains * op->channels * op->img_dim * L;
  return w;
}

static int phase_setup(const nhmc_op* op, PhaseArgs& a, size_t& smem_r, size_t& smem_c) {
  const int L = op->img_dim + 2 * op->pad;
  if (!fft_geom(L, a.g)) {
    set_error("phase: padded size %d is not r*2^k with odd r <= 15", L);
    return NHMC_ERR_UNSUPPORTED;
  }
  a.D = op->img_dim; a.pad = op->pad; a.channels = op->channels; a.no_clamp = op->no_clamp;
  a.twg = reinterpret_cast<const float2*>(op->tc_tables);
  smem_r = fft_smem(L, NCR); smem_c = fft_smem(L, NCC);
  if (smem_c > 200 * 1024) { set_error("phase: padded size %d too large for shared memory", L); return NHMC_ERR_UNSUPPORTED; }
  if (smem_r > 48 * 1024) {
    cudaFuncSetAttribute(phase_rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
    cudaFuncSetAttribute(phase_rows_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
    cudaFuncSetAttribute(phase_rows_inv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
  }
  if (smem_c > 48 * 1024) {
    cudaFuncSetAttribute(phase_cols_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
    cudaFuncSetAttribute(phase_cols_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
    cudaFuncSetAttribute(phase_cols_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
  }
  return NHMC_OK;
}

int phase_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* extra, cudaStream_t st) {
  PhaseArgs a{}; size_t smem_r, smem_c;
  if (int e = phase_setup(op, a, smem_r, smem_c)) return e;
  const PhaseWs w = phase_ws(op, n_chains, extra);
  const int mats = n_chains * op->channels, L = a.g.L;
  a.x = x; a.out = y; a.rw = w.rw; a.no_clamp = 1;   // H() itself never clamps
  phase_rows_fwd_kernel<<<dim3((a.D + NCR - 1) / NCR, mats), kFftThreads, smem_r, st>>>(a);
  phase_cols_kernel<0><<<dim3(L / NCC, mats), kFftThreads, smem_c, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

int phase_pinv(const nhmc_op* op, const float* y, float* out, int n_chains, void* extra, cudaStream_t st) {
  PhaseArgs a{}; size_t smem_r, smem_c;
  if (int e = phase_setup(op, a, smem_r, smem_c)) return e;
  const PhaseWs w = phase_ws(op, n_chains, extra);
  const int mats = n_chains * op->channels, L = a.g.L;
  a.y = y; a.out = out; a.cw = w.cw;
  phase_cols_kernel<2><<<dim3(L / NCC, mats), kFftThreads, smem_c, st>>>(a);
  phase_rows_inv_kernel<1><<<dim3((a.D + NCR - 1) / NCR, mats), kFftThreads, smem_r, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

int phase_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss,
                    int n_chains, void* workspace, cudaStream_t st) {
  PhaseArgs a{}; size_t smem_r, smem_c;
  if (int e = phase_setup(op, a, smem_r, smem_c)) return e;
  const PhaseWs w = phase_ws(op, n_chains, (char*)workspace + red_ws_aligned(n_chains));
  const int mats = n_chains * op->channels, L = a.g.L;
  a.x = x_hat; a.y = y; a.out = g; a.xc_out = xc_out; a.rw = w.rw; a.cw = w.cw;
  a.loss = loss; a.partials = w.partials; a.tickets = w.tickets;
  phase_rows_fwd_kernel<<<dim3((a.D + NCR - 1) / NCR, mats), kFftThreads, smem_r, st>>>(a);
  phase_cols_kernel<1><<<dim3(L / NCC, mats), kFftThreads, smem_c, st>>>(a);
  phase_rows_inv_kernel<0><<<dim3((a.D + NCR - 1) / NCR, mats), kFftThreads, smem_r, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

__global__ void phase_twiddle_kernel(float2* table, int L) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
    float sn, cs;
    sincospif(2.0f * (float)j / (float)L, &sn, &cs);
    table[j] = make_float2(cs, -sn);
  }
}
size_t phase_tables_bytes(const nhmc_op* op) { return (size_t)(op->img_dim + 2 * op->pad) * sizeof(float2); }
int phase_prepare_tables(const nhmc_op* op, void* tables, cudaStream_t st) {
  const int L = op->img_dim + 2 * op->pad;
  phase_twiddle_kernel<<<(L + 255) / 256, 256, 0, st>>>(reinterpret_cast<float2*>(tables), L);
  return NHMC_LAUNCH_CHECK();
}

}  // namespace nhmc
