// K9: GroupNorm(32) fused with its follow-up (SiLU / FiLM + SiLU), forward and input-backward.
// Replaces GroupNorm32 + SiLU (+ `h*(1+scale)+shift`) of the ADM U-Net
// (guided_diffusion/nn.py:17-19; unet_ffhq.py:253-257,277-280,319-323,356,683-685): fp32 statistics,
// activations in the network dtype, one statistics pass + one apply pass instead of
// cast -> native_group_norm -> cast -> mul/add -> silu (5+ full passes in the reference).
//
// HBM-bound.  Layouts: channels-last [B, HW, C] (what the bf16 convolutions want on sm_100) and
// NCHW [B, C, HW].  Channels-last mapping: a thread owns VEC consecutive channels (16 B) of a pixel and
// walks down the pixels of its CTA's chunk, so every access is a full 16-byte vector and the
// per-channel affine (rstd*gamma*(1+scale), ...) is computed once per thread.  Statistics: fp32
// per-thread partials -> shared-memory ordered reduce -> fp64 per-(sample, group) finish by the last CTA
// (deterministic).  Weights are frozen on this path (sampling), so backward produces dx only.
#include <stdlib.h>

#include "common.cuh"

namespace nhmc {

template <typename T> struct VecOf { static constexpr int N = 16 / sizeof(T); };

template <typename T, int N> struct Pack { float v[N]; };

template <typename T> __device__ __forceinline__ void load_vec(const T* p, float (&v)[16 / sizeof(T)]);
template <> __device__ __forceinline__ void load_vec<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <> __device__ __forceinline__ void load_vec<__half>(const __half* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void store_vec(T* p, const float (&v)[16 / sizeof(T)]);
template <> __device__ __forceinline__ void store_vec<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 t;
  uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = t;
}
template <> __device__ __forceinline__ void store_vec<__half>(__half* p, const float (&v)[8]) {
  uint4 t;
  uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = t;
}

// raw 16-byte streaming load (read-only path) and in-register unpack: keeps U loads in flight per thread
__device__ __forceinline__ uint4 ld_raw(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <typename T> __device__ __forceinline__ void unpack(const uint4& t, float (&v)[16 / sizeof(T)]);
template <> __device__ __forceinline__ void unpack<float>(const uint4& t, float (&v)[4]) {
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void unpack<__half>(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__device__ __forceinline__ float sigmoidf_fast(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }

struct GnArgs {
  const void* x; const void* dy; void* out;
  const void* add;   // backward: optional tensor added to dx in the same pass (skip-branch gradient), or null
  const void* add2;  // backward, single-source statistics + apply path only: a second addend (the gradient that reaches the
                     // block input through the U-Net's long skip connection), or null
  // virtual concatenation (unet_ffhq.py:730): when x1 != null the normalised tensor is cat([x, x1], channel) with
  // c_split channels in x and C - c_split in x1, each a dense channels-last block of its own; the backward then
  // writes dx / reads add per block (out, out1 / add, add1; add and add1 are independently nullable)
  const void* x1; void* out1; const void* add1; int c_split;
  const float* gamma; const float* beta; const float* scale; const float* shift;  // scale/shift: [B,C] fp32 or null
  const float* in_bias;  // fp32 or null: the normalised tensor is x + in_bias[b*in_bias_stride + c] (conv bias, and for
                         // the additive-embedding ResBlock variant the per-sample embedding, folded in)
  int64_t in_bias_stride;  // 0: one [C] vector shared by all samples
  float* stats;      // [B][G][2]: fwd (mean, rstd)
  float* bstats;     // [B][G][2]: bwd (sum du / n, sum du*u / n)
  double* partials;  // [B][kGnMaxBlocks][G][2]
  unsigned int* tickets;  // [B]
  int C, G; int64_t HW; float eps;
  int64_t film_stride;  // elements between two samples' rows of scale / shift (>= C)
  int b0;            // first sample of this launch (two-kernel path launched per L2-sized group of samples)
  int reps;          // statistics + apply kernels: replicas of the fp64 accumulators (CTA blk adds to replica blk % reps of
                     // its sample, slot = replica * gridDim.y + sample; 0 or 1 = none): spreads the atomics of the many CTAs
                     // of one sample (B = 1: 444) over more addresses
  int rows_per_cta_iter;  // pixel rows covered by one CTA pass (channels-last)
  int64_t rows_per_cta;   // pixel rows owned by a CTA (channels-last)
};

// per-element pieces --------------------------------------------------------------------------
// sigmoid: one MUFU (tanh.approx, rel. error 2^-11 < half a bf16/fp16 ulp) for 16-bit activations -- the SFU
// pipe, not HBM, bounded the apply kernels with the 2-MUFU ex2+rcp form; fp32 activations keep the exact form.
template <typename T> __device__ __forceinline__ float sigmoid_t(float z) {
  if (sizeof(T) == 2) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(0.5f, t, 0.5f);
  }
  return sigmoidf_fast(z);
}
template <int MODE, typename T> __device__ __forceinline__ float gn_act(float z) {
  return MODE == NHMC_GN_PLAIN ? z : z * sigmoid_t<T>(z);
}
// Channels-last kernels fold the 1/2 of sigmoid(z) = (1 + tanh(z/2))/2 into the per-channel affine for 16-bit
// activations: the affine yields h = z/2 directly, silu(z) = h + h*tanh(h) is one MUFU + one FMA per element and
// silu'(z) = 1/2 + (tanh(h) + h*(1 - tanh(h)^2))/2 one MUFU + three FMAs (these kernels run within 2x of the issue
// limit, so instruction count is bandwidth).  fp32 activations keep the exact sigmoid.
template <int MODE, typename T> struct AffineHalf { static constexpr bool value = MODE != NHMC_GN_PLAIN && sizeof(T) == 2; };
template <int MODE, typename T> __device__ __forceinline__ float gn_act_a(float w) {   // w = z/2 if AffineHalf else z
  if (AffineHalf<MODE, T>::value) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(w));
    return fmaf(w, t, w);
  }
  return MODE == NHMC_GN_PLAIN ? w : w * sigmoidf_fast(w);
}
template <int MODE, typename T> __device__ __forceinline__ float gn_dact_a(float w) {
  if (MODE == NHMC_GN_PLAIN) return 1.0f;
  if (AffineHalf<MODE, T>::value) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(w));
    const float u = fmaf(-t, t, 1.0f);
    return fmaf(0.5f, fmaf(w, u, t), 0.5f);
  }
  const float s = sigmoidf_fast(w);
  return s * (1.0f + w * (1.0f - s));
}
template <int MODE, typename T> __device__ __forceinline__ float gn_dact(float z) {  // d act / dz
  if (MODE == NHMC_GN_PLAIN) return 1.0f;
  const float s = sigmoid_t<T>(z);
  return s * (1.0f + z * (1.0f - s));
}

// Tuning constants of the streaming (statistics / apply) kernels: 16-byte loads in flight per thread (one source: U1, two
// sources: U2 per source) and the resident CTAs per SM the register allocation is held to.  Chosen by measurement
// (profiles/r2_gn_stream_variants.md); overridable at compile time for that sweep.
#ifndef NHMC_GN_U1
#define NHMC_GN_U1 8
#endif
#ifndef NHMC_GN_U2
#define NHMC_GN_U2 4
#endif
#ifndef NHMC_GN_MINB_STATS
#define NHMC_GN_MINB_STATS 3
#endif
#ifndef NHMC_GN_MINB_APPLY
#define NHMC_GN_MINB_APPLY 3
#endif

// ============================== channels-last ==============================================
// A CTA owns `rows_per_cta` consecutive pixels of one sample = ONE contiguous byte range of the NHWC
// tensor.  Thread t always lands on the same VEC channels (a pass of blockDim * 16 B is a whole number of
// pixel rows), so the per-channel affine lives in registers.

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// This thread's 16-byte vector of a channels-last [rows, C] tensor: p = address in row 0, stride = bytes per row.
// The tensor is one dense block, or -- virtual concatenation -- two dense blocks split at channel c_split.
struct RowPtr { char* p; int64_t stride; };
template <typename T>
__device__ __forceinline__ RowPtr row_ptr(const void* block, int width, int c_in_block, int64_t first_row) {
  RowPtr r;
  r.stride = (int64_t)width * sizeof(T);
  r.p = const_cast<char*>(reinterpret_cast<const char*>(block)) + first_row * r.stride + (int64_t)c_in_block * sizeof(T);
  return r;
}
// Which block of a (possibly split) tensor holds this thread's channels c0 .. c0+V-1.
struct SegSel { bool second; int width, cc; };
__device__ __forceinline__ SegSel seg_select(const void* x1, int c_split, int C, int c0) {
  SegSel s;
  s.second = x1 != nullptr && c0 >= c_split;
  s.width = x1 != nullptr ? (s.second ? C - c_split : c_split) : C;
  s.cc = s.second ? c0 - c_split : c0;
  return s;
}

// Streams rows [r_begin, r_end) of x (and dy for the backward kernels): this thread visits rows r_begin + prow,
// + rows, ... and calls body(x_vec, dy_vec, ep) where ep[i] is the address of this thread's vector of the i-th
// extra tensor (outputs, addends) in the same row.  Plain 16-byte read-only loads, U (x NSRC) of them in flight
// per thread before the first use; every address advances by pointer bumps (no per-load multiplies: these
// kernels sit within 2x of the issue limit).  Measured on B200 this simple form streams at 6-7 TB/s, while the
// earlier mbarrier/TMA shared-memory ring topped out at ~4.5 TB/s (59 % issue slots busy on ring bookkeeping,
// 57 % DRAM throughput).  TMA stays where it buys something: the resident kernel below, which needs the slab in
// shared memory anyway.
template <int NSRC, int NX, bool REV, typename Body>
__device__ __forceinline__ void stream_rows(const RowPtr x, const RowPtr d, const RowPtr (&ex)[NX], int64_t r_begin,
                                            int64_t r_end, int prow, int rows, Body body) {
  constexpr int U = NSRC == 2 ? NHMC_GN_U2 : NHMC_GN_U1;
  const int64_t r0 = r_begin + prow;
  const int n = r_end > r0 ? (int)((r_end - r0 + rows - 1) / rows) : 0;      // rows this thread visits
  // REV: last row first (so that a re-read starts on what the previous kernel touched last).  Measured on B200: no gain
  // for the apply kernels (the L2 does not behave like an LRU over a 268 MB stream), so both kernels walk forwards.
  const int64_t first = REV ? r0 + (int64_t)(n > 0 ? n - 1 : 0) * rows : r0;
  const int64_t dir = REV ? -1 : 1;
  const char* px = x.p + first * x.stride;
  const char* pd = NSRC == 2 ? d.p + first * d.stride : nullptr;
  const int64_t sx = dir * rows * x.stride, sd = dir * rows * d.stride;
  char* pe[NX];
  int64_t se[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) { pe[i] = ex[i].p + first * ex[i].stride; se[i] = dir * rows * ex[i].stride; }
  for (int k = 0; k < n; k += U) {
    uint4 vx[U], vd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      vx[u] = make_uint4(0, 0, 0, 0); vd[u] = make_uint4(0, 0, 0, 0);
      if (k + u < n) {
        vx[u] = ld_raw(px);
        if (NSRC == 2) vd[u] = ld_raw(pd);
      }
      px += sx;
      if (NSRC == 2) pd += sd;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (k + u < n) body(vx[u], vd[u], pe);
#pragma unroll
      for (int i = 0; i < NX; ++i) pe[i] += se[i];
    }
  }
}

// PHASE 0: forward statistics (sum x, sum x^2)            -> fp64 accumulators [B][64][2]
// PHASE 1: backward statistics (sum du, sum du*u)          -> same accumulators
template <typename T, int MODE, int PHASE>
__global__ void __launch_bounds__(256, NHMC_GN_MINB_STATS)
gn_nhwc_stats_kernel(GnArgs a) {
  pdl_wait(); pdl_trigger();
  constexpr int V = VecOf<T>::N;
  constexpr float kAff = AffineHalf<MODE, T>::value ? 0.5f : 1.0f;
  extern __shared__ __align__(128) unsigned char dyn[];
  const int b = blockIdx.y + a.b0, blk = blockIdx.x;
  const float* in_bias = a.in_bias ? a.in_bias + (int64_t)b * a.in_bias_stride : nullptr;
  const int cvec = a.C / V;
  const int cv = threadIdx.x % cvec, prow = threadIdx.x / cvec, rows = blockDim.x / cvec;
  const int c0 = cv * V;
  const int Cg = a.C / a.G;
  float s0[V], s1[V], ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s0[i] = 0.f; s1[i] = 0.f; ca[i] = 0.f; cb[i] = 0.f; }
  if (PHASE == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (a.in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
      ca[i] = rstd * gam * sc * kAff;                // z = ca*x + cb (z/2 for 16-bit SiLU, see AffineHalf)
      cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    }
  }
  const int64_t r_begin = (int64_t)blk * a.rows_per_cta;
  int64_t r_end = r_begin + a.rows_per_cta; if (r_end > a.HW) r_end = a.HW;
  const SegSel sel = seg_select(a.x1, a.c_split, a.C, c0);
  const RowPtr xr = row_ptr<T>(sel.second ? a.x1 : a.x, sel.width, sel.cc, (int64_t)b * a.HW);
  const RowPtr dr = row_ptr<T>(a.dy, a.C, c0, (int64_t)b * a.HW);
  // PHASE 1 accumulates t0 = sum dy*act'(z), t1 = sum dy*act'(z)*x per channel; the per-channel
  // affine (gamma(1+s), mean, rstd) is applied once after the loop.
  const RowPtr none[1] = {{nullptr, 0}};
  stream_rows<PHASE == 1 ? 2 : 1, 1, false>(xr, dr, none, r_begin, r_end, prow, rows,
      [&](const uint4& rx, const uint4& rd, char* const (&)[1]) {
        float xv[V];
        unpack<T>(rx, xv);
        if (PHASE == 0) {
#pragma unroll
          for (int i = 0; i < V; ++i) { s0[i] += xv[i]; s1[i] = fmaf(xv[i], xv[i], s1[i]); }
        } else {
          float dv[V];
          unpack<T>(rd, dv);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float d = dv[i] * gn_dact_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
            s0[i] += d; s1[i] = fmaf(d, xv[i], s1[i]);
          }
        }
      });
  if (PHASE == 0 && in_bias) {
    // statistics of x + bias from the raw sums: this thread saw `cnt` pixels of each of its channels
    const int64_t span = r_end - r_begin - prow;
    const int64_t cnt = span > 0 ? (span + rows - 1) / rows : 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float bia = in_bias[c0 + i];
      s1[i] = s1[i] + 2.0f * bia * s0[i] + (float)cnt * bia * bia;
      s0[i] = s0[i] + (float)cnt * bia;
    }
  }
  if (PHASE == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (a.in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float cgm = a.gamma[c] * sc;             // du = d * gamma * (1+scale); u = (x-mean)*rstd
      const float t0 = s0[i], t1 = s1[i];
      s0[i] = cgm * t0;
      s1[i] = cgm * rstd * (t1 - mean * t0);
    }
  }
  // ordered reduce over the CTA's pixel rows, then over the channels of each group
  float* sm0 = reinterpret_cast<float*>(dyn); float* sm1 = sm0 + (size_t)rows * a.C;
#pragma unroll
  for (int i = 0; i < V; ++i) { sm0[(size_t)prow * a.C + c0 + i] = s0[i]; sm1[(size_t)prow * a.C + c0 + i] = s1[i]; }
  __syncthreads();
  // one fp64 atomic pair per (CTA, group): no serial tail, no partial-sum array.  The accumulators are
  // zero on entry (scratch is zeroed at allocation; the apply kernel re-zeroes them after its last read).
  for (int g = threadIdx.x; g < a.G; g += blockDim.x) {
    double t0 = 0.0, t1 = 0.0;
    for (int c = g * Cg; c < (g + 1) * Cg; ++c)
      for (int r = 0; r < rows; ++r) { t0 += (double)sm0[(size_t)r * a.C + c]; t1 += (double)sm1[(size_t)r * a.C + c]; }
    const int64_t slot = a.reps > 1 ? (int64_t)(blk % a.reps) * gridDim.y + b : b;
    atomicAdd(a.partials + (slot * 64 + g) * 2, t0);
    atomicAdd(a.partials + (slot * 64 + g) * 2 + 1, t1);
  }
}

// PHASE 0: y = act(ca*x + cb) ; PHASE 1: dx = rstd*(du - S1/n - u*S2/n)
template <typename T, int MODE, int PHASE, bool ADD2 = false>
__global__ void __launch_bounds__(256, NHMC_GN_MINB_APPLY)
gn_nhwc_apply_kernel(GnArgs a) {
  pdl_wait(); pdl_trigger();
  constexpr int V = VecOf<T>::N;
  constexpr float kAff = AffineHalf<MODE, T>::value ? 0.5f : 1.0f;
  const int b = blockIdx.y + a.b0, blk = blockIdx.x;
  const float* in_bias = a.in_bias ? a.in_bias + (int64_t)b * a.in_bias_stride : nullptr;
  const int cvec = a.C / V;
  const int cv = threadIdx.x % cvec, prow = threadIdx.x / cvec, rows = blockDim.x / cvec;
  const int c0 = cv * V;
  const int Cg = a.C / a.G;
  float ca[V], cb[V], k1[V], k2[V], k3[V];
  const double n = (double)Cg * (double)a.HW;
  // group statistics once per CTA (G <= 64 threads touch fp64), everybody else reads floats from shared memory
  __shared__ float g_mean[64], g_rstd[64], g_m1[64], g_m2[64];
  if (threadIdx.x < a.G) {
    const int g = threadIdx.x;
    const double inv_n = 1.0 / n;
    double a0 = __ldcg(a.partials + ((int64_t)b * 64 + g) * 2), a1 = __ldcg(a.partials + ((int64_t)b * 64 + g) * 2 + 1);
    for (int rep = 1; rep < a.reps; ++rep) {
      const int64_t slot = (int64_t)rep * gridDim.y + b;
      a0 += __ldcg(a.partials + (slot * 64 + g) * 2); a1 += __ldcg(a.partials + (slot * 64 + g) * 2 + 1);
    }
    a0 *= inv_n; a1 *= inv_n;
    if (PHASE == 0) {
      double var = a1 - a0 * a0; if (var < 0.0) var = 0.0;
      g_mean[g] = (float)a0; g_rstd[g] = (float)(1.0 / sqrt(var + (double)a.eps));
    } else {
      g_mean[g] = a.stats[((int64_t)b * a.G + g) * 2]; g_rstd[g] = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      g_m1[g] = (float)a0; g_m2[g] = (float)a1;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = c0 + i, g = c / Cg;
    const float mean = g_mean[g] - (a.in_bias ? in_bias[c] : 0.0f), rstd = g_rstd[g];
    const float gam = a.gamma[c], bet = a.beta[c];
    const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
    const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
    ca[i] = rstd * gam * sc * kAff;
    cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    k1[i] = k2[i] = k3[i] = 0.f;
    if (PHASE == 1) {
      // dx = rstd*(du - m1 - u*m2), du = dy*act'(z)*gamma*(1+s), u = (x-mean)*rstd
      //    = k1*dy*act'(z) + k2*x + k3
      const float m1 = g_m1[g], m2 = g_m2[g];
      k1[i] = rstd * gam * sc;
      k2[i] = -rstd * rstd * m2;
      k3[i] = -rstd * m1 + rstd * rstd * m2 * mean;
    }
  }
  const int64_t r_begin = (int64_t)blk * a.rows_per_cta;
  int64_t r_end = r_begin + a.rows_per_cta; if (r_end > a.HW) r_end = a.HW;
  const int64_t first = (int64_t)b * a.HW;
  const SegSel sel = seg_select(a.x1, a.c_split, a.C, c0);
  const RowPtr xr = row_ptr<T>(sel.second ? a.x1 : a.x, sel.width, sel.cc, first);
  const RowPtr dr = row_ptr<T>(a.dy, a.C, c0, first);
  // forward: y is one dense tensor; backward: dx (and add) follow the layout of x
  const RowPtr orow = PHASE == 0 ? row_ptr<T>(a.out, a.C, c0, first)
                                 : row_ptr<T>(sel.second ? a.out1 : a.out, sel.width, sel.cc, first);
  const void* addp = sel.second ? a.add1 : a.add;
  const bool has_add = PHASE == 1 && addp != nullptr;
  constexpr int NX = ADD2 ? 3 : 2;
  const RowPtr extra[3] = {orow, row_ptr<T>(addp, sel.width, sel.cc, first), row_ptr<T>(ADD2 ? a.add2 : nullptr, a.C, c0, first)};
  const RowPtr (&extra_nx)[NX] = reinterpret_cast<const RowPtr (&)[NX]>(extra);
  stream_rows<PHASE == 1 ? 2 : 1, NX, false>(xr, dr, extra_nx, r_begin, r_end, prow, rows,
      [&](const uint4& rx, const uint4& rd, char* const (&ep)[NX]) {
        float xv[V], ov[V];
        unpack<T>(rx, xv);
        if (PHASE == 0) {
#pragma unroll
          for (int i = 0; i < V; ++i) ov[i] = gn_act_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
        } else {
          float dv[V];
          unpack<T>(rd, dv);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float d = dv[i] * gn_dact_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
            ov[i] = fmaf(k1[i], d, fmaf(k2[i], xv[i], k3[i]));
          }
          if (ADD2) {            // both addends present (host contract): issue the two loads together
            float ad[V], ae[V];
            const uint4 r1 = ld_raw(ep[1]);
            const uint4 r2 = ld_raw(ep[ADD2 ? 2 : 1]);
            unpack<T>(r1, ad); unpack<T>(r2, ae);
#pragma unroll
            for (int i = 0; i < V; ++i) ov[i] += ad[i] + ae[i];
          } else if (has_add) {
            float ad[V];
            unpack<T>(ld_raw(ep[1]), ad);
#pragma unroll
            for (int i = 0; i < V; ++i) ov[i] += ad[i];
          }
        }
        store_vec<T>(reinterpret_cast<T*>(ep[0]), ov);
      });
  // Every CTA read the accumulators before streaming; the last CTA to finish publishes (mean, rstd) for the
  // backward pass and re-zeroes the accumulators for the next call on this stream (no wait on the hot path).
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicInc(a.tickets + b, (unsigned int)(gridDim.x - 1)) == (unsigned int)(gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    for (int g = threadIdx.x; g < a.G; g += blockDim.x) {
      double* acc = a.partials + ((int64_t)b * 64 + g) * 2;
      if (PHASE == 0) {
        a.stats[((int64_t)b * a.G + g) * 2] = g_mean[g];
        a.stats[((int64_t)b * a.G + g) * 2 + 1] = g_rstd[g];
      }
      acc[0] = 0.0; acc[1] = 0.0;
      for (int rep = 1; rep < a.reps; ++rep) {
        double* racc = a.partials + (((int64_t)rep * gridDim.y + b) * 64 + g) * 2;
        racc[0] = 0.0; racc[1] = 0.0;
      }
    }
  }
}

// ============================== channels-last, shared-memory resident ========================
// ONE kernel, ONE read and ONE write of the tensor (the algorithmic minimum): a CTA pulls its contiguous
// slab of pixels into shared memory with TMA bulk copies (the whole slab is in flight at once), reduces
// its per-group partial sums while the chunks land, publishes them with fp64 atomics, waits until every
// CTA of the same SAMPLE has done so (per-sample arrival counter), then normalises straight out of shared
// memory.  Nothing is read twice, no second launch.
//
// Forward progress without a cooperative launch: a CTA takes its logical index from a global ticket when it
// STARTS, so logical indices follow start order; the CTAs of the lowest unfinished sample are therefore
// either running or next in line, and the host sizes slabs so that one sample's CTAs always fit on the chip
// together (occupancy API).  The spin is bounded anyway: on timeout an error flag is raised, never a hang.
constexpr int kResChunks = 16;   // mbarriers per CTA (chunks of the slab)

struct GnResArgs {
  GnArgs a;
  int nb;                   // CTAs per sample
  int chunk_passes;         // passes (blockDim * 16 B) per chunk
  unsigned int total_ctas;
  unsigned int* ticket;     // [1]   start-order ticket, wraps to 0 by itself
  unsigned int* arrive;     // [B]   CTAs of the sample whose partial sums are published
  unsigned int* done;       // [B]   CTAs of the sample that have read the totals (wraps to 0 by itself)
  unsigned int* err;        // [1]   set to 1 if a wait timed out
  int reps;                 // replicas of the fp64 accumulators per (sample, group): CTA blk adds to replica blk % reps, so that
                            // hundreds of CTAs of one sample (B = 1: 592) do not serialise on 64 addresses; readers sum them
  int batch;                // samples of this launch (replica r of sample b lives in accumulator slot r * batch + b)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Group totals of per-channel partial sums (s0, s1) of this CTA -> g_acc[g][0..1] (fp64), then -- when the
// sample spans several CTAs -- through the global accumulators and the per-sample barrier.  Ends with the
// sample-wide totals in g_acc, visible to every thread.
// cluster helpers: hardware barrier across the CTAs of a thread-block cluster and reads of a peer CTA's shared memory
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ double ld_dsmem_f64(const double* local_smem, unsigned rank) {
  unsigned addr = smem_u32(local_smem), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}

template <int V, bool CLUSTER>
__device__ __forceinline__ void group_totals(const float (&s0)[V], const float (&s1)[V], const GnResArgs& r, int b,
                                             int c0, int cvec, int rows, float4* part, double (*g_acc)[2],
                                             double (*g_part)[2]) {
  const GnArgs& a = r.a;
  const int Cg = a.C / a.G;
  const int g0 = c0 / Cg;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);      // (slot0.s0, slot0.s1, slot1.s0, slot1.s1)
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if ((c0 + i) / Cg == g0) { q.x += s0[i]; q.y += s1[i]; } else { q.z += s0[i]; q.w += s1[i]; }
  }
  part[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < a.G) {
    const int g = threadIdx.x;
    const int cv_lo = (g * Cg) / V, cv_hi = ((g + 1) * Cg - 1) / V;
    double t0 = 0.0, t1 = 0.0;
    for (int cv = cv_lo; cv <= cv_hi; ++cv) {
      const int j = g - (cv * V) / Cg;             // 0 or 1 (host guarantees a thread spans <= 2 groups)
      for (int pr = 0; pr < rows; ++pr) {
        const float4 t = part[pr * cvec + cv];
        t0 += (double)(j ? t.z : t.x); t1 += (double)(j ? t.w : t.y);
      }
    }
    if (CLUSTER) {
      g_part[g][0] = t0; g_part[g][1] = t1;        // read by every CTA of the cluster through DSMEM
    } else if (r.nb > 1) {
      const int64_t slot = (int64_t)(blockIdx.x % (unsigned)r.reps) * r.batch + b;
      atomicAdd(a.partials + (slot * 64 + g) * 2, t0);
      atomicAdd(a.partials + (slot * 64 + g) * 2 + 1, t1);
      __threadfence();
    } else {
      g_acc[g][0] = t0; g_acc[g][1] = t1;
    }
  }
  __syncthreads();
  if (CLUSTER) {
    // one sample = one thread-block cluster: hardware barrier, then every CTA sums the nb partials of each group
    // in rank order out of its peers' shared memory -- no global atomics, no spinning, deterministic
    cluster_arrive();
    cluster_wait();
    if (threadIdx.x < a.G) {
      const int g = threadIdx.x;
      double t0 = 0.0, t1 = 0.0;
      for (int k = 0; k < r.nb; ++k) { t0 += ld_dsmem_f64(&g_part[g][0], (unsigned)k); t1 += ld_dsmem_f64(&g_part[g][1], (unsigned)k); }
      g_acc[g][0] = t0; g_acc[g][1] = t1;
    }
    cluster_arrive();        // peers may exit (and release their shared memory) only after everybody has read: waited at kernel end
    __syncthreads();
    return;
  }
  if (r.nb > 1) {
    if (threadIdx.x == 0) {
      atomicAdd(r.arrive + b, 1u);
      const long long t_start = clock64();
      while (ld_acquire_u32(r.arrive + b) < (unsigned int)r.nb) {
        __nanosleep(64);
        if (clock64() - t_start > (1ll << 31)) { atomicExch(r.err, 1u); break; }   // ~1 s: never hang the device
      }
    }
    __syncthreads();
    if (threadIdx.x < a.G) {
      const int g = threadIdx.x;
      double t0 = 0.0, t1 = 0.0;
      for (int rep = 0; rep < r.reps; ++rep) {       // fixed order: deterministic given the replica contents
        const int64_t slot = (int64_t)rep * r.batch + b;
        t0 += __ldcg(a.partials + (slot * 64 + g) * 2); t1 += __ldcg(a.partials + (slot * 64 + g) * 2 + 1);
      }
      g_acc[g][0] = t0; g_acc[g][1] = t1;
    }
    __syncthreads();
    // the last CTA of the sample to have read the totals re-zeroes them for the next call on this stream
    if (threadIdx.x == 0) {
      const bool last = atomicInc(r.done + b, (unsigned int)(r.nb - 1)) == (unsigned int)(r.nb - 1);
      if (last) {
        for (int rep = 0; rep < r.reps; ++rep) {
          const int64_t slot = (int64_t)rep * r.batch + b;
          for (int g = 0; g < a.G; ++g) { a.partials[(slot * 64 + g) * 2] = 0.0; a.partials[(slot * 64 + g) * 2 + 1] = 0.0; }
        }
        r.arrive[b] = 0u;
      }
    }
  }
}

// PHASE 0: y = act(gn(x + in_bias));  PHASE 1: dx = gn_bwd(dy) (+ add)
template <typename T, int MODE, int PHASE, bool CLUSTER>
__global__ void __launch_bounds__(256, PHASE == 1 ? 3 : 4)
gn_nhwc_resident_kernel(const GnResArgs r) {
  constexpr int V = VecOf<T>::N;
  constexpr float kAff = AffineHalf<MODE, T>::value ? 0.5f : 1.0f;
  constexpr int NSRC = PHASE == 1 ? 2 : 1;
  const GnArgs& a = r.a;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bars[kResChunks];
  __shared__ float4 part[256];
  __shared__ double g_acc[32][2];      // the resident path takes G <= 32 (static shared memory is part of the budget)
  __shared__ float g_c0[32], g_c1[32];
  __shared__ double g_part[CLUSTER ? 32 : 1][2];
  __shared__ unsigned int s_lid;
  if (threadIdx.x == 0) {
    // cluster launch: sample = cluster index, CTA of the sample = rank in the cluster (grid is [B * nb], cluster [nb])
    s_lid = CLUSTER ? blockIdx.x : atomicInc(r.ticket, r.total_ctas - 1u);
    for (int k = 0; k < kResChunks; ++k) mbar_init(bars + k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int b = (int)(s_lid / (unsigned int)r.nb), blk = (int)(s_lid % (unsigned int)r.nb);
  const float* in_bias = a.in_bias ? a.in_bias + (int64_t)b * a.in_bias_stride : nullptr;
  const int cvec = a.C / V;
  const int cv = threadIdx.x % cvec, rows = blockDim.x / cvec;
  const int c0 = cv * V;
  const int Cg = a.C / a.G;
  const int64_t r_begin = (int64_t)blk * a.rows_per_cta;
  int64_t r_end = r_begin + a.rows_per_cta; if (r_end > a.HW) r_end = a.HW;
  const int64_t row_bytes = (int64_t)a.C * sizeof(T);
  const int64_t base = ((int64_t)b * a.HW + r_begin) * row_bytes;
  const unsigned total = r_end > r_begin ? (unsigned)((r_end - r_begin) * row_bytes) : 0u;
  const unsigned pass_bytes = blockDim.x * 16u;
  const unsigned chunk_bytes = (unsigned)r.chunk_passes * pass_bytes;
  const unsigned n_pass = (total + pass_bytes - 1) / pass_bytes;
  const unsigned slab = (unsigned)(((a.rows_per_cta * row_bytes + 127) / 128) * 128);   // dy lives behind x
  unsigned char* sx = dyn;
  unsigned char* sd = dyn + slab;
  const char* xsrc = reinterpret_cast<const char*>(a.x) + base;
  const char* dsrc = PHASE == 1 ? reinterpret_cast<const char*>(a.dy) + base : nullptr;
  if (threadIdx.x == 0) {
    for (unsigned k = 0, off = 0; off < total; ++k, off += chunk_bytes) {
      const unsigned bytes = total - off < chunk_bytes ? total - off : chunk_bytes;
      mbar_expect_tx(bars + k, bytes * NSRC);
      bulk_g2s(sx + off, xsrc + off, bytes, bars + k);
      if (NSRC == 2) bulk_g2s(sd + off, dsrc + off, bytes, bars + k);
    }
  }
  const unsigned my_off = threadIdx.x * 16u;
  float s0[V], s1[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s0[i] = 0.f; s1[i] = 0.f; }

  if (PHASE == 0) {
    // ---- statistics while the chunks land ----
    for (unsigned ps = 0; ps < n_pass; ++ps) {
      if (ps % (unsigned)r.chunk_passes == 0) mbar_wait(bars + ps / (unsigned)r.chunk_passes, 0);
      const unsigned o = ps * pass_bytes + my_off;
      if (o < total) {
        float xv[V];
        unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
#pragma unroll
        for (int i = 0; i < V; ++i) { s0[i] += xv[i]; s1[i] = fmaf(xv[i], xv[i], s1[i]); }
      }
    }
    if (in_bias) {
      const float cnt = my_off < total ? (float)((total - my_off + pass_bytes - 1) / pass_bytes) : 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float bia = in_bias[c0 + i];
        s1[i] = s1[i] + 2.0f * bia * s0[i] + cnt * bia * bia;
        s0[i] = s0[i] + cnt * bia;
      }
    }
    group_totals<V, CLUSTER>(s0, s1, r, b, c0, cvec, rows, part, g_acc, g_part);
    if (threadIdx.x < a.G) {
      const int g = threadIdx.x;
      const double inv_n = 1.0 / ((double)Cg * (double)a.HW);
      const double m = g_acc[g][0] * inv_n;
      double var = g_acc[g][1] * inv_n - m * m; if (var < 0.0) var = 0.0;
      g_c0[g] = (float)m; g_c1[g] = (float)(1.0 / sqrt(var + (double)a.eps));
      if (blk == 0) {
        a.stats[((int64_t)b * a.G + g) * 2] = g_c0[g];
        a.stats[((int64_t)b * a.G + g) * 2 + 1] = g_c1[g];
      }
    }
    __syncthreads();
    float ca[V], cb[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = g_c0[g] - (a.in_bias ? in_bias[c] : 0.0f), rstd = g_c1[g];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
      ca[i] = rstd * gam * sc * kAff;
      cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    }
    char* dst = reinterpret_cast<char*>(a.out) + base;
#pragma unroll 4
    for (unsigned ps = 0; ps < n_pass; ++ps) {
      const unsigned o = ps * pass_bytes + my_off;
      if (o < total) {
        float xv[V], ov[V];
        unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
#pragma unroll
        for (int i = 0; i < V; ++i) ov[i] = gn_act_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
        store_vec<T>(reinterpret_cast<T*>(dst + o), ov);
      }
    }
  } else {
    // ---- backward: d = dy * act'(z) replaces dy in shared memory; t0 = sum d, t1 = sum d*x per channel ----
    float ca[V], cb[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (a.in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
      ca[i] = rstd * gam * sc * kAff;
      cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    }
    for (unsigned ps = 0; ps < n_pass; ++ps) {
      if (ps % (unsigned)r.chunk_passes == 0) mbar_wait(bars + ps / (unsigned)r.chunk_passes, 0);
      const unsigned o = ps * pass_bytes + my_off;
      if (o < total) {
        float xv[V], dv[V];
        unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
        unpack<T>(*reinterpret_cast<const uint4*>(sd + o), dv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          dv[i] = dv[i] * gn_dact_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
          s0[i] += dv[i]; s1[i] = fmaf(dv[i], xv[i], s1[i]);
        }
        if (MODE != NHMC_GN_PLAIN) store_vec<T>(reinterpret_cast<T*>(sd + o), dv);   // own 16 bytes: no sync needed
      }
    }
    float k1[V], k2[V], k3[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (a.in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float cgm = a.gamma[c] * sc;             // du = d * gamma * (1+scale); u = (x-mean)*rstd
      const float t0 = s0[i], t1 = s1[i];
      s0[i] = cgm * t0;
      s1[i] = cgm * rstd * (t1 - mean * t0);
      k1[i] = rstd * cgm; k2[i] = mean; k3[i] = rstd;   // finished below once the group means are known
    }
    group_totals<V, CLUSTER>(s0, s1, r, b, c0, cvec, rows, part, g_acc, g_part);
    if (threadIdx.x < a.G) {
      const int g = threadIdx.x;
      const double inv_n = 1.0 / ((double)Cg * (double)a.HW);
      g_c0[g] = (float)(g_acc[g][0] * inv_n); g_c1[g] = (float)(g_acc[g][1] * inv_n);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      // dx = rstd*(du - m1 - u*m2) = k1*d + k2*x + k3
      const int g = (c0 + i) / Cg;
      const float m1 = g_c0[g], m2 = g_c1[g], mean = k2[i], rstd = k3[i];
      k2[i] = -rstd * rstd * m2;
      k3[i] = -rstd * m1 + rstd * rstd * m2 * mean;
    }
    char* dst = reinterpret_cast<char*>(a.out) + base;
    const char* asrc = a.add ? reinterpret_cast<const char*>(a.add) + base : nullptr;
    constexpr int U = 4;
    for (unsigned ps0 = 0; ps0 < n_pass; ps0 += U) {
      uint4 av[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned o = (ps0 + u) * pass_bytes + my_off;
        av[u] = make_uint4(0, 0, 0, 0);
        if (asrc && o < total) av[u] = ld_raw(asrc + o);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned o = (ps0 + u) * pass_bytes + my_off;
        if (o < total) {
          float xv[V], dv[V], ov[V], ad[V];
          unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
          unpack<T>(*reinterpret_cast<const uint4*>(sd + o), dv);
          unpack<T>(av[u], ad);
#pragma unroll
          for (int i = 0; i < V; ++i) ov[i] = fmaf(k1[i], dv[i], fmaf(k2[i], xv[i], k3[i])) + ad[i];
          store_vec<T>(reinterpret_cast<T*>(dst + o), ov);
        }
      }
    }
  }
  if (CLUSTER) cluster_wait();    // every peer has finished reading this CTA's partial sums
}

// ============================== channels-last, resident by channel slice ======================
// Tensors too large for the single-wave resident kernel (the 128x128 / 256x256 maps) are cut into UNITS = (sample,
// slice of whole groups): groups are independent, so a slice of C/cs channels is a GroupNorm problem of its own.
// A unit is sized so that at least two units sit on the chip at once (each on its own set of SMs, one CTA per SM,
// ~220 KB of shared memory): while one unit waits on its barrier and writes its result, the next ones are loading,
// so HBM never idles, and -- unlike the statistics + apply pair -- x (and dy) are read exactly ONCE.  Rows of a slice
// are not contiguous in memory (stride C), so the loads are per-thread 16-byte cp.async copies (8 lanes cover a
// 128-byte segment); a thread only ever touches the bytes it copied itself, hence no block barriers in the hot loops.
// Units take their index from the start-order ticket (same forward-progress argument as above); bounded spin.
struct GnSliceArgs {
  GnArgs a;
  int cs;                   // slices per sample
  int Cs;                   // channels per slice (multiple of C/G and of the vector width)
  int nb;                   // CTAs per unit
  int64_t rows_per_cta;
  int chunk_passes;         // passes per cp.async commit group (4 groups)
  unsigned int total_ctas;
  unsigned int* ticket; unsigned int* arrive; unsigned int* done; unsigned int* err;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kSliceThreads = 256;   // 3 CTAs per SM (24 warps): the backward is within 2x of the issue limit, occupancy matters

// SOLO: every unit fits ONE CTA (nb == 1, the 8x8 .. 64x64 maps): no ticket, no global accumulators, no barrier --
// the unit index is the block index and the group totals never leave shared memory, so launches of any number of
// waves are safe and a call is one short dependency chain (load -> reduce -> normalise -> store).
// SYNC 2: a unit is ONE thread-block cluster of nb <= 8 CTAs (the 64x64 maps, whose units exceed one CTA's share of shared
// memory): the CTAs exchange their per-group partial sums through distributed shared memory behind the hardware cluster
// barrier -- still no ticket, no global accumulators, no spinning, and clusters never wait on each other.
template <typename T, int MODE, int PHASE, int SYNC = 0>
__global__ void __launch_bounds__(kSliceThreads, 3)
gn_nhwc_slice_kernel(const GnSliceArgs r) {
  constexpr bool SOLO = SYNC == 1;
  constexpr bool CLUSTER = SYNC == 2;
  if (SOLO) { pdl_wait(); pdl_trigger(); }     // units never wait on each other: safe to let the next kernel's CTAs in early
  if (CLUSTER) pdl_wait();
  constexpr int V = VecOf<T>::N;
  constexpr float kAff = AffineHalf<MODE, T>::value ? 0.5f : 1.0f;
  constexpr int NSRC = PHASE == 1 ? 2 : 1;
  const GnArgs& a = r.a;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ float4 part[kSliceThreads];
  __shared__ double g_acc[32][2];
  __shared__ float g_c0[32], g_c1[32];
  __shared__ unsigned int s_lid;
  __shared__ double g_part[CLUSTER ? 32 : 1][2];
  if (!SOLO && !CLUSTER) {
    if (threadIdx.x == 0) s_lid = atomicInc(r.ticket, r.total_ctas - 1u);
    __syncthreads();
  }
  const int unit = SOLO ? (int)blockIdx.x : CLUSTER ? (int)(blockIdx.x / (unsigned int)r.nb) : (int)(s_lid / (unsigned int)r.nb);
  const int blk = SOLO ? 0 : CLUSTER ? (int)(blockIdx.x % (unsigned int)r.nb) : (int)(s_lid % (unsigned int)r.nb);
  const int b = unit / r.cs, sl = unit % r.cs;
  const float* in_bias = a.in_bias ? a.in_bias + (int64_t)b * a.in_bias_stride : nullptr;
  const int cvec = r.Cs / V;                                   // vectors per row of the slice
  const int cv = threadIdx.x % cvec, rows = blockDim.x / cvec;
  const int c0 = sl * r.Cs + cv * V;                           // first channel of this thread (absolute)
  const int Cg = a.C / a.G;
  const int g_first = (sl * r.Cs) / Cg, g_count = r.Cs / Cg;   // groups of this slice
  const int64_t r_begin = (int64_t)blk * r.rows_per_cta;
  int64_t r_end = r_begin + r.rows_per_cta; if (r_end > a.HW) r_end = a.HW;
  const int n_rows = r_end > r_begin ? (int)(r_end - r_begin) : 0;
  const int64_t row_stride = (int64_t)a.C * sizeof(T);
  const int prow = threadIdx.x / cvec;
  const int64_t gbase = ((int64_t)b * a.HW + r_begin + prow) * row_stride + (int64_t)c0 * sizeof(T);   // this thread's first row
  const int n_mine = n_rows > prow ? (n_rows - prow + rows - 1) / rows : 0;     // rows this thread owns: prow, prow+rows, ...
  const unsigned pass_bytes = blockDim.x * 16u;
  const unsigned slab = (unsigned)((((r.rows_per_cta + rows - 1) / rows) * (int64_t)pass_bytes + 127) / 128 * 128);
  unsigned char* sx = dyn;
  unsigned char* sd = dyn + slab;
  const char* xg = reinterpret_cast<const char*>(a.x) + gbase;
  const char* dg = PHASE == 1 ? reinterpret_cast<const char*>(a.dy) + gbase : nullptr;
  const unsigned my_off = threadIdx.x * 16u;
  // ---- issue every copy of this thread, four commit groups
  {
    int k = 0;
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {
      const int k_end = (grp + 1) * r.chunk_passes < n_mine ? (grp + 1) * r.chunk_passes : n_mine;
      for (; k < k_end; ++k) {
        const int64_t go = (int64_t)(k * rows) * row_stride;
        cp_async16(sx + (unsigned)k * pass_bytes + my_off, xg + go);
        if (NSRC == 2) cp_async16(sd + (unsigned)k * pass_bytes + my_off, dg + go);
      }
      cp_async_commit();
    }
  }
  float s0[V], s1[V], ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s0[i] = 0.f; s1[i] = 0.f; ca[i] = 0.f; cb[i] = 0.f; }
  if (PHASE == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
      ca[i] = rstd * gam * sc * kAff;
      cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    }
  }
  // ---- pass 1 over the landed chunks (own bytes only: no block barrier)
  auto pass1 = [&](int k_lo, int k_hi) {
    for (int k = k_lo; k < k_hi; ++k) {
      const unsigned o = (unsigned)k * pass_bytes + my_off;
      float xv[V];
      unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
      if (PHASE == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) { s0[i] += xv[i]; s1[i] = fmaf(xv[i], xv[i], s1[i]); }
      } else {
        float dv[V];
        unpack<T>(*reinterpret_cast<const uint4*>(sd + o), dv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          dv[i] = dv[i] * gn_dact_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
          s0[i] += dv[i]; s1[i] = fmaf(dv[i], xv[i], s1[i]);
        }
        if (MODE != NHMC_GN_PLAIN) store_vec<T>(reinterpret_cast<T*>(sd + o), dv);
      }
    }
  };
  {
    const int cp = r.chunk_passes;
    auto hi = [&](int g) { return (g + 1) * cp < n_mine ? (g + 1) * cp : n_mine; };
    cp_async_wait<3>(); pass1(0, hi(0));
    cp_async_wait<2>(); pass1(hi(0), hi(1));
    cp_async_wait<1>(); pass1(hi(1), hi(2));
    cp_async_wait<0>(); pass1(hi(2), n_mine);
  }
  if (PHASE == 0 && in_bias) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float bia = in_bias[c0 + i];
      s1[i] = s1[i] + 2.0f * bia * s0[i] + (float)n_mine * bia * bia;
      s0[i] = s0[i] + (float)n_mine * bia;
    }
  }
  float k1[V], k2[V], k3[V];
  if (PHASE == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, g = c / Cg;
      const float mean = a.stats[((int64_t)b * a.G + g) * 2] - (in_bias ? in_bias[c] : 0.0f);
      const float rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float cgm = a.gamma[c] * sc;
      const float t0 = s0[i], t1 = s1[i];
      s0[i] = cgm * t0;
      s1[i] = cgm * rstd * (t1 - mean * t0);
      k1[i] = rstd * cgm; k2[i] = mean; k3[i] = rstd;
    }
  }
  // ---- per-CTA group partials -> global fp64 accumulators -> unit barrier -> unit totals
  {
    const int g0 = c0 / Cg;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if ((c0 + i) / Cg == g0) { q.x += s0[i]; q.y += s1[i]; } else { q.z += s0[i]; q.w += s1[i]; }
    }
    part[threadIdx.x] = q;
    __syncthreads();
    if ((int)threadIdx.x < g_count) {
      const int g = g_first + threadIdx.x;
      const int cl = g * Cg - sl * r.Cs;                        // first channel of the group, slice-relative
      const int cv_lo = cl / V, cv_hi = (cl + Cg - 1) / V;
      double t0 = 0.0, t1 = 0.0;
      for (int v = cv_lo; v <= cv_hi; ++v) {
        const int j = g - (sl * r.Cs + v * V) / Cg;
        for (int pr = 0; pr < rows; ++pr) {
          const float4 t = part[pr * cvec + v];
          t0 += (double)(j ? t.z : t.x); t1 += (double)(j ? t.w : t.y);
        }
      }
      if (SOLO) {
        g_acc[threadIdx.x][0] = t0; g_acc[threadIdx.x][1] = t1;
      } else if (CLUSTER) {
        g_part[threadIdx.x][0] = t0; g_part[threadIdx.x][1] = t1;       // read by every CTA of the cluster through DSMEM
      } else {
        atomicAdd(a.partials + ((int64_t)b * 64 + g) * 2, t0);
        atomicAdd(a.partials + ((int64_t)b * 64 + g) * 2 + 1, t1);
        __threadfence();
      }
    }
    __syncthreads();
    if (CLUSTER) {
      cluster_arrive();
      cluster_wait();
      if ((int)threadIdx.x < g_count) {
        double t0 = 0.0, t1 = 0.0;
        for (int k = 0; k < r.nb; ++k) {           // rank order: deterministic
          t0 += ld_dsmem_f64(&g_part[threadIdx.x][0], (unsigned)k); t1 += ld_dsmem_f64(&g_part[threadIdx.x][1], (unsigned)k);
        }
        g_acc[threadIdx.x][0] = t0; g_acc[threadIdx.x][1] = t1;
      }
      cluster_arrive();        // peers may exit (and release their shared memory) only after everybody has read: waited at kernel end
      __syncthreads();
    }
    if (!SOLO && !CLUSTER) {
    if (threadIdx.x == 0) {
      atomicAdd(r.arrive + unit, 1u);
      const long long t_start = clock64();
      while (ld_acquire_u32(r.arrive + unit) < (unsigned int)r.nb) {
        __nanosleep(64);
        if (clock64() - t_start > (1ll << 31)) { atomicExch(r.err, 1u); break; }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < g_count) {
      const int g = g_first + threadIdx.x;
      g_acc[threadIdx.x][0] = __ldcg(a.partials + ((int64_t)b * 64 + g) * 2);
      g_acc[threadIdx.x][1] = __ldcg(a.partials + ((int64_t)b * 64 + g) * 2 + 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const bool last = atomicInc(r.done + unit, (unsigned int)(r.nb - 1)) == (unsigned int)(r.nb - 1);
      if (last) {
        for (int g = g_first; g < g_first + g_count; ++g) { a.partials[((int64_t)b * 64 + g) * 2] = 0.0; a.partials[((int64_t)b * 64 + g) * 2 + 1] = 0.0; }
        r.arrive[unit] = 0u;
      }
    }
    }
  }
  const double inv_n = 1.0 / ((double)Cg * (double)a.HW);
  if ((int)threadIdx.x < g_count) {
    const int gl = threadIdx.x, g = g_first + gl;
    if (PHASE == 0) {
      const double m = g_acc[gl][0] * inv_n;
      double var = g_acc[gl][1] * inv_n - m * m; if (var < 0.0) var = 0.0;
      g_c0[gl] = (float)m; g_c1[gl] = (float)(1.0 / sqrt(var + (double)a.eps));
      if (blk == 0) {
        a.stats[((int64_t)b * a.G + g) * 2] = g_c0[gl];
        a.stats[((int64_t)b * a.G + g) * 2 + 1] = g_c1[gl];
      }
    } else {
      g_c0[gl] = (float)(g_acc[gl][0] * inv_n); g_c1[gl] = (float)(g_acc[gl][1] * inv_n);
    }
  }
  __syncthreads();
  char* og = reinterpret_cast<char*>(a.out) + gbase;
  if (PHASE == 0) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = c0 + i, gl = c / Cg - g_first;
      const float mean = g_c0[gl] - (in_bias ? in_bias[c] : 0.0f), rstd = g_c1[gl];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
      const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
      ca[i] = rstd * gam * sc * kAff;
      cb[i] = ((bet - mean * rstd * gam) * sc + sh) * kAff;
    }
#pragma unroll 4
    for (int k = 0; k < n_mine; ++k) {
      float xv[V], ov[V];
      unpack<T>(*reinterpret_cast<const uint4*>(sx + (unsigned)k * pass_bytes + my_off), xv);
#pragma unroll
      for (int i = 0; i < V; ++i) ov[i] = gn_act_a<MODE, T>(fmaf(ca[i], xv[i], cb[i]));
      store_vec<T>(reinterpret_cast<T*>(og + (int64_t)(k * rows) * row_stride), ov);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int gl = (c0 + i) / Cg - g_first;
      const float m1 = g_c0[gl], m2 = g_c1[gl], mean = k2[i], rstd = k3[i];
      k2[i] = -rstd * rstd * m2;
      k3[i] = -rstd * m1 + rstd * rstd * m2 * mean;
    }
    const char* ag = a.add ? reinterpret_cast<const char*>(a.add) + gbase : nullptr;
    constexpr int U = 4;
    for (int k0 = 0; k0 < n_mine; k0 += U) {
      uint4 av[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        av[u] = make_uint4(0, 0, 0, 0);
        if (ag && k0 + u < n_mine) av[u] = ld_raw(ag + (int64_t)((k0 + u) * rows) * row_stride);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u;
        if (k < n_mine) {
          const unsigned o = (unsigned)k * pass_bytes + my_off;
          float xv[V], dv[V], ov[V], ad[V];
          unpack<T>(*reinterpret_cast<const uint4*>(sx + o), xv);
          unpack<T>(*reinterpret_cast<const uint4*>(sd + o), dv);
          unpack<T>(av[u], ad);
#pragma unroll
          for (int i = 0; i < V; ++i) ov[i] = fmaf(k1[i], dv[i], fmaf(k2[i], xv[i], k3[i])) + ad[i];
          store_vec<T>(reinterpret_cast<T*>(og + (int64_t)(k * rows) * row_stride), ov);
        }
      }
    }
  }
  if (CLUSTER) cluster_wait();    // every peer has finished reading this CTA's partial sums
}

// ============================== NCHW =========================================================
// One CTA per (sample, group): the group is a contiguous slab of Cg*HW elements (<= 1 MB at FFHQ-256),
// read twice by the same CTA so the second read is an L2 hit.
template <typename T, int MODE, int PHASE>
__global__ void __launch_bounds__(256)
gn_nchw_kernel(GnArgs a) {
  const int g = blockIdx.x, b = blockIdx.y;
  const float* in_bias = a.in_bias ? a.in_bias + (int64_t)b * a.in_bias_stride : nullptr;
  const int Cg = a.C / a.G;
  const int64_t n = (int64_t)Cg * a.HW;
  const int64_t base = ((int64_t)b * a.C + (int64_t)g * Cg) * a.HW;
  const T* xb = reinterpret_cast<const T*>(a.x) + base;
  const T* dyb = PHASE == 1 ? reinterpret_cast<const T*>(a.dy) + base : nullptr;
  T* ob = reinterpret_cast<T*>(a.out) + base;
  __shared__ double red[2 * 32];
  __shared__ float bc[2];
  float mean, rstd;
  if (PHASE == 0) {
    double s[2] = {0.0, 0.0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = Elem<T>::to_f(xb[i]) + (a.in_bias ? in_bias[g * Cg + (int)(i / a.HW)] : 0.0f);
      s[0] += v; s[1] += (double)v * v;
    }
    block_sum<double, 2>(s, red);
    if (threadIdx.x == 0) {
      const double m = s[0] / (double)n;
      double var = s[1] / (double)n - m * m; if (var < 0.0) var = 0.0;
      bc[0] = (float)m; bc[1] = (float)(1.0 / sqrt(var + (double)a.eps));
      a.stats[((int64_t)b * a.G + g) * 2] = bc[0];
      a.stats[((int64_t)b * a.G + g) * 2 + 1] = bc[1];
    }
    __syncthreads();
    mean = bc[0]; rstd = bc[1];
  } else {
    mean = a.stats[((int64_t)b * a.G + g) * 2]; rstd = a.stats[((int64_t)b * a.G + g) * 2 + 1];
  }
  auto coef = [&](int c, float& ca, float& cb, float& cgm) {
    const float gam = a.gamma[c], bet = a.beta[c];
    const float sc = (MODE == NHMC_GN_FILM_SILU) ? 1.0f + a.scale[(int64_t)b * a.film_stride + c] : 1.0f;
    const float sh = (MODE == NHMC_GN_FILM_SILU) ? a.shift[(int64_t)b * a.film_stride + c] : 0.0f;
    ca = rstd * gam * sc; cb = (bet - mean * rstd * gam) * sc + sh; cgm = gam * sc;
  };
  if (PHASE == 0) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      float ca, cb, cgm; coef(g * Cg + (int)(i / a.HW), ca, cb, cgm);
      const float bia = a.in_bias ? in_bias[g * Cg + (int)(i / a.HW)] : 0.0f;
      ob[i] = Elem<T>::from_f(gn_act<MODE, T>(fmaf(ca, Elem<T>::to_f(xb[i]) + bia, cb)));
    }
  } else {
    double s[2] = {0.0, 0.0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      float ca, cb, cgm; coef(g * Cg + (int)(i / a.HW), ca, cb, cgm);
      const float xv = Elem<T>::to_f(xb[i]) + (a.in_bias ? in_bias[g * Cg + (int)(i / a.HW)] : 0.0f);
      const float du = Elem<T>::to_f(dyb[i]) * gn_dact<MODE, T>(fmaf(ca, xv, cb)) * cgm;
      s[0] += du; s[1] += (double)du * (double)((xv - mean) * rstd);
    }
    block_sum<double, 2>(s, red);
    if (threadIdx.x == 0) { bc[0] = (float)(s[0] / (double)n); bc[1] = (float)(s[1] / (double)n); }
    __syncthreads();
    const float m1 = bc[0], m2 = bc[1];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      float ca, cb, cgm; coef(g * Cg + (int)(i / a.HW), ca, cb, cgm);
      const float xv = Elem<T>::to_f(xb[i]) + (a.in_bias ? in_bias[g * Cg + (int)(i / a.HW)] : 0.0f);
      const float du = Elem<T>::to_f(dyb[i]) * gn_dact<MODE, T>(fmaf(ca, xv, cb)) * cgm;
      const float u = (xv - mean) * rstd;
      ob[i] = Elem<T>::from_f(rstd * (du - m1 - u * m2));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GN scratch (fp64 accumulators, tickets, barrier counters) is CALLER-OWNED device memory (include/nhmc.h:
// nhmc_gn_workspace_bytes, zero-filled once): the library keeps no device allocation and no mutable state of its own,
// so two host threads / samplers / lanes never share accumulators unless they share a workspace.  The execution-path
// switches that used to be process globals arrive as a per-call flag word.
struct GnCfg { bool resident, cluster, pdl; int slice, solo, lanes; };
struct GnScratch { double* partials; unsigned int* tickets; float* bstats; unsigned int* res; int batch; GnCfg cfg; };  // res: [ticket, err, ticket2, pad, arrive[nb], done[nb]]

constexpr size_t kGnWsPerSample = 64 * 2 * sizeof(double) + 64 * 2 * sizeof(float) + 3 * sizeof(unsigned int) + 4;   // 1552 B
constexpr size_t kGnWsHeader = 64;

static GnCfg parse_flags(unsigned int f) {
  GnCfg c;
  c.resident = !(f & NHMC_GN_F_NO_RESIDENT);
  c.cluster = !(f & NHMC_GN_F_NO_CLUSTER);
  c.slice = (f & NHMC_GN_F_SLICE_ALL) ? 2 : ((f & NHMC_GN_F_SLICE_BWD) ? 1 : 0);
  c.solo = (f & NHMC_GN_F_NO_SOLO) ? 0 : ((f & NHMC_GN_F_NO_SOLO_CLUSTER) ? 1 : 2);
  c.pdl = !(f & NHMC_F_NO_PDL);
  c.lanes = (int)((f >> NHMC_GN_F_LANES_SHIFT) & 7u) + 1;
  return c;
}

static int get_scratch(void* workspace, size_t workspace_bytes, unsigned int flags, int batch, GnScratch& out) {
  NHMC_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "GroupNorm workspace must be a 16-byte aligned device buffer");
  NHMC_CHECK_ARG(workspace_bytes >= kGnWsHeader + kGnWsPerSample * (size_t)(batch < 64 ? 64 : batch),
                 "GroupNorm workspace too small for this batch: size it with nhmc_gn_workspace_bytes(max_batch)");
  const size_t nb = (workspace_bytes - kGnWsHeader) / kGnWsPerSample;
  char* p = reinterpret_cast<char*>(workspace);
  out.res = reinterpret_cast<unsigned int*>(p);                         // 4 header words, then arrive[nb], done[nb] behind the arrays
  out.partials = reinterpret_cast<double*>(p + kGnWsHeader);
  out.bstats = reinterpret_cast<float*>(out.partials + nb * 64 * 2);
  out.tickets = reinterpret_cast<unsigned int*>(out.bstats + nb * 64 * 2);
  out.batch = (int)nb;
  out.cfg = parse_flags(flags);
  return NHMC_OK;
}
// arrive / done counters of the barrier kernels live behind the ticket array
static inline unsigned int* ws_arrive(const GnScratch& s) { return s.tickets + s.batch; }
static inline unsigned int* ws_done(const GnScratch& s) { return s.tickets + 2 * (size_t)s.batch; }

struct DevInfo { int sms, smem_sm, smem_optin, smem_reserved; bool ok; };
static DevInfo dev_info() {
  static DevInfo info[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  DevInfo& d = info[dev];
  if (!d.ok) {
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
    d.ok = d.sms > 0;
  }
  return d;
}


// Returns NHMC_OK after launching, NHMC_ERR_UNSUPPORTED when the shape does not fit the resident scheme
// (the caller then takes the two-kernel path).
template <typename T, int MODE, int PHASE>
static int run_nhwc_resident(const GnArgs& a0, const GnScratch& s, int batch, int threads, cudaStream_t st) {
  constexpr int V = VecOf<T>::N;
  constexpr int NSRC = PHASE == 1 ? 2 : 1;
  if (!s.cfg.resident || threads > 256 || a0.G > 32) return NHMC_ERR_UNSUPPORTED;
  const int Cg = a0.C / a0.G, cvec = a0.C / V;
  for (int cv = 0; cv < cvec; ++cv)
    if ((cv * V + V - 1) / Cg - (cv * V) / Cg > 1) return NHMC_ERR_UNSUPPORTED;   // a thread may span at most 2 groups
  const DevInfo d = dev_info();
  if (!d.ok) return NHMC_ERR_UNSUPPORTED;
  auto kern = gn_nhwc_resident_kernel<T, MODE, PHASE, false>;
  static int static_smem = -1;   // per instantiation
  if (static_smem < 0) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return NHMC_ERR_UNSUPPORTED; }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes) != cudaSuccess) {
      cudaGetLastError(); return NHMC_ERR_UNSUPPORTED;
    }
    static_smem = (int)fa.sharedSizeBytes;
  }
  const int64_t row_bytes = (int64_t)a0.C * sizeof(T);
  const int64_t pass_bytes = (int64_t)threads * 16;
  // (a) One sample per thread-block cluster (<= 8 CTAs, the portable limit): the per-sample reduction runs over
  // distributed shared memory behind the hardware cluster barrier -- no global atomics, no ticket, no spinning.
  // Taken whenever a sample's slab fits in 8 CTAs' shared memory (the 8x8 .. 32x32 maps at FFHQ-256).
  if (s.cfg.cluster) {
    auto ckern = gn_nhwc_resident_kernel<T, MODE, PHASE, true>;
    static int c_static = -1;
    if (c_static < 0) {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, ckern) == cudaSuccess &&
          cudaFuncSetAttribute(ckern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes) == cudaSuccess)
        c_static = (int)fa.sharedSizeBytes;
      else { cudaGetLastError(); c_static = 1 << 30; }
    }
    const int64_t sample_bytes = a0.HW * row_bytes;
    const int64_t limit = ((int64_t)(d.smem_optin - c_static) / NSRC) & ~(int64_t)127;
    int64_t nb = (sample_bytes + 32767) / 32768;                    // ~32 KiB of input per CTA ...
    if (nb > 8) nb = 8;                                             // ... but at most 8 CTAs per sample
    if (nb < 1) nb = 1;
    int64_t rows_per_cta = (a0.HW + nb - 1) / nb;
    nb = (a0.HW + rows_per_cta - 1) / rows_per_cta;
    if (limit > 0 && rows_per_cta * row_bytes <= limit) {
      const int64_t slab = ((rows_per_cta * row_bytes + 127) / 128) * 128;
      GnResArgs r{};
      r.a = a0;
      r.a.rows_per_cta = rows_per_cta;
      r.nb = (int)nb;
      const int64_t n_pass = (rows_per_cta * row_bytes + pass_bytes - 1) / pass_bytes;
      r.chunk_passes = (int)((n_pass + kResChunks - 1) / kResChunks);
      r.total_ctas = (unsigned int)(nb * batch);
      r.reps = 1; r.batch = batch;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(r.total_ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (size_t)slab * NSRC; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)nb; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, ckern, &cfg) == cudaSuccess && n_clusters >= 1) {
        if (int e = cuda_status(cudaLaunchKernelEx(&cfg, ckern, r), "cluster launch")) return e;
        return NHMC_OK;
      }
      cudaGetLastError();
    }
  }
  // (b) Single wave only: every CTA of the launch must be on the chip at once (k CTAs per SM).  Larger tensors go
  // through the statistics + apply kernels, whose second read is served by the L2 for what still fits there
  // (measured in situ: the resident scheme loses once a launch needs several waves, because the per-sample
  // barrier then serialises load -> reduce -> store phases that the two-kernel path overlaps).
  if (s.cfg.lanes > 1) return NHMC_ERR_UNSUPPORTED;   // concurrent lanes never run the global-barrier kernel
  for (int k = 4; k >= 1; --k) {
    int64_t dyn_limit = d.smem_sm / k - d.smem_reserved;
    if (dyn_limit > d.smem_optin) dyn_limit = d.smem_optin;
    dyn_limit -= static_smem;
    const int64_t slab_limit = (dyn_limit / NSRC) & ~(int64_t)127;
    if (slab_limit < row_bytes) continue;
    const int64_t slots = ((int64_t)k * d.sms) / s.cfg.lanes;         // this lane's share of the chip
    int64_t nb = slots / batch;                                   // CTAs per sample that keep the launch in one wave
    if (nb < 1) continue;
    int64_t rows_per_cta = (a0.HW + nb - 1) / nb;
    if (rows_per_cta * row_bytes < 8192) rows_per_cta = (8192 + row_bytes - 1) / row_bytes;   // never below 8 KiB per CTA
    if (rows_per_cta > a0.HW) rows_per_cta = a0.HW;
    if (rows_per_cta * row_bytes > slab_limit) continue;
    nb = (a0.HW + rows_per_cta - 1) / rows_per_cta;
    const int64_t slab = ((rows_per_cta * row_bytes + 127) / 128) * 128;
    const size_t dyn = (size_t)slab * NSRC;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, dyn) != cudaSuccess) { cudaGetLastError(); continue; }
    if (((int64_t)occ * d.sms) / s.cfg.lanes < nb * batch) continue;
    GnResArgs r{};
    r.a = a0;
    r.a.rows_per_cta = rows_per_cta;
    r.nb = (int)nb;
    const int64_t n_pass = (rows_per_cta * row_bytes + pass_bytes - 1) / pass_bytes;
    r.chunk_passes = (int)((n_pass + kResChunks - 1) / kResChunks);
    r.total_ctas = (unsigned int)(nb * batch);
    r.batch = batch;
    r.reps = (int)(nb >= 64 ? 8 : (nb >= 16 ? 4 : 1));             // replicas only where a sample has many CTAs
    while (r.reps > 1 && (int64_t)r.reps * batch > s.batch) r.reps >>= 1;
    r.ticket = s.res; r.err = s.res + 1; r.arrive = ws_arrive(s); r.done = ws_done(s);
    // cooperative launch: the driver guarantees that every CTA of the grid is resident at once (or fails the launch), so
    // the per-sample spin barrier cannot starve behind CTAs of other kernels
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(r.total_ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cuda_status(cudaLaunchKernelEx(&cfg, kern, r), "resident GroupNorm launch");
  }
  return NHMC_ERR_UNSUPPORTED;
}


// Resident-by-channel-slice kernel for tensors that do not fit the single-wave resident kernel.  NHMC_ERR_UNSUPPORTED
// when the shape cannot be cut into units of which at least two fit on the chip at once.
template <typename T, int MODE, int PHASE>
static int run_nhwc_slice(const GnArgs& a0, const GnScratch& s, int batch, cudaStream_t st) {
  constexpr int V = VecOf<T>::N;
  constexpr int NSRC = PHASE == 1 ? 2 : 1;
  const int mode = s.cfg.slice;
  if (!s.cfg.resident || mode == 0 || a0.x1 || a0.G > 32) return NHMC_ERR_UNSUPPORTED;
  const int64_t row_bytes = (int64_t)a0.C * sizeof(T);
  const int64_t sample_bytes = a0.HW * row_bytes;
  if (PHASE == 0 && (mode < 2 || (int64_t)batch * sample_bytes < ((int64_t)200 << 20))) return NHMC_ERR_UNSUPPORTED;
  const DevInfo d = dev_info();
  if (!d.ok) return NHMC_ERR_UNSUPPORTED;
  auto kern = gn_nhwc_slice_kernel<T, MODE, PHASE>;
  static int static_smem = -1;
  if (static_smem < 0) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes) != cudaSuccess) {
      cudaGetLastError(); return NHMC_ERR_UNSUPPORTED;
    }
    static_smem = (int)fa.sharedSizeBytes;
  }
  const int Cg = a0.C / a0.G;
  for (int occ = 3; occ >= 1; --occ) {
    int64_t dyn_limit = d.smem_sm / occ - d.smem_reserved;
    if (dyn_limit > d.smem_optin) dyn_limit = d.smem_optin;
    dyn_limit -= static_smem;
    const int64_t slab_limit = (dyn_limit / NSRC) & ~(int64_t)127;
    const int64_t slots = (int64_t)occ * d.sms;
    for (int cs = 1; cs <= 8; cs <<= 1) {
      if (a0.G % cs || (int64_t)batch * cs > s.batch) continue;
      const int Cs = a0.C / cs;
      if (Cs % Cg || Cs % V) continue;
      const int cvec = Cs / V;
      if (cvec > kSliceThreads) continue;
      const int rows = kSliceThreads / cvec, threads = rows * cvec;
      bool ok = true;
      for (int cv = 0; cv < cvec && ok; ++cv) ok = ((cv * V + V - 1) / Cg - (cv * V) / Cg) <= 1;   // <= 2 groups per thread
      if (!ok) continue;
      const int64_t pass_bytes = (int64_t)threads * 16;
      const int64_t unit_bytes = sample_bytes / cs;
      int64_t nb = (unit_bytes + slab_limit - 1) / slab_limit;
      if (nb < 1) nb = 1;
      if (nb > slots / 2) continue;                                // at least two units on the chip at once
      const int64_t U = slots / nb;
      nb = slots / U;                                              // the resident units cover all slots
      int64_t rows_per_cta = 0, slab = 0;
      for (; nb <= slots / 2; ++nb) {
        rows_per_cta = (a0.HW + nb - 1) / nb;
        slab = (((rows_per_cta + rows - 1) / rows) * pass_bytes + 127) / 128 * 128;
        if (slab <= slab_limit) break;
      }
      if (nb > slots / 2 || slab > slab_limit) continue;
      nb = (a0.HW + rows_per_cta - 1) / rows_per_cta;
      int got = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, threads, (size_t)slab * NSRC) != cudaSuccess) { cudaGetLastError(); continue; }
      if ((int64_t)got * d.sms < 2 * nb) continue;                 // two whole units must be resident
      GnSliceArgs r{};
      r.a = a0;
      r.cs = cs; r.Cs = Cs; r.nb = (int)nb; r.rows_per_cta = rows_per_cta;
      const int64_t n_pass = (rows_per_cta + rows - 1) / rows;
      r.chunk_passes = (int)((n_pass + 3) / 4);
      r.total_ctas = (unsigned int)(nb * batch * cs);
      r.ticket = s.res + 2; r.err = s.res + 1; r.arrive = ws_arrive(s); r.done = ws_done(s);
      kern<<<r.total_ctas, threads, (size_t)slab * NSRC, st>>>(r);
      return NHMC_LAUNCH_CHECK();
    }
  }
  return NHMC_ERR_UNSUPPORTED;
}


// One CTA per (sample, slice of whole groups) with the unit resident in shared memory: the small maps (8x8 .. 64x64),
// where a call is bound by its dependency chain, not by bandwidth.  Picks the WIDEST slice (longest contiguous run per
// pixel) that still yields >= ~one CTA per SM and whose unit fits 1/3 of an SM's shared memory (3 CTAs per SM overlap
// each other's load / reduce / store phases); NHMC_ERR_UNSUPPORTED otherwise.
template <typename T, int MODE, int PHASE>
static int run_nhwc_solo(const GnArgs& a0, const GnScratch& s, int batch, cudaStream_t st) {
  constexpr int V = VecOf<T>::N;
  constexpr int NSRC = PHASE == 1 ? 2 : 1;
  if (!s.cfg.solo || a0.x1 || a0.G > 32) return NHMC_ERR_UNSUPPORTED;
  const DevInfo d = dev_info();
  if (!d.ok) return NHMC_ERR_UNSUPPORTED;
  auto kern = gn_nhwc_slice_kernel<T, MODE, PHASE, 1>;
  static int static_smem = -1;
  if (static_smem < 0) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes) != cudaSuccess) {
      cudaGetLastError(); return NHMC_ERR_UNSUPPORTED;
    }
    static_smem = (int)fa.sharedSizeBytes;
  }
  const int Cg = a0.C / a0.G;
  const int64_t budget = (d.smem_sm / 3 - d.smem_reserved - static_smem) & ~(int64_t)127;    // 3 CTAs per SM (~70 KiB)
  int best_cs = 0, best_threads = 0; int64_t best_slab = 0;
  for (int cs = 1; cs <= a0.G; cs <<= 1) {
    if (a0.G % cs) continue;
    const int Cs = a0.C / cs;
    if (Cs % Cg || Cs % V) continue;
    const int cvec = Cs / V;
    if (cvec > kSliceThreads) continue;
    bool ok = true;
    for (int cv = 0; cv < cvec && ok; ++cv) ok = ((cv * V + V - 1) / Cg - (cv * V) / Cg) <= 1;   // <= 2 groups per thread
    if (!ok) continue;
    int rows = kSliceThreads / cvec;
    if (rows > a0.HW) rows = (int)a0.HW;
    const int threads = rows * cvec;
    const int64_t n_pass = (a0.HW + rows - 1) / rows;
    const int64_t slab = ((n_pass * threads * 16 + 127) / 128) * 128;
    if (slab * NSRC > budget) continue;
    // measured (tools/kernel_microbench.py --small --warm --graph-calls 20, SOLO on/off): the one-CTA kernel wins while a
    // unit is <= 64 KiB AND a pixel's run is at least one 32-byte sector; with narrower runs only for tiny units
    const int64_t unit_bytes = slab * NSRC, run_bytes = (int64_t)Cs * sizeof(T);
    if (!(unit_bytes <= 8192 || (unit_bytes <= 65536 && run_bytes >= 32))) continue;
    best_cs = cs; best_threads = threads; best_slab = slab;
    if ((int64_t)batch * cs >= (int64_t)d.sms - 20) break;           // enough CTAs: keep the widest such slice
  }
  if (!best_cs) {
    // ---- one thread-block cluster per unit (2, 4 or 8 CTAs of <= 64 KiB each): the 64x64 maps ----
    // measured (tools/kernel_microbench.py --small --warm --graph-calls 20 --solo 1/2): wins on the 64x64 maps with >= 64-byte
    // runs per pixel while x (+ dy) stay <= 36 MB in total (10.2 vs 15.2 us forward, 24.4 vs 26.9 us backward at
    // [8,256,64,64]); loses on 32x32 maps (resident kernel: 16.5 vs 22.6 us), with 32-byte runs and beyond that size
    if (s.cfg.solo < 2 || a0.HW < 4096 || (int64_t)batch * a0.HW * a0.C * (int64_t)sizeof(T) * NSRC > ((int64_t)36 << 20)) return NHMC_ERR_UNSUPPORTED;
    auto ckern = gn_nhwc_slice_kernel<T, MODE, PHASE, 2>;
    static int c_static = -1;
    if (c_static < 0) {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, ckern) == cudaSuccess &&
          cudaFuncSetAttribute(ckern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - (int)fa.sharedSizeBytes) == cudaSuccess)
        c_static = (int)fa.sharedSizeBytes;
      else { cudaGetLastError(); c_static = 1 << 30; }
    }
    if (c_static >= (1 << 30)) return NHMC_ERR_UNSUPPORTED;
    for (int cs = 1; cs <= a0.G; cs <<= 1) {                       // widest slice first
      if (a0.G % cs) continue;
      const int Cs = a0.C / cs;
      if (Cs % Cg || Cs % V || (int64_t)Cs * sizeof(T) < 64) continue;
      const int cvec = Cs / V;
      if (cvec > kSliceThreads) continue;
      bool ok = true;
      for (int cv = 0; cv < cvec && ok; ++cv) ok = ((cv * V + V - 1) / Cg - (cv * V) / Cg) <= 1;
      if (!ok) continue;
      const int rows = kSliceThreads / cvec, threads = rows * cvec;
      for (int nb = 2; nb <= 8; nb <<= 1) {
        const int64_t rows_per_cta = (a0.HW + nb - 1) / nb;
        const int64_t n_pass = (rows_per_cta + rows - 1) / rows;
        const int64_t slab = ((n_pass * threads * 16 + 127) / 128) * 128;
        if (slab * NSRC > 65536) continue;
        if ((a0.HW + rows_per_cta - 1) / rows_per_cta != nb) continue;      // every rank owns rows
        GnSliceArgs r{};
        r.a = a0;
        r.cs = cs; r.Cs = Cs; r.nb = nb; r.rows_per_cta = rows_per_cta;
        r.chunk_passes = (int)((n_pass + 3) / 4);
        r.total_ctas = (unsigned int)(batch * cs * nb);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(r.total_ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (size_t)slab * NSRC; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)nb; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, ckern, &cfg) != cudaSuccess || n_clusters < 1) { cudaGetLastError(); continue; }
        return cuda_status(cudaLaunchKernelEx(&cfg, ckern, r), "gn cluster-unit launch");
      }
    }
    return NHMC_ERR_UNSUPPORTED;
  }
  GnSliceArgs r{};
  r.a = a0;
  r.cs = best_cs; r.Cs = a0.C / best_cs; r.nb = 1; r.rows_per_cta = a0.HW;
  const int rows = best_threads / (r.Cs / V);
  const int64_t n_pass = (a0.HW + rows - 1) / rows;
  r.chunk_passes = (int)((n_pass + 3) / 4);
  r.total_ctas = (unsigned int)(batch * best_cs);
  // Programmatic launch only for launches of small units or at most one CTA per SM: with several large-unit CTAs per SM
  // the dependents that slip in early crowd some SMs and the launch ends unbalanced (measured: 15.5 -> 18.4 us at
  // [8,512,32,32] backward); a plain launch spreads evenly over the idle chip.
  if ((int64_t)r.total_ctas <= d.sms || (int64_t)best_slab * NSRC <= 16384)
    return cuda_status(launch_pdl(s.cfg.pdl, kern, dim3(r.total_ctas), dim3(best_threads), (size_t)best_slab * NSRC, st, r), "gn one-CTA launch");
  kern<<<r.total_ctas, best_threads, (size_t)best_slab * NSRC, st>>>(r);
  return NHMC_LAUNCH_CHECK();
}

template <typename T, int MODE, int PHASE>
static int run_nhwc(GnArgs a, const GnScratch& s, int batch, cudaStream_t st) {
  constexpr int V = VecOf<T>::N;
  const int cvec = a.C / V;
  int rows = 256 / cvec; if (rows < 1) rows = 1;
  const int threads = cvec * rows;
  NHMC_CHECK_ARG(threads <= 1024, "too many channels for the channels-last GroupNorm mapping");
  if (!a.x1 && !a.add2) {
    int e = run_nhwc_solo<T, MODE, PHASE>(a, s, batch, st);
    if (e != NHMC_ERR_UNSUPPORTED) return e;
    e = run_nhwc_resident<T, MODE, PHASE>(a, s, batch, threads, st);
    if (e != NHMC_ERR_UNSUPPORTED) return e;
    e = run_nhwc_slice<T, MODE, PHASE>(a, s, batch, st);
    if (e != NHMC_ERR_UNSUPPORTED) return e;
  }
  // Two kernels (statistics, apply).  The grid is ONE balanced wave: (CTAs per SM the statistics kernel can keep
  // resident) x (SM count) CTAs in total, each streaming a contiguous row range of one sample.
  const int64_t bytes_per_row = (int64_t)a.C * sizeof(T);
  const size_t red = (size_t)rows * a.C * 2 * sizeof(float);
  NHMC_CHECK_ARG(red <= 48 * 1024, "GroupNorm reduction scratch too large");
  static int occ = -1;             // per instantiation (depends on registers and the reduction scratch only)
  if (occ < 0) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, gn_nhwc_stats_kernel<T, MODE, PHASE>, 256, 16384) != cudaSuccess) { cudaGetLastError(); o = 2; }
    occ = o < 1 ? 1 : o;
  }
  const DevInfo d = dev_info();
  const int sms = d.ok ? d.sms : kSMs;
  int64_t nblk = ((int64_t)occ * sms) / batch;                     // CTAs per sample
  if (nblk < 1) nblk = 1;
  int64_t rows_per_cta = (a.HW + nblk - 1) / nblk;
  if (rows_per_cta * bytes_per_row < 16384) rows_per_cta = (16384 + bytes_per_row - 1) / bytes_per_row;   // >= 16 KiB per CTA
  rows_per_cta = ((rows_per_cta + rows - 1) / rows) * rows;
  nblk = (a.HW + rows_per_cta - 1) / rows_per_cta;
  a.rows_per_cta = rows_per_cta;
  a.rows_per_cta_iter = rows;
  a.b0 = 0;
  a.reps = nblk >= 64 ? 8 : (nblk >= 16 ? 4 : 1);
  while (a.reps > 1 && (int64_t)a.reps * batch > s.batch) a.reps >>= 1;
  const dim3 grid((unsigned)nblk, (unsigned)batch);
  if (int e = cuda_status(launch_pdl(s.cfg.pdl, gn_nhwc_stats_kernel<T, MODE, PHASE>, grid, dim3(threads), red, st, a), "gn statistics launch")) return e;
  if (PHASE == 1 && a.add2) {
    NHMC_CHECK_ARG(a.add && !a.x1, "dx_add2 needs dx_add and a single-source tensor");
    return cuda_status(launch_pdl(s.cfg.pdl, gn_nhwc_apply_kernel<T, MODE, PHASE, PHASE == 1>, grid, dim3(threads), 0, st, a), "gn apply launch");
  }
  return cuda_status(launch_pdl(s.cfg.pdl, gn_nhwc_apply_kernel<T, MODE, PHASE>, grid, dim3(threads), 0, st, a), "gn apply launch");
}

template <typename T, int MODE, int PHASE>
static int run_any(int layout, GnArgs a, const GnScratch& s, int batch, cudaStream_t st) {
  if (layout == NHMC_LAYOUT_NHWC) return run_nhwc<T, MODE, PHASE>(a, s, batch, st);
  NHMC_CHECK_ARG(!a.add && !a.add2, "dx_add needs the channels-last layout");
  gn_nchw_kernel<T, MODE, PHASE><<<dim3(a.G, batch), 256, 0, st>>>(a);
  return NHMC_LAUNCH_CHECK();
}

template <int PHASE>
static int dispatch(int mode, int layout, int dtype, GnArgs a, const GnScratch& s, int batch, cudaStream_t st) {
#define GN_CASE(T)                                                                      \
  switch (mode) {                                                                       \
    case NHMC_GN_PLAIN: return run_any<T, NHMC_GN_PLAIN, PHASE>(layout, a, s, batch, st);   \
    case NHMC_GN_SILU: return run_any<T, NHMC_GN_SILU, PHASE>(layout, a, s, batch, st);     \
    case NHMC_GN_FILM_SILU: return run_any<T, NHMC_GN_FILM_SILU, PHASE>(layout, a, s, batch, st); \
  }
  switch (dtype) {
    case NHMC_F32: GN_CASE(float) break;
    case NHMC_BF16: GN_CASE(__nv_bfloat16) break;
    case NHMC_F16: GN_CASE(__half) break;
  }
#undef GN_CASE
  set_error("gn: unknown mode %d / dtype %d", mode, dtype);
  return NHMC_ERR_ARG;
}

static int check_common(int mode, int layout, int dtype, const void* x, const void* out, const float* gamma,
                        const float* beta, const float* scale, const float* shift, int batch, int channels,
                        int64_t hw, int groups) {
  NHMC_CHECK_ARG(x && out && gamma && beta, "null pointer");
  NHMC_CHECK_ARG(batch > 0 && channels > 0 && hw > 0 && groups > 0 && groups <= 64 && channels % groups == 0, "bad geometry");
  NHMC_CHECK_ARG(mode != NHMC_GN_FILM_SILU || (scale && shift), "FiLM needs scale and shift");
  NHMC_CHECK_ARG(layout == NHMC_LAYOUT_NCHW || layout == NHMC_LAYOUT_NHWC, "unknown layout");
  if (layout == NHMC_LAYOUT_NHWC) {
    const int V = dtype == NHMC_F32 ? 4 : 8;
    NHMC_CHECK_ARG(channels % V == 0 && aligned16(x) && aligned16(out), "channels-last needs C % (16B/elem) == 0 and 16-byte alignment");
  }
  return NHMC_OK;
}

}  // namespace nhmc

using namespace nhmc;

extern "C" int nhmc_gn_fwd(int mode, int layout, int dtype, const void* x, void* y, const float* gamma,
                           const float* beta, const float* scale, const float* shift, int64_t film_stride, const float* in_bias,
                           int64_t in_bias_stride, float* stats_out, int batch, int channels, int64_t hw, int groups, float eps,
                           void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream) {
  if (int e = check_common(mode, layout, dtype, x, y, gamma, beta, scale, shift, batch, channels, hw, groups)) return e;
  NHMC_CHECK_ARG(stats_out, "stats_out required");
  GnScratch s;
  if (int e = get_scratch(workspace, workspace_bytes, flags, batch, s)) return e;
  GnArgs a{};
  a.x = x; a.out = y; a.gamma = gamma; a.beta = beta; a.scale = scale; a.shift = shift; a.in_bias = in_bias; a.in_bias_stride = in_bias_stride; a.stats = stats_out;
  a.partials = s.partials; a.tickets = s.tickets; a.C = channels; a.G = groups; a.HW = hw; a.eps = eps;
  a.film_stride = film_stride > 0 ? film_stride : channels;
  NHMC_CHECK_ARG(a.film_stride >= channels, "film_stride must be >= channels");
  return dispatch<0>(mode, layout, dtype, a, s, batch, (cudaStream_t)stream);
}

extern "C" int nhmc_gn_bwd(int mode, int layout, int dtype, const void* x, const void* dy, void* dx,
                           const float* gamma, const float* beta, const float* scale, const float* shift,
                           int64_t film_stride, const float* in_bias, int64_t in_bias_stride, const float* stats, float* dscale,
                           float* dshift, const void* dx_add, const void* dx_add2, int batch,
                           int channels, int64_t hw, int groups, void* workspace, size_t workspace_bytes, unsigned int flags,
                           nhmc_stream_t stream) {
  if (int e = check_common(mode, layout, dtype, x, dx, gamma, beta, scale, shift, batch, channels, hw, groups)) return e;
  NHMC_CHECK_ARG(dy && stats, "dy and stats required");
  NHMC_CHECK_ARG(!dscale && !dshift, "dscale/dshift are not produced on the sampling path (timestep embedding carries no gradient)");
  NHMC_CHECK_ARG(layout != NHMC_LAYOUT_NHWC || aligned16(dy), "dy must be 16-byte aligned");
  NHMC_CHECK_ARG(!dx_add || aligned16(dx_add), "dx_add must be 16-byte aligned");
  NHMC_CHECK_ARG(!dx_add2 || (dx_add && aligned16(dx_add2)), "dx_add2 needs dx_add and 16-byte alignment");
  GnScratch s;
  if (int e = get_scratch(workspace, workspace_bytes, flags, batch, s)) return e;
  GnArgs a{};
  a.x = x; a.dy = dy; a.out = dx; a.gamma = gamma; a.beta = beta; a.scale = scale; a.shift = shift; a.in_bias = in_bias; a.in_bias_stride = in_bias_stride;
  a.stats = const_cast<float*>(stats);
  a.bstats = s.bstats;
  a.partials = s.partials; a.tickets = s.tickets; a.C = channels; a.G = groups; a.HW = hw; a.eps = 0.f;
  a.film_stride = film_stride > 0 ? film_stride : channels;
  NHMC_CHECK_ARG(a.film_stride >= channels, "film_stride must be >= channels");
  a.add = dx_add; a.add2 = dx_add2;
  return dispatch<1>(mode, layout, dtype, a, s, batch, (cudaStream_t)stream);
}

// Virtual concatenation: GroupNorm over cat([x0, x1], channel) without materialising the concatenation
// (channels-last only, statistics + apply kernels).
extern "C" int nhmc_gn_cat_fwd(int mode, int dtype, const void* x0, const void* x1, int c0, int c1, void* y,
                               const float* gamma, const float* beta, const float* scale, const float* shift,
                               int64_t film_stride, float* stats_out, int batch, int64_t hw, int groups, float eps,
                               void* workspace, size_t workspace_bytes, unsigned int flags, nhmc_stream_t stream) {
  const int channels = c0 + c1;
  if (int e = check_common(mode, NHMC_LAYOUT_NHWC, dtype, x0, y, gamma, beta, scale, shift, batch, channels, hw, groups)) return e;
  const int V = dtype == NHMC_F32 ? 4 : 8;
  NHMC_CHECK_ARG(x1 && c0 > 0 && c1 > 0 && c0 % V == 0 && c1 % V == 0 && aligned16(x1), "both blocks need C % (16B/elem) == 0 and 16-byte alignment");
  NHMC_CHECK_ARG(stats_out, "stats_out required");
  GnScratch s;
  if (int e = get_scratch(workspace, workspace_bytes, flags, batch, s)) return e;
  GnArgs a{};
  a.x = x0; a.x1 = x1; a.c_split = c0; a.out = y; a.gamma = gamma; a.beta = beta; a.scale = scale; a.shift = shift; a.stats = stats_out;
  a.partials = s.partials; a.tickets = s.tickets; a.C = channels; a.G = groups; a.HW = hw; a.eps = eps;
  a.film_stride = film_stride > 0 ? film_stride : channels;
  return dispatch<0>(mode, NHMC_LAYOUT_NHWC, dtype, a, s, batch, (cudaStream_t)stream);
}

extern "C" int nhmc_gn_cat_bwd(int mode, int dtype, const void* x0, const void* x1, int c0, int c1, const void* dy,
                               void* dx0, void* dx1, const float* gamma, const float* beta, const float* scale,
                               const float* shift, int64_t film_stride, const float* stats, const void* add0,
                               const void* add1, int batch, int64_t hw, int groups, void* workspace, size_t workspace_bytes,
                               unsigned int flags, nhmc_stream_t stream) {
  const int channels = c0 + c1;
  if (int e = check_common(mode, NHMC_LAYOUT_NHWC, dtype, x0, dx0, gamma, beta, scale, shift, batch, channels, hw, groups)) return e;
  const int V = dtype == NHMC_F32 ? 4 : 8;
  NHMC_CHECK_ARG(x1 && dx1 && c0 > 0 && c1 > 0 && c0 % V == 0 && c1 % V == 0 && aligned16(x1) && aligned16(dx1),
                 "both blocks need C % (16B/elem) == 0 and 16-byte alignment");
  NHMC_CHECK_ARG(dy && stats && aligned16(dy), "dy (16-byte aligned) and stats required");
  NHMC_CHECK_ARG((!add0 || aligned16(add0)) && (!add1 || aligned16(add1)), "add blocks must be 16-byte aligned");
  GnScratch s;
  if (int e = get_scratch(workspace, workspace_bytes, flags, batch, s)) return e;
  GnArgs a{};
  a.x = x0; a.x1 = x1; a.c_split = c0; a.dy = dy; a.out = dx0; a.out1 = dx1; a.add = add0; a.add1 = add1;
  a.gamma = gamma; a.beta = beta; a.scale = scale; a.shift = shift; a.stats = const_cast<float*>(stats); a.bstats = s.bstats;
  a.partials = s.partials; a.tickets = s.tickets; a.C = channels; a.G = groups; a.HW = hw; a.eps = 0.f;
  a.film_stride = film_stride > 0 ? film_stride : channels;
  return dispatch<1>(mode, NHMC_LAYOUT_NHWC, dtype, a, s, batch, (cudaStream_t)stream);
}

extern "C" size_t nhmc_gn_workspace_bytes(int max_batch) {
  // room for 8 accumulator replicas per sample (samples spanning many CTAs), never below 64 slots
  size_t nb = (size_t)(max_batch > 0 ? max_batch : 1) * 8;
  if (nb < 64) nb = 64;
  return kGnWsHeader + kGnWsPerSample * nb;
}

extern "C" int nhmc_gn_check(const void* workspace, size_t workspace_bytes, int32_t* flag_out, nhmc_stream_t stream) {
  // Copies the workspace's timeout flag (0 = no barrier wait of the resident GroupNorm kernels ever timed out) to
  // *flag_out (device or pinned host memory) on `stream`; nothing synchronises.
  NHMC_CHECK_ARG(workspace && flag_out && workspace_bytes >= kGnWsHeader, "workspace and flag_out required");
  const unsigned int* res = reinterpret_cast<const unsigned int*>(workspace);
  return cuda_status(cudaMemcpyAsync(flag_out, res + 1, sizeof(int32_t), cudaMemcpyDefault, (cudaStream_t)stream), "gn check");
}
