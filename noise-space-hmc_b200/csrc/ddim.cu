// K2/K3: few-step DDIM epilogue, forward and backward, one vectorised pass each.
// Replaces Unconditional.cal_x0 / map_back (algos/unconditional.py:17-28; x0-prediction formula
// of functions/denoising.py:62) and their autograd graph (about 12 + 15 ATen launches per step).
//
// fwd : reads xt (fp32) + eps (net dtype), writes xt' (fp32) + 1 bit/element clip mask.
// bwd : reads g' (= g_a + g_b, fp32) + the mask, writes g_xt (fp32) + g_eps (net dtype).
// A warp owns 128 consecutive elements per step (float4 per lane); the clip mask of that group is
// four ballots -> 16 bytes, so the backward pass never re-reads xt/eps.
#include "common.cuh"

namespace nhmc {

template <typename TE>
__global__ void __launch_bounds__(kThreads)
ddim_fwd_kernel(const float* __restrict__ xt, const TE* __restrict__ eps, int64_t eps_stride,
                float* __restrict__ out, uint32_t* __restrict__ mask, float c1, float sa, float c2, float ss,
                int64_t n_groups, int64_t n_per_chain, int64_t mask_words) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  const float* xb = xt + (int64_t)b * n_per_chain;
  const TE* eb = eps + (int64_t)b * eps_stride;
  float* ob = out + (int64_t)b * n_per_chain;
  uint32_t* mb = mask ? mask + (int64_t)b * mask_words : nullptr;
  for (int64_t gidx = warp0; gidx < n_groups; gidx += nwarps) {
    const int64_t e0 = gidx * 128 + lane * 4;
    const bool ok = e0 < n_per_chain;  // n_per_chain % 4 == 0 on this path
    float4 x = ok ? ld_ro(reinterpret_cast<const float4*>(xb + e0)) : make_float4(0, 0, 0, 0);
    float4 e = ok ? load4<TE>(eb + e0) : make_float4(0, 0, 0, 0);
    float xs[4] = {x.x, x.y, x.z, x.w};
    const float es[4] = {e.x, e.y, e.z, e.w};
    uint32_t bits[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // (xt - eps*c1)/sa ; clip ; ss*x0 + c2*eps   -- reference operation order, no contraction
      const float x0raw = __fdiv_rn(__fsub_rn(xs[j], __fmul_rn(es[j], c1)), sa);
      const bool in = (x0raw >= -1.0f) && (x0raw <= 1.0f);
      const float x0 = fminf(fmaxf(x0raw, -1.0f), 1.0f);
      xs[j] = __fadd_rn(__fmul_rn(ss, x0), __fmul_rn(c2, es[j]));
      bits[j] = __ballot_sync(0xffffffffu, ok && in);
    }
    if (ok) st_stream(reinterpret_cast<float4*>(ob + e0), make_float4(xs[0], xs[1], xs[2], xs[3]));
    if (mb && lane < 4) {
      const uint32_t w = lane == 0 ? bits[0] : lane == 1 ? bits[1] : lane == 2 ? bits[2] : bits[3];
      mb[gidx * 4 + lane] = w;
    }
  }
}

template <typename TE, bool HAS_GB, bool HAS_GE>
__global__ void __launch_bounds__(kThreads)
ddim_bwd_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b, const uint32_t* __restrict__ mask,
                float* __restrict__ g_xt, TE* __restrict__ g_eps, int64_t eps_stride, float c1, float sa, float c2,
                float ss, int64_t n_groups, int64_t n_per_chain, int64_t mask_words) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  const float* ga = g_a + (int64_t)b * n_per_chain;
  const float* gb = HAS_GB ? g_b + (int64_t)b * n_per_chain : nullptr;
  const uint32_t* mb = mask + (int64_t)b * mask_words;
  float* ox = g_xt + (int64_t)b * n_per_chain;
  TE* oe = HAS_GE ? g_eps + (int64_t)b * eps_stride : nullptr;
  for (int64_t gidx = warp0; gidx < n_groups; gidx += nwarps) {
    const int64_t e0 = gidx * 128 + lane * 4;
    if (e0 >= n_per_chain) continue;
    const uint4 mw = __ldg(reinterpret_cast<const uint4*>(mb + gidx * 4));  // same 16 B for the whole warp
    float4 g = ld_ro(reinterpret_cast<const float4*>(ga + e0));
    if (HAS_GB) {
      const float4 t = ld_ro(reinterpret_cast<const float4*>(gb + e0));
      g.x = __fadd_rn(g.x, t.x); g.y = __fadd_rn(g.y, t.y); g.z = __fadd_rn(g.z, t.z); g.w = __fadd_rn(g.w, t.w);
    }
    const float gs[4] = {g.x, g.y, g.z, g.w};
    const uint32_t ms[4] = {mw.x, mw.y, mw.z, mw.w};
    float gx[4], ge[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float m = ((ms[j] >> lane) & 1u) ? 1.0f : 0.0f;
      // autograd order: d(ss*x0) -> clamp mask -> /sa ; eps gets c2*g' and -c1 * that
      const float gx0 = __fmul_rn(__fmul_rn(ss, gs[j]), m);
      gx[j] = __fdiv_rn(gx0, sa);
      ge[j] = __fsub_rn(__fmul_rn(c2, gs[j]), __fmul_rn(gx[j], c1));
    }
    st_stream(reinterpret_cast<float4*>(ox + e0), make_float4(gx[0], gx[1], gx[2], gx[3]));
    if (HAS_GE) store4<TE>(oe + e0, make_float4(ge[0], ge[1], ge[2], ge[3]));
  }
}

// Latent path (main_sampling_latent.py:651,670): the clip of the DDIM output in front of the decoder and its gradient
// mask behind the decoder's VJP, one small kernel each instead of clamp / copy / two compares / and / cast / mul.
__global__ void clip_fwd_kernel(const float* __restrict__ z, float* __restrict__ zc, float* __restrict__ zc2, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(z[i], -1.0f), 1.0f);
    zc[i] = v;
    if (zc2) zc2[i] = v;
  }
}
template <typename TG>
__global__ void clip_bwd_kernel(const TG* __restrict__ g, const float* __restrict__ z, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float zv = z[i];
    out[i] = (zv >= -1.0f && zv <= 1.0f) ? Elem<TG>::to_f(g[i]) : 0.0f;
  }
}

static inline dim3 ddim_grid(int64_t n_groups, int n_chains) {
  // one warp per 128-element group and step; 8 warps per CTA, 2 groups per warp
  return dim3(cta_count(n_groups, 16, n_chains), n_chains);
}

}  // namespace nhmc

using namespace nhmc;

extern "C" size_t nhmc_ddim_mask_bytes(int64_t n_per_chain) { return (size_t)((n_per_chain + 127) / 128) * 16; }

extern "C" int nhmc_ddim_fwd(const float* xt, const void* eps, int eps_dtype, int64_t eps_chain_stride,
                             float* xt_next, uint32_t* mask_out, float c1, float sa, float c2, float ss,
                             int n_chains, int64_t n_per_chain, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(xt && eps && xt_next, "null pointer");
  NHMC_CHECK_ARG(n_chains > 0 && n_per_chain > 0 && n_per_chain % 4 == 0, "n_per_chain must be a positive multiple of 4");
  NHMC_CHECK_ARG(eps_chain_stride >= n_per_chain && eps_chain_stride % 4 == 0, "bad eps stride");
  NHMC_CHECK_ARG(aligned16(xt) && aligned16(xt_next) && aligned16(eps) && (!mask_out || aligned16(mask_out)),
                 "pointers must be 16-byte aligned");
  NHMC_CHECK_ARG(sa > 0.f, "sqrt(abar_t) must be positive");
  const int64_t n_groups = (n_per_chain + 127) / 128;
  const int64_t mask_words = n_groups * 4;
  const dim3 grid = ddim_grid(n_groups, n_chains);
  cudaStream_t st = (cudaStream_t)stream;
  switch (eps_dtype) {
    case NHMC_F32:
      ddim_fwd_kernel<float><<<grid, kThreads, 0, st>>>(xt, (const float*)eps, eps_chain_stride, xt_next, mask_out,
                                                        c1, sa, c2, ss, n_groups, n_per_chain, mask_words);
      break;
    case NHMC_BF16:
      ddim_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(xt, (const __nv_bfloat16*)eps, eps_chain_stride,
                                                                xt_next, mask_out, c1, sa, c2, ss, n_groups,
                                                                n_per_chain, mask_words);
      break;
    case NHMC_F16:
      ddim_fwd_kernel<__half><<<grid, kThreads, 0, st>>>(xt, (const __half*)eps, eps_chain_stride, xt_next, mask_out,
                                                         c1, sa, c2, ss, n_groups, n_per_chain, mask_words);
      break;
    default:
      set_error("nhmc_ddim_fwd: unknown eps dtype %d", eps_dtype);
      return NHMC_ERR_ARG;
  }
  return NHMC_LAUNCH_CHECK();
}

template <typename TE>
static int launch_bwd(const float* g_a, const float* g_b, const uint32_t* mask, float* g_xt, void* g_eps,
                      int64_t stride, float c1, float sa, float c2, float ss, int n_chains, int64_t n, cudaStream_t st) {
  const int64_t n_groups = (n + 127) / 128;
  const int64_t mw = n_groups * 4;
  const dim3 grid = ddim_grid(n_groups, n_chains);
  TE* ge = reinterpret_cast<TE*>(g_eps);
#define BW(GB, GE) ddim_bwd_kernel<TE, GB, GE><<<grid, kThreads, 0, st>>>(g_a, g_b, mask, g_xt, ge, stride, c1, sa, c2, ss, n_groups, n, mw)
  if (g_b) { if (g_eps) BW(true, true); else BW(true, false); }
  else     { if (g_eps) BW(false, true); else BW(false, false); }
#undef BW
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_ddim_bwd(const float* g_a, const float* g_b, const uint32_t* mask, float* g_xt, void* g_eps_out,
                             int eps_dtype, int64_t eps_chain_stride, float c1, float sa, float c2, float ss,
                             int n_chains, int64_t n_per_chain, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(g_a && mask && g_xt, "null pointer");
  NHMC_CHECK_ARG(n_chains > 0 && n_per_chain > 0 && n_per_chain % 4 == 0, "n_per_chain must be a positive multiple of 4");
  NHMC_CHECK_ARG(!g_eps_out || (eps_chain_stride >= n_per_chain && eps_chain_stride % 4 == 0), "bad eps stride");
  NHMC_CHECK_ARG(aligned16(g_a) && aligned16(g_xt) && aligned16(mask) && (!g_b || aligned16(g_b)) &&
                     (!g_eps_out || aligned16(g_eps_out)), "pointers must be 16-byte aligned");
  NHMC_CHECK_ARG(sa > 0.f, "sqrt(abar_t) must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  switch (eps_dtype) {
    case NHMC_F32: return launch_bwd<float>(g_a, g_b, mask, g_xt, g_eps_out, eps_chain_stride, c1, sa, c2, ss, n_chains, n_per_chain, st);
    case NHMC_BF16: return launch_bwd<__nv_bfloat16>(g_a, g_b, mask, g_xt, g_eps_out, eps_chain_stride, c1, sa, c2, ss, n_chains, n_per_chain, st);
    case NHMC_F16: return launch_bwd<__half>(g_a, g_b, mask, g_xt, g_eps_out, eps_chain_stride, c1, sa, c2, ss, n_chains, n_per_chain, st);
    default: set_error("nhmc_ddim_bwd: unknown eps dtype %d", eps_dtype); return NHMC_ERR_ARG;
  }
}

extern "C" int nhmc_clip_fwd(const float* z, float* zc, float* zc_copy, int64_t n, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(z && zc && n > 0, "bad arguments");
  int64_t blocks = (n + 255) / 256; if (blocks > kSMs * 8) blocks = kSMs * 8;
  clip_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, zc, zc_copy, n);
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_clip_bwd(const void* g, int g_dtype, const float* z, float* out, int64_t n, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(g && z && out && n > 0, "bad arguments");
  int64_t blocks = (n + 255) / 256; if (blocks > kSMs * 8) blocks = kSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  switch (g_dtype) {
    case NHMC_F32: clip_bwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)g, z, out, n); break;
    case NHMC_BF16: clip_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)g, z, out, n); break;
    case NHMC_F16: clip_bwd_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)g, z, out, n); break;
    default: set_error("nhmc_clip_bwd: unknown dtype %d", g_dtype); return NHMC_ERR_ARG;
  }
  return NHMC_LAUNCH_CHECK();
}
