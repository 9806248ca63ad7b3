// K1/K6: fused leapfrog position/momentum update with in-register energy accumulation.
// Replaces the elementwise ATen launches of main_sampling.py:697,702,706-707,713,715,717
// (hmc) and main_sampling_latent.py:654,663-676 (hmc_latent).
//
// HBM-bound: 3 reads (x, p, g[, g_b]) + 2 writes (x', p) of N fp32 per chain = 20 B/element
// (FIRST/MIDDLE), 3 reads + 1 write for LAST.  One pass, float4 accesses, 4 vectors in flight
// per thread, grid = (blocks_per_chain, chains).
#include "common.cuh"

namespace nhmc {

struct LfCoef {
  float k;     // 1/(2 sigma_y^2)
  float kick;  // eps/2 (FIRST) or eps (MIDDLE/LAST)
  float half;  // eps/2 (LAST: undo half of the last kick)
  float drift; // eps * m^-1
};

// p - kick*(x + k*g), each product/sum rounded separately like the reference's eager ops.
__device__ __forceinline__ float force1(float x, float g, float k) { return __fadd_rn(x, __fmul_rn(k, g)); }

template <int MODE, bool HAS_GB>
__global__ void __launch_bounds__(kThreads)
leapfrog_kernel(const float* __restrict__ x_in, float* x_out, float* p, const float* __restrict__ g_a,
                const float* __restrict__ g_b, const double* __restrict__ loss_in, nhmc_chain* chains,
                double mass, int64_t n_vec, int64_t n_per_chain, double* partials, unsigned int* tickets) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const nhmc_chain ch = chains[b];
  LfCoef c;
  c.k = (float)(1.0 / (2.0 * ch.sigma_y * ch.sigma_y));
  c.kick = (float)(MODE == NHMC_LF_FIRST ? ch.epsilon / 2.0 : ch.epsilon);
  c.half = (float)(ch.epsilon / 2.0);
  c.drift = (float)(ch.epsilon * (1.0 / mass));

  const int64_t base = (int64_t)b * n_per_chain;
  const float4* xi = reinterpret_cast<const float4*>(x_in + base);
  float4* xo = reinterpret_cast<float4*>(x_out + base);
  float4* pp = reinterpret_cast<float4*>(p + base);
  const float4* ga = reinterpret_cast<const float4*>(g_a + base);
  const float4* gb = HAS_GB ? reinterpret_cast<const float4*>(g_b + base) : nullptr;

  double sx = 0.0, sp = 0.0;
  constexpr int U = 4;
  const int64_t stride = (int64_t)nblk * kThreads;
  for (int64_t i0 = (int64_t)blk * kThreads + threadIdx.x; i0 < n_vec; i0 += stride * U) {
    float4 vx[U], vp[U], vg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n_vec) {
        vx[u] = ld_stream(xi + i);
        vp[u] = ld_stream(pp + i);
        vg[u] = ld_ro(ga + i);
        if (HAS_GB) {
          float4 t = ld_ro(gb + i);
          vg[u].x = __fadd_rn(vg[u].x, t.x); vg[u].y = __fadd_rn(vg[u].y, t.y);
          vg[u].z = __fadd_rn(vg[u].z, t.z); vg[u].w = __fadd_rn(vg[u].w, t.w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n_vec) continue;
      float xs[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
      float ps[4] = {vp[u].x, vp[u].y, vp[u].z, vp[u].w};
      const float gs[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (MODE == NHMC_LF_FIRST) {  // energies of the state the momentum was drawn at
          sx += (double)xs[e] * (double)xs[e];
          sp += (double)ps[e] * (double)ps[e];
        }
        const float F = force1(xs[e], gs[e], c.k);
        ps[e] = __fsub_rn(ps[e], __fmul_rn(c.kick, F));
        if (MODE == NHMC_LF_LAST) {
          ps[e] = __fadd_rn(ps[e], __fmul_rn(c.half, F));
          sx += (double)xs[e] * (double)xs[e];
          sp += (double)ps[e] * (double)ps[e];
        } else {
          xs[e] = __fadd_rn(xs[e], __fmul_rn(c.drift, ps[e]));
        }
      }
      st_stream(pp + i, make_float4(ps[0], ps[1], ps[2], ps[3]));
      if (MODE != NHMC_LF_LAST) st_stream(xo + i, make_float4(xs[0], xs[1], xs[2], xs[3]));
    }
  }

  if (MODE != NHMC_LF_MIDDLE) {
    __shared__ double red[2 * 32];
    double v[2] = {sx, sp};
    block_sum<double, 2>(v, red);
    if (grid_finish<2>(v, partials + (int64_t)b * kMaxBlocksPerChain * 2, tickets + b, blk, nblk)) {
      const int w = (MODE == NHMC_LF_FIRST) ? 0 : 1;
      chains[b].sum_x2[w] = v[0];
      chains[b].sum_p2[w] = v[1];
      if (loss_in) chains[b].loss[w] = loss_in[b];
    }
  }
}

// scalar fallback for n_per_chain % 4 != 0 or unaligned pointers (never hit at FFHQ-256 / latent sizes)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
leapfrog_scalar_kernel(const float* x_in, float* x_out, float* p, const float* g_a, const float* g_b,
                       const double* loss_in, nhmc_chain* chains, double mass, int64_t n_per_chain,
                       double* partials, unsigned int* tickets) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const nhmc_chain ch = chains[b];
  const float k = (float)(1.0 / (2.0 * ch.sigma_y * ch.sigma_y));
  const float kick = (float)(MODE == NHMC_LF_FIRST ? ch.epsilon / 2.0 : ch.epsilon);
  const float half = (float)(ch.epsilon / 2.0);
  const float drift = (float)(ch.epsilon * (1.0 / mass));
  const int64_t base = (int64_t)b * n_per_chain;
  double sx = 0.0, sp = 0.0;
  for (int64_t i = (int64_t)blk * kThreads + threadIdx.x; i < n_per_chain; i += (int64_t)nblk * kThreads) {
    float x = x_in[base + i], pv = p[base + i];
    float g = g_a[base + i];
    if (g_b) g = __fadd_rn(g, g_b[base + i]);
    if (MODE == NHMC_LF_FIRST) { sx += (double)x * x; sp += (double)pv * pv; }
    const float F = force1(x, g, k);
    pv = __fsub_rn(pv, __fmul_rn(kick, F));
    if (MODE == NHMC_LF_LAST) {
      pv = __fadd_rn(pv, __fmul_rn(half, F));
      sx += (double)x * x; sp += (double)pv * pv;
    } else {
      x_out[base + i] = __fadd_rn(x, __fmul_rn(drift, pv));
    }
    p[base + i] = pv;
  }
  if (MODE != NHMC_LF_MIDDLE) {
    __shared__ double red[2 * 32];
    double v[2] = {sx, sp};
    block_sum<double, 2>(v, red);
    if (grid_finish<2>(v, partials + (int64_t)b * kMaxBlocksPerChain * 2, tickets + b, blk, nblk)) {
      const int w = (MODE == NHMC_LF_FIRST) ? 0 : 1;
      chains[b].sum_x2[w] = v[0];
      chains[b].sum_p2[w] = v[1];
      if (loss_in) chains[b].loss[w] = loss_in[b];
    }
  }
}

__global__ void begin_iteration_kernel(nhmc_chain* chains, int n, int sched) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  nhmc_chain& c = chains[b];
  if (!c.active) return;
  if (sched == NHMC_SCHED_PIXEL) {
    // main_sampling.py:683-689 (epochs = 60, anneal constant 1.6)
    const int epochs = 60;
    if (c.epoch < epochs) {
      const double f = 1.0 - (double)c.epoch / (double)epochs;
      // python: sigma_0 + 1.6 * (1 - epoch/epochs) ** 2, every operation rounded separately
      c.sigma_y = __dadd_rn(c.sigma_0, __dmul_rn(1.6, __dmul_rn(f, f)));
    } else if (c.epoch == epochs) {
      c.sigma_y = c.sigma_0;
      if (c.tau > 0.1) { c.tau = 0.1; c.epsilon = 0.01; }
    }
  }
  c.iters += 1;
}

}  // namespace nhmc

using namespace nhmc;

extern "C" size_t nhmc_leapfrog_workspace_bytes(int n_chains) {
  if (n_chains < 1) n_chains = 1;
  return (size_t)n_chains * kMaxBlocksPerChain * 2 * sizeof(double) + (size_t)n_chains * sizeof(unsigned int) + 64;
}

extern "C" int nhmc_begin_iteration(nhmc_chain* chains, int n_chains, int sched, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(chains && n_chains > 0, "bad chain array");
  NHMC_CHECK_ARG(sched == NHMC_SCHED_PIXEL || sched == NHMC_SCHED_LATENT, "unknown schedule");
  begin_iteration_kernel<<<(n_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chains, n_chains, sched);
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_leapfrog(int mode, const float* x_in, float* x_out, float* p, const float* g_a,
                             const float* g_b, const double* loss_in, nhmc_chain* chains, double mass,
                             int n_chains, int64_t n_per_chain, void* workspace, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(mode >= NHMC_LF_FIRST && mode <= NHMC_LF_LAST, "unknown mode");
  NHMC_CHECK_ARG(x_in && p && g_a && chains && workspace, "null pointer");
  NHMC_CHECK_ARG(mode == NHMC_LF_LAST || x_out, "x_out required");
  NHMC_CHECK_ARG(n_chains > 0 && n_per_chain > 0 && mass > 0.0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = reinterpret_cast<double*>(workspace);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(partials + (size_t)n_chains * kMaxBlocksPerChain * 2);
  const bool vec = (n_per_chain % 4 == 0) && aligned16(x_in) && aligned16(p) && aligned16(g_a) &&
                   (!g_b || aligned16(g_b)) && (mode == NHMC_LF_LAST || aligned16(x_out));
  if (vec) {
    const int64_t n_vec = n_per_chain / 4;
    dim3 grid(blocks_per_chain(n_vec, n_chains, 4), n_chains);
#define LF_LAUNCH(M, GB)                                                                                 \
  leapfrog_kernel<M, GB><<<grid, kThreads, 0, st>>>(x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_vec, \
                                                    n_per_chain, partials, tickets)
    if (mode == NHMC_LF_FIRST) { if (g_b) LF_LAUNCH(NHMC_LF_FIRST, true); else LF_LAUNCH(NHMC_LF_FIRST, false); }
    else if (mode == NHMC_LF_MIDDLE) { if (g_b) LF_LAUNCH(NHMC_LF_MIDDLE, true); else LF_LAUNCH(NHMC_LF_MIDDLE, false); }
    else { if (g_b) LF_LAUNCH(NHMC_LF_LAST, true); else LF_LAUNCH(NHMC_LF_LAST, false); }
#undef LF_LAUNCH
  } else {
    dim3 grid(blocks_per_chain(n_per_chain, n_chains, 4), n_chains);
    if (mode == NHMC_LF_FIRST)
      leapfrog_scalar_kernel<NHMC_LF_FIRST><<<grid, kThreads, 0, st>>>(x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_per_chain, partials, tickets);
    else if (mode == NHMC_LF_MIDDLE)
      leapfrog_scalar_kernel<NHMC_LF_MIDDLE><<<grid, kThreads, 0, st>>>(x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_per_chain, partials, tickets);
    else
      leapfrog_scalar_kernel<NHMC_LF_LAST><<<grid, kThreads, 0, st>>>(x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_per_chain, partials, tickets);
  }
  return NHMC_LAUNCH_CHECK();
}
