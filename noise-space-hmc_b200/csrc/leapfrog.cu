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

// Diagonal mass (hmc_test_conditioning, main_sampling.py:776-894): per-element inverse mass, kinetic energy
// sum(inv_m * p^2), drift x + (eps * p) * inv_m, and the trajectory's Welford mean / M2 of the positions (:845-849)
// updated in the same pass (MIDDLE / LAST: the position the gradient was evaluated at is x_in).
struct LfDiag {
  const float* inv_mass;   // [B, N]
  float* wf_mean;          // [B, N] or nullptr
  float* wf_m2;            // [B, N]
};

template <int MODE, bool HAS_GB, bool DIAG>
__global__ void __launch_bounds__(kThreads)
leapfrog_kernel(const float* __restrict__ x_in, float* x_out, float* p, const float* __restrict__ g_a,
                const float* __restrict__ g_b, const double* __restrict__ loss_in, nhmc_chain* chains,
                double mass, int64_t n_vec, int64_t n_per_chain, double* partials, unsigned int* tickets, LfDiag dg) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const double ch_sigma_y = chains[b].sigma_y, ch_eps = chains[b].epsilon;
  LfCoef c;
  c.k = (float)(1.0 / (2.0 * ch_sigma_y * ch_sigma_y));
  c.kick = (float)(MODE == NHMC_LF_FIRST ? ch_eps / 2.0 : ch_eps);
  c.half = (float)(ch_eps / 2.0);
  c.drift = DIAG ? (float)ch_eps : (float)(ch_eps * (1.0 / mass));
  // Welford: on from epoch - burn > epochs // 3 (main_sampling.py:845), step index l = chains[b].reserved (0-based)
  bool wf = false; float wf_n = 1.0f;
  if (DIAG && MODE != NHMC_LF_FIRST && dg.wf_mean) {
    wf = (chains[b].epoch - 5) > 13;
    wf_n = (float)(chains[b].reserved + 1);
  }

  const int64_t base = (int64_t)b * n_per_chain;
  const float4* xi = reinterpret_cast<const float4*>(x_in + base);
  float4* xo = reinterpret_cast<float4*>(x_out + base);
  float4* pp = reinterpret_cast<float4*>(p + base);
  const float4* ga = reinterpret_cast<const float4*>(g_a + base);
  const float4* gb = HAS_GB ? reinterpret_cast<const float4*>(g_b + base) : nullptr;
  const float4* im = DIAG ? reinterpret_cast<const float4*>(dg.inv_mass + base) : nullptr;
  float4* wm = (DIAG && wf) ? reinterpret_cast<float4*>(dg.wf_mean + base) : nullptr;
  float4* w2 = (DIAG && wf) ? reinterpret_cast<float4*>(dg.wf_m2 + base) : nullptr;

  double sx = 0.0, sp = 0.0;
  constexpr int U = 4;
  const int64_t stride = (int64_t)nblk * kThreads;
  for (int64_t i0 = (int64_t)blk * kThreads + threadIdx.x; i0 < n_vec; i0 += stride * U) {
    float4 vx[U], vp[U], vg[U], vm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n_vec) {
        vx[u] = ld_stream(xi + i);
        vp[u] = ld_stream(pp + i);
        vg[u] = ld_ro(ga + i);
        if (DIAG) vm[u] = ld_ro(im + i);
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
      float ms[4] = {1.f, 1.f, 1.f, 1.f};
      if (DIAG) { ms[0] = vm[u].x; ms[1] = vm[u].y; ms[2] = vm[u].z; ms[3] = vm[u].w; }
      if (DIAG && wf) {   // running mean / M2 of the trajectory's positions, fp32 op by op (main_sampling.py:846-849)
        float4 mv = ld_stream(wm + i), m2 = ld_stream(w2 + i);
        float me[4] = {mv.x, mv.y, mv.z, mv.w}, qe[4] = {m2.x, m2.y, m2.z, m2.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float delta = __fsub_rn(xs[e], me[e]);
          me[e] = __fadd_rn(me[e], __fdiv_rn(delta, wf_n));
          qe[e] = __fadd_rn(qe[e], __fmul_rn(delta, __fsub_rn(xs[e], me[e])));
        }
        st_stream(wm + i, make_float4(me[0], me[1], me[2], me[3]));
        st_stream(w2 + i, make_float4(qe[0], qe[1], qe[2], qe[3]));
      }
      // energy sums: products formed exactly in fp64 from the fp32 values, four terms summed per vector before they
      // join the running sums (one dependent DADD per vector instead of four)
      double ex = 0.0, ep = 0.0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (MODE == NHMC_LF_FIRST) {  // energies of the state the momentum was drawn at
          ex += (double)xs[e] * (double)xs[e];
          ep += DIAG ? (double)__fmul_rn(ms[e], __fmul_rn(ps[e], ps[e])) : (double)ps[e] * (double)ps[e];
        }
        const float F = force1(xs[e], gs[e], c.k);
        ps[e] = __fsub_rn(ps[e], __fmul_rn(c.kick, F));
        if (MODE == NHMC_LF_LAST) {
          ps[e] = __fadd_rn(ps[e], __fmul_rn(c.half, F));
          ex += (double)xs[e] * (double)xs[e];
          ep += DIAG ? (double)__fmul_rn(ms[e], __fmul_rn(ps[e], ps[e])) : (double)ps[e] * (double)ps[e];
        } else {
          xs[e] = DIAG ? __fadd_rn(xs[e], __fmul_rn(__fmul_rn(c.drift, ps[e]), ms[e]))
                       : __fadd_rn(xs[e], __fmul_rn(c.drift, ps[e]));
        }
      }
      sx += ex; sp += ep;
      st_stream(pp + i, make_float4(ps[0], ps[1], ps[2], ps[3]));
      if (MODE != NHMC_LF_LAST) st_stream(xo + i, make_float4(xs[0], xs[1], xs[2], xs[3]));
    }
  }

  if (MODE != NHMC_LF_MIDDLE) {
    __shared__ double red[2 * 32];
    double v[2] = {sx, sp};
    block_sum<double, 2>(v, red);
    if (grid_finish_par<2>(v, partials + (int64_t)b * kMaxBlocksPerChain * 2, tickets + b, blk, nblk, red)) {
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
  } else if (sched == NHMC_SCHED_COND) {
    // main_sampling.py:809-817 (burn = 5, epochs = 40): python doubles, every operation rounded separately
    const int burn = 5, epochs = 40;
    if (c.epoch < burn) {
      c.sigma_y = __dadd_rn(c.sigma_0, 0.9);
    } else if (c.epoch < epochs) {
      const double f = 1.0 - (double)(c.epoch - burn) / (double)epochs;
      c.sigma_y = __dadd_rn(c.sigma_0, __dmul_rn(0.9, pow(f, 3.0)));
    } else if (c.epoch == epochs) {
      c.sigma_y = c.sigma_0;
      if (c.tau > 0.1) { c.tau = 0.1; c.epsilon = 0.01; }
    }
  }
  c.reserved = 0;      // leapfrog step index l of the trajectory that starts now (diagonal-mass Welford update)
  c.iters += 1;
}

__global__ void advance_step_kernel(nhmc_chain* chains, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n) chains[b].reserved += 1;
}

}  // namespace nhmc

using namespace nhmc;

extern "C" size_t nhmc_leapfrog_workspace_bytes(int n_chains) {
  if (n_chains < 1) n_chains = 1;
  return (size_t)n_chains * kMaxBlocksPerChain * 2 * sizeof(double) + (size_t)n_chains * sizeof(unsigned int) + 64;
}

extern "C" int nhmc_begin_iteration(nhmc_chain* chains, int n_chains, int sched, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(chains && n_chains > 0, "bad chain array");
  NHMC_CHECK_ARG(sched == NHMC_SCHED_PIXEL || sched == NHMC_SCHED_LATENT || sched == NHMC_SCHED_COND, "unknown schedule");
  begin_iteration_kernel<<<(n_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chains, n_chains, sched);
  return NHMC_LAUNCH_CHECK();
}

static int leapfrog_launch(int mode, const float* x_in, float* x_out, float* p, const float* g_a, const float* g_b,
                           const double* loss_in, nhmc_chain* chains, double mass, int n_chains, int64_t n_per_chain,
                           void* workspace, const LfDiag* dg, cudaStream_t st) {
  NHMC_CHECK_ARG(mode >= NHMC_LF_FIRST && mode <= NHMC_LF_LAST, "unknown mode");
  NHMC_CHECK_ARG(x_in && p && g_a && chains && workspace, "null pointer");
  NHMC_CHECK_ARG(mode == NHMC_LF_LAST || x_out, "x_out required");
  NHMC_CHECK_ARG(n_chains > 0 && n_per_chain > 0 && mass > 0.0, "bad sizes");
  double* partials = reinterpret_cast<double*>(workspace);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(partials + (size_t)n_chains * kMaxBlocksPerChain * 2);
  const bool vec = (n_per_chain % 4 == 0) && aligned16(x_in) && aligned16(p) && aligned16(g_a) &&
                   (!g_b || aligned16(g_b)) && (mode == NHMC_LF_LAST || aligned16(x_out)) &&
                   (!dg || (aligned16(dg->inv_mass) && (!dg->wf_mean || (aligned16(dg->wf_mean) && aligned16(dg->wf_m2)))));
  NHMC_CHECK_ARG(vec || !dg, "the diagonal-mass kernel needs n_per_chain % 4 == 0 and 16-byte aligned buffers");
  if (vec) {
    const int64_t n_vec = n_per_chain / 4;
    int nblk = blocks_per_chain(n_vec, n_chains, 4);
    if (mode != NHMC_LF_MIDDLE) {
      // modes that end in a cross-block reduction: ONE resident wave (8 CTAs of 256 threads per SM), so that the
      // per-CTA block reduce + ticket is amortised over >= ~10 vectors per thread instead of ~2.5
      int64_t cap = (int64_t)kSMs * 8 / n_chains;
      if (cap < 1) cap = 1;
      if (nblk > cap) nblk = (int)cap;
    }
    dim3 grid(nblk, n_chains);
    const LfDiag d0 = dg ? *dg : LfDiag{nullptr, nullptr, nullptr};
#define LF_LAUNCH(M, GB, DG)                                                                                     \
  leapfrog_kernel<M, GB, DG><<<grid, kThreads, 0, st>>>(x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_vec, \
                                                        n_per_chain, partials, tickets, d0)
#define LF_MODE(M)                                                              \
  do {                                                                          \
    if (dg) { if (g_b) LF_LAUNCH(M, true, true); else LF_LAUNCH(M, false, true); } \
    else    { if (g_b) LF_LAUNCH(M, true, false); else LF_LAUNCH(M, false, false); } \
  } while (0)
    if (mode == NHMC_LF_FIRST) LF_MODE(NHMC_LF_FIRST);
    else if (mode == NHMC_LF_MIDDLE) LF_MODE(NHMC_LF_MIDDLE);
    else LF_MODE(NHMC_LF_LAST);
#undef LF_MODE
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

extern "C" int nhmc_leapfrog(int mode, const float* x_in, float* x_out, float* p, const float* g_a,
                             const float* g_b, const double* loss_in, nhmc_chain* chains, double mass,
                             int n_chains, int64_t n_per_chain, void* workspace, nhmc_stream_t stream) {
  return leapfrog_launch(mode, x_in, x_out, p, g_a, g_b, loss_in, chains, mass, n_chains, n_per_chain, workspace, nullptr,
                         (cudaStream_t)stream);
}

extern "C" int nhmc_leapfrog_diag(int mode, const float* x_in, float* x_out, float* p, const float* g_a, const float* g_b,
                                  const double* loss_in, nhmc_chain* chains, const float* inv_mass, float* wf_mean,
                                  float* wf_m2, int n_chains, int64_t n_per_chain, void* workspace, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(inv_mass && ((wf_mean == nullptr) == (wf_m2 == nullptr)), "inv_mass required; wf_mean and wf_m2 come together");
  const LfDiag dg{inv_mass, wf_mean, wf_m2};
  if (int e = leapfrog_launch(mode, x_in, x_out, p, g_a, g_b, loss_in, chains, 1.0, n_chains, n_per_chain, workspace, &dg,
                              (cudaStream_t)stream)) return e;
  if (mode != NHMC_LF_FIRST) {   // the trajectory's step index advances after every MIDDLE / LAST step
    advance_step_kernel<<<(n_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chains, n_chains);
    return NHMC_LAUNCH_CHECK();
  }
  return NHMC_OK;
}
