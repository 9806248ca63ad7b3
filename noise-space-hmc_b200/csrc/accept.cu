// K7: Metropolis accept/reject and the annealing schedule, applied on device with no host sync.
// Replaces main_sampling.py:697,717-731,743-749 (pixel) and main_sampling_latent.py:676-733
// (latent): two `.item()` round trips, python-side schedule updates and a clone per iteration.
#include "common.cuh"

namespace nhmc {

// fp32 Hamiltonian in the reference's evaluation order: 0.5*sum(x^2) + k*loss + 0.5*sum(p^2)*m^-1,
// each term rounded to fp32 the way the eager ops round it.
__device__ __forceinline__ float hamiltonian(double sx2, double sp2, double loss, double sigma_y, double mass) {
  const float prior = __fmul_rn(0.5f, (float)sx2);
  const float data = __fmul_rn((float)(1.0 / (2.0 * sigma_y * sigma_y)), (float)loss);
  const float kin = __fmul_rn(__fmul_rn(0.5f, (float)sp2), (float)(1.0 / mass));
  return __fadd_rn(__fadd_rn(prior, data), kin);
}

__global__ void accept_kernel(nhmc_chain* chains, const float* __restrict__ u, double mass, int sched, int n_slots, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  nhmc_chain c = chains[b];
  c.accept = 0; c.collect = 0; c.slot = -1;
  if (!c.active) { chains[b] = c; return; }
  c.H[0] = hamiltonian(c.sum_x2[0], c.sum_p2[0], c.loss[0], c.sigma_y, mass);
  c.H[1] = hamiltonian(c.sum_x2[1], c.sum_p2[1], c.loss[1], c.sigma_y, mass);
  const float dH = __fsub_rn(c.H[1], c.H[0]);
  c.prob = fminf(1.0f, expf(-dH));   // NaN energy -> prob NaN -> reject, like `u < nan` in python
  c.u = u[b];
  const bool acc = c.u < c.prob;
  c.accept = acc ? 1 : 0;
  if (sched == NHMC_SCHED_PIXEL) {
    const int epochs = 60, sampling = 20;
    if (acc) {
      c.accepted += 1;
      c.rejected = 0;
      if (c.epoch >= epochs + sampling) {
        c.collect = 1;
        c.slot = c.n_collected % n_slots;
        c.n_collected += 1;
      }
      c.epoch += 1;
      if (c.epoch >= epochs + 2 * sampling) c.active = 0;
    } else {
      c.rejected += 1;
      if (c.rejected >= 2) { c.tau *= 0.95; c.epsilon *= 0.95; }
    }
  } else if (sched == NHMC_SCHED_COND) {
    // hmc_test_conditioning (main_sampling.py:856-892): burn 5, epochs 40, sampling 10, 85 accepted iterations in all
    const int burn = 5, epochs = 40, sampling = 10;
    if (acc) {
      c.accepted += 1;
      c.rejected = 0;
      if (c.epoch >= epochs + sampling) {
        c.collect = 1;
        c.slot = c.n_collected % n_slots;
        c.n_collected += 1;
      }
      c.epoch += 1;
      if (c.epoch >= burn + epochs + 4 * sampling) c.active = 0;
    } else {
      c.rejected += 1;
      if (c.rejected >= 2) { c.tau *= 0.95; c.epsilon *= 0.95; }
    }
  } else {
    const int epochs = 50, sampling = 10;
    if (acc) {
      c.accepted += 1;
      c.rejected = 0;
      if (c.epoch < epochs) {
        c.sigma_y = c.sigma_y_init * pow(c.sigma_0 / c.sigma_y_init, (double)c.epoch / (double)epochs);
      } else {
        c.sigma_y = c.sigma_0;
        c.tau = 0.1;
        c.epsilon = 0.01;
        if (c.accepted > 1) {  // the previously accepted sample exists (the reference would raise otherwise)
          c.collect = 1;
          c.slot = c.n_collected % n_slots;
          c.n_collected += 1;
        }
      }
    } else {
      c.rejected += 1;
      if (c.rejected >= 2) { c.tau *= 0.9; c.epsilon *= 0.9; c.rejected = 0; }
    }
    c.epoch += 1;  // `for epoch in range(70)`: rejects consume an epoch too
    if (c.epoch >= epochs + 2 * sampling) c.active = 0;
  }
  chains[b] = c;
}

__global__ void __launch_bounds__(kThreads)
commit_kernel(const nhmc_chain* __restrict__ chains, float* __restrict__ x, const float* __restrict__ x_prop,
              const float* __restrict__ x_hat, float* __restrict__ x_keep, float* __restrict__ samples, int sched,
              int n_slots, int64_t n_vec) {
  const int b = blockIdx.y;
  const int acc = chains[b].accept, col = chains[b].collect, slot = chains[b].slot;
  if (!acc) return;  // both schedules only ever store on an accepting iteration
  float4* xb = reinterpret_cast<float4*>(x) + (int64_t)b * n_vec;
  const float4* xp = reinterpret_cast<const float4*>(x_prop) + (int64_t)b * n_vec;
  const float4* xh = reinterpret_cast<const float4*>(x_hat) + (int64_t)b * n_vec;
  float4* xk = x_keep ? reinterpret_cast<float4*>(x_keep) + (int64_t)b * n_vec : nullptr;
  float4* sm = (col && samples) ? reinterpret_cast<float4*>(samples) + ((int64_t)b * n_slots + slot) * n_vec : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * kThreads) {
    st_stream(xb + v, ld_ro(xp + v));
    const float4 h = ld_ro(xh + v);
    if (sched != NHMC_SCHED_LATENT) {
      if (sm) st_stream(sm + v, h);
    } else {
      if (sm) st_stream(sm + v, ld_stream(xk + v));  // one accept behind (main_sampling_latent.py:705-713)
      if (xk) st_stream(xk + v, h);
    }
  }
}

// the gradient evaluation at the top of every iteration (main_sampling.py:693-695) repeats the last
// evaluation of the previous ACCEPTED trajectory (same x, same loss, same gradient).  Keeping (loss, g) of the current
// state per chain makes it 20 instead of 21 evaluations per iteration: accepted chains take the proposal's values.
__global__ void __launch_bounds__(kThreads)
commit_grad_kernel(const nhmc_chain* __restrict__ chains, float* __restrict__ g_state, const float* __restrict__ g_a,
                   const float* __restrict__ g_b, double* __restrict__ loss_state, const double* __restrict__ loss_prop,
                   int64_t n_vec) {
  const int b = blockIdx.y;
  if (!chains[b].accept) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) loss_state[b] = loss_prop[b];
  float4* gs = reinterpret_cast<float4*>(g_state) + (int64_t)b * n_vec;
  const float4* ga = reinterpret_cast<const float4*>(g_a) + (int64_t)b * n_vec;
  const float4* gb = g_b ? reinterpret_cast<const float4*>(g_b) + (int64_t)b * n_vec : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * kThreads) {
    float4 t = ld_ro(ga + v);
    if (gb) {   // same rounding as the leapfrog kernel's g_a + g_b
      const float4 u = ld_ro(gb + v);
      t.x = __fadd_rn(t.x, u.x); t.y = __fadd_rn(t.y, u.y); t.z = __fadd_rn(t.z, u.z); t.w = __fadd_rn(t.w, u.w);
    }
    st_stream(gs + v, t);
  }
}

}  // namespace nhmc

using namespace nhmc;

extern "C" int nhmc_commit_grad(const nhmc_chain* chains, float* g_state, const float* g_a, const float* g_b,
                                double* loss_state, const double* loss_prop, int n_chains, int64_t n_per_chain,
                                nhmc_stream_t stream) {
  NHMC_CHECK_ARG(chains && g_state && g_a && loss_state && loss_prop && n_chains > 0, "bad arguments");
  NHMC_CHECK_ARG(n_per_chain > 0 && n_per_chain % 4 == 0, "n_per_chain must be a positive multiple of 4");
  NHMC_CHECK_ARG(aligned16(g_state) && aligned16(g_a) && (!g_b || aligned16(g_b)), "pointers must be 16-byte aligned");
  const int64_t n_vec = n_per_chain / 4;
  dim3 grid(cta_count(n_vec, kThreads * 4, n_chains), n_chains);
  commit_grad_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(chains, g_state, g_a, g_b, loss_state, loss_prop, n_vec);
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_accept(nhmc_chain* chains, const float* u, double mass, int sched, int n_slots, int n_chains,
                           nhmc_stream_t stream) {
  NHMC_CHECK_ARG(chains && u && n_chains > 0 && n_slots > 0 && mass > 0.0, "bad arguments");
  NHMC_CHECK_ARG(sched == NHMC_SCHED_PIXEL || sched == NHMC_SCHED_LATENT || sched == NHMC_SCHED_COND, "unknown schedule");
  accept_kernel<<<(n_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chains, u, mass, sched, n_slots, n_chains);
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_commit(const nhmc_chain* chains, float* x, const float* x_prop, const float* x_hat, float* x_keep,
                           float* samples, int sched, int n_slots, int n_chains, int64_t n_per_chain,
                           nhmc_stream_t stream) {
  NHMC_CHECK_ARG(chains && x && x_prop && x_hat && n_chains > 0 && n_slots > 0, "bad arguments");
  NHMC_CHECK_ARG(n_per_chain > 0 && n_per_chain % 4 == 0, "n_per_chain must be a positive multiple of 4");
  NHMC_CHECK_ARG(sched != NHMC_SCHED_LATENT || x_keep, "x_keep required for the latent schedule");
  NHMC_CHECK_ARG(aligned16(x) && aligned16(x_prop) && aligned16(x_hat) && (!x_keep || aligned16(x_keep)) &&
                     (!samples || aligned16(samples)), "pointers must be 16-byte aligned");
  const int64_t n_vec = n_per_chain / 4;
  dim3 grid(cta_count(n_vec, kThreads * 4, n_chains), n_chains);
  commit_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(chains, x, x_prop, x_hat, x_keep, samples, sched, n_slots, n_vec);
  return NHMC_LAUNCH_CHECK();
}
