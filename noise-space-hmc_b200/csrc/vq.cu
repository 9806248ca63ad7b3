// Nearest-codebook lookup of the VQ-f4 first stage (hmc_latent path): replaces the [N, K] distance matrix +
// argmin + embedding gather of taming's VectorQuantizer2.forward (used by VQModelInterface.decode,
// ldm/models/autoencoder.py:274-282) with one pass: a thread owns one latent pixel, the codebook streams through
// shared memory.  d(k) = (sum z^2 + sum e_k^2) - 2 z.e_k in fp32, first minimum wins (torch.argmin).
// N = B*H*W pixels x K codes x D dims: 32,768 x 8,192 x 3 at FFHQ-256 -- 1.6 GFLOP instead of a 1 GB matrix.
#include "common.cuh"

namespace nhmc {

constexpr int kVqMaxD = 16;
constexpr int kVqTile = 1024;   // codes per shared-memory tile

template <int D>
__global__ void __launch_bounds__(256)
vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ codebook, float* __restrict__ zq,
                  int32_t* __restrict__ idx_out, int64_t P, int K) {
  __shared__ float e_s[kVqTile * D];
  __shared__ float ee_s[kVqTile];
  const int b = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < P;
  float zv[D];
  float zz = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    zv[d] = live ? z[((int64_t)b * D + d) * P + p] : 0.f;
    zz = fmaf(zv[d], zv[d], zz);
  }
  float best = INFINITY;
  int best_k = 0;
  for (int k0 = 0; k0 < K; k0 += kVqTile) {
    const int nk = K - k0 < kVqTile ? K - k0 : kVqTile;
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) e_s[i] = codebook[(int64_t)k0 * D + i];
    __syncthreads();
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(e_s[i * D + d], e_s[i * D + d], s);
      ee_s[i] = s;
    }
    __syncthreads();
    for (int i = 0; i < nk; ++i) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) dot = fmaf(zv[d], e_s[i * D + d], dot);
      const float dist = __fsub_rn(__fadd_rn(zz, ee_s[i]), 2.0f * dot);
      if (dist < best) { best = dist; best_k = k0 + i; }
    }
  }
  if (live) {
    if (idx_out) idx_out[(int64_t)b * P + p] = best_k;
#pragma unroll
    for (int d = 0; d < D; ++d) {   // the reference returns z + (e - z).detach(): the same value up to fp32 rounding, reproduced here
      const float e = __ldg(codebook + (int64_t)best_k * D + d);
      zq[((int64_t)b * D + d) * P + p] = __fadd_rn(zv[d], __fsub_rn(e, zv[d]));
    }
  }
}

}  // namespace nhmc

using namespace nhmc;

extern "C" int nhmc_vq_nearest(const float* z, const float* codebook, float* zq, int32_t* idx_out, int batch, int dim,
                               int64_t pixels, int n_codes, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(z && codebook && zq && batch > 0 && pixels > 0 && n_codes > 0, "bad arguments");
  NHMC_CHECK_ARG(dim >= 1 && dim <= kVqMaxD, "embedding dimension out of range");
  const dim3 grid((unsigned)((pixels + 255) / 256), (unsigned)batch);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dim) {
#define VQ_CASE(D) case D: vq_nearest_kernel<D><<<grid, 256, 0, st>>>(z, codebook, zq, idx_out, pixels, n_codes); break;
    VQ_CASE(1) VQ_CASE(2) VQ_CASE(3) VQ_CASE(4) VQ_CASE(8)
#undef VQ_CASE
    default:
      set_error("vq_nearest: embedding dimension %d not instantiated (1, 2, 3, 4, 8)", dim);
      return NHMC_ERR_UNSUPPORTED;
  }
  return NHMC_LAUNCH_CHECK();
}
