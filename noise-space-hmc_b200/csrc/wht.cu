// Walsh-Hadamard compressive sensing (Hfuncs.py:611-651, --deg cs<ratio>): y[b, k*C + c] = FWHT(x_c)[perm[k]] / D
// for the first D*D/ratio entries of a pixel permutation; H^T = H^+ (all singular values are 1): scatter + FWHT.
//
// The D*D-point natural-order transform factors as H_{D*D} = H_D (x) H_D: a pass along rows (index bits 0..log2 D - 1,
// butterflies h = 1 .. D/2) and a pass along columns (the upper bits), in the SAME stage order and with the same
// (a + b, a - b) pairing as the reference's loop (:613-621), so the fp32 results are bit-identical.  A row lives in the
// registers of one warp (low bits in-register, high bits by shuffle); the column pass stages a [D x 32] tile through
// shared memory so that global accesses stay 128-byte row segments.  HBM-bound: 2 passes x (read + write) per transform.
#include "common.cuh"

namespace nhmc {

// v[j] holds element (lane*EPL + j) [LANE_LOW = false] or (j*32 + lane) [LANE_LOW = true] of an (EPL*32)-point vector
template <int EPL, bool LANE_LOW>
__device__ __forceinline__ void fwht_warp(float (&v)[EPL], int lane) {
  auto in_reg = [&]() {
#pragma unroll
    for (int h = 1; h < EPL; h <<= 1) {
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        if ((j & h) == 0) {
          const float a = v[j], b = v[j | h];
          v[j] = __fadd_rn(a, b);
          v[j | h] = __fsub_rn(a, b);
        }
      }
    }
  };
  auto across = [&]() {
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
      const bool upper = (lane & h) != 0;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const float o = __shfl_xor_sync(0xffffffffu, v[j], h);
        v[j] = upper ? __fsub_rn(o, v[j]) : __fadd_rn(v[j], o);
      }
    }
  };
  if (LANE_LOW) { across(); in_reg(); } else { in_reg(); across(); }   // ascending butterfly span in both layouts
}

// rows pass: one warp per image row; optional clamp of the input to [-1, 1]
template <int EPL>
__global__ void __launch_bounds__(256)
wht_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n_rows, int clamp) {
  constexpr int D = EPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  float v[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    float t = in[row * D + lane * EPL + j];
    if (clamp) t = fminf(fmaxf(t, -1.0f), 1.0f);
    v[j] = t;
  }
  fwht_warp<EPL, false>(v, lane);
#pragma unroll
  for (int j = 0; j < EPL; ++j) out[row * D + lane * EPL + j] = v[j];
}

enum { WHT_EPI_SCALE = 0, WHT_EPI_GRAD = 1 };
// columns pass over a [D x 32] tile of one image; result scaled by 1/D.  WHT_EPI_GRAD: out = -2 * v * 1[-1 <= xhat <= 1]
template <int EPL, int EPI>
__global__ void __launch_bounds__(256)
wht_cols_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ xhat, int no_clamp) {
  constexpr int D = EPL * 32;
  extern __shared__ float tile[];                     // [D][33]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t img = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const float* src = in + img * D * D;
  for (int r = warp; r < D; r += 8) tile[r * 33 + lane] = src[(int64_t)r * D + c0 + lane];
  __syncthreads();
  const float inv = 1.0f / (float)D;
  for (int c = warp; c < 32; c += 8) {
    float v[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[j] = tile[(j * 32 + lane) * 33 + c];
    fwht_warp<EPL, true>(v, lane);
#pragma unroll
    for (int j = 0; j < EPL; ++j) tile[(j * 32 + lane) * 33 + c] = v[j] * inv;
  }
  __syncthreads();
  for (int r = warp; r < D; r += 8) {
    const int64_t off = img * D * D + (int64_t)r * D + c0 + lane;
    float v = tile[r * 33 + lane];
    if (EPI == WHT_EPI_GRAD) {
      const float xh = xhat[off];
      v = -(2.0f * v) * ((no_clamp || (xh >= -1.0f && xh <= 1.0f)) ? 1.0f : 0.0f);
    }
    out[off] = v;
  }
}

// y[b, k*C + c] = T[b, c, perm[k]]
__global__ void cs_gather_kernel(const float* __restrict__ T, const int32_t* __restrict__ perm, float* __restrict__ y, int C,
                                 int64_t P, int n_meas) {
  const int b = blockIdx.y;
  const int64_t M = (int64_t)n_meas * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / C), c = (int)(e % C);
    y[(int64_t)b * M + e] = T[((int64_t)b * C + c) * P + perm[k]];
  }
}
// temp[b, c, perm[k]] = r[b, k*C + c]  (temp pre-zeroed)
__global__ void cs_scatter_kernel(const float* __restrict__ r, const int32_t* __restrict__ perm, float* __restrict__ temp, int C,
                                  int64_t P, int n_meas) {
  const int b = blockIdx.y;
  const int64_t M = (int64_t)n_meas * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / C), c = (int)(e % C);
    temp[((int64_t)b * C + c) * P + perm[k]] = r[(int64_t)b * M + e];
  }
}
// r = y - T[perm]; loss[b] = sum r^2; temp[b, c, perm[k]] = r (temp pre-zeroed)
__global__ void __launch_bounds__(kThreads)
cs_resid_scatter_kernel(const float* __restrict__ T, const int32_t* __restrict__ perm, const float* __restrict__ y,
                        float* __restrict__ temp, int C, int64_t P, int n_meas, double* loss, RedWs ws) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t M = (int64_t)n_meas * C;
  double acc = 0.0;
  for (int64_t e = (int64_t)blk * kThreads + threadIdx.x; e < M; e += (int64_t)nblk * kThreads) {
    const int k = (int)(e / C), c = (int)(e % C);
    const int64_t pos = ((int64_t)b * C + c) * P + perm[k];
    const float r = __fsub_rn(y[(int64_t)b * M + e], T[pos]);
    acc += (double)r * (double)r;
    temp[pos] = r;
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk)) loss[b] = vsum[0];
}

template <int EPL>
static int wht2d(const float* in, float* mid, float* out, int mats, int clamp, int epi, const float* xhat, int no_clamp,
                 cudaStream_t st) {
  constexpr int D = EPL * 32;
  const int64_t n_rows = (int64_t)mats * D;
  wht_rows_kernel<EPL><<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(in, mid, n_rows, clamp);
  const size_t smem = (size_t)D * 33 * sizeof(float);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(wht_cols_kernel<EPL, WHT_EPI_SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(wht_cols_kernel<EPL, WHT_EPI_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const dim3 grid(D / 32, mats);
  if (epi == WHT_EPI_GRAD) wht_cols_kernel<EPL, WHT_EPI_GRAD><<<grid, 256, smem, st>>>(mid, out, xhat, no_clamp);
  else wht_cols_kernel<EPL, WHT_EPI_SCALE><<<grid, 256, smem, st>>>(mid, out, nullptr, 0);
  return NHMC_LAUNCH_CHECK();
}

static int wht2d_dispatch(int D, const float* in, float* mid, float* out, int mats, int clamp, int epi, const float* xhat,
                          int no_clamp, cudaStream_t st) {
  switch (D) {
    case 32: return wht2d<1>(in, mid, out, mats, clamp, epi, xhat, no_clamp, st);
    case 64: return wht2d<2>(in, mid, out, mats, clamp, epi, xhat, no_clamp, st);
    case 128: return wht2d<4>(in, mid, out, mats, clamp, epi, xhat, no_clamp, st);
    case 256: return wht2d<8>(in, mid, out, mats, clamp, epi, xhat, no_clamp, st);
    case 512: return wht2d<16>(in, mid, out, mats, clamp, epi, xhat, no_clamp, st);
  }
  set_error("Walsh-Hadamard operator: img_dim %d not supported (32, 64, 128, 256, 512)", D);
  return NHMC_ERR_UNSUPPORTED;
}

size_t cs_workspace_bytes(const nhmc_op* op, int n_chains) {
  return (size_t)2 * n_chains * op->channels * op->img_dim * op->img_dim * sizeof(float) + 256;
}

int cs_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* ws, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  const int64_t P = (int64_t)D * D;
  float* t0 = reinterpret_cast<float*>(ws);
  float* t1 = t0 + (size_t)mats * P;
  if (int e = wht2d_dispatch(D, x, t0, t1, mats, 0, WHT_EPI_SCALE, nullptr, 0, st)) return e;
  const int64_t M = (int64_t)op->n_kept * C;
  if (M > 0) cs_gather_kernel<<<dim3(cta_count(M, kThreads, n_chains), n_chains), kThreads, 0, st>>>(t1, op->kept_pix, y, C, P, op->n_kept);
  return NHMC_LAUNCH_CHECK();
}

int cs_adjoint(const nhmc_op* op, const float* r, float* out, int n_chains, void* ws, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  const int64_t P = (int64_t)D * D;
  float* t0 = reinterpret_cast<float*>(ws);
  float* t1 = t0 + (size_t)mats * P;
  if (int e = cuda_status(cudaMemsetAsync(t0, 0, (size_t)mats * P * sizeof(float), st), "cs memset")) return e;
  const int64_t M = (int64_t)op->n_kept * C;
  if (M > 0) cs_scatter_kernel<<<dim3(cta_count(M, kThreads, n_chains), n_chains), kThreads, 0, st>>>(r, op->kept_pix, t0, C, P, op->n_kept);
  return wht2d_dispatch(D, t0, t1, out, mats, 0, WHT_EPI_SCALE, nullptr, 0, st);
}

int cs_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss, int n_chains,
                 void* workspace, cudaStream_t st);

__global__ void cs_clamp_copy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int nc) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = nc ? x[e] : fminf(fmaxf(x[e], -1.0f), 1.0f);
}

int cs_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out, double* loss, int n_chains,
                 void* workspace, cudaStream_t st) {
  const int D = op->img_dim, C = op->channels, mats = n_chains * C;
  const int64_t P = (int64_t)D * D;
  const RedWs rw = red_ws(workspace, n_chains);
  float* t0 = reinterpret_cast<float*>((char*)workspace + red_ws_aligned(n_chains));
  float* t1 = t0 + (size_t)mats * P;
  // T = FWHT(clamp(x_hat)) -> t1
  if (int e = wht2d_dispatch(D, x_hat, t0, t1, mats, op->no_clamp ? 0 : 1, WHT_EPI_SCALE, nullptr, 0, st)) return e;
  // r = y - T[perm] scattered into a zeroed image (t0), loss
  if (int e = cuda_status(cudaMemsetAsync(t0, 0, (size_t)mats * P * sizeof(float), st), "cs memset")) return e;
  const int64_t M = (int64_t)op->n_kept * C;
  int nblk = cta_count(M > 0 ? M : 1, kThreads * 4, n_chains);
  cs_resid_scatter_kernel<<<dim3(nblk, n_chains), kThreads, 0, st>>>(t1, op->kept_pix, y, t0, C, P, op->n_kept, loss, rw);
  // g = -2 FWHT(scatter(r)) .* mask   (t0 -> rows -> t1 -> cols + epilogue -> g)
  if (int e = wht2d_dispatch(D, t0, t1, g, mats, 0, WHT_EPI_GRAD, x_hat, op->no_clamp, st)) return e;
  if (xc_out) {
    const int64_t total = (int64_t)mats * P;
    cs_clamp_copy_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x_hat, xc_out, total, op->no_clamp);
  }
  return NHMC_LAUNCH_CHECK();
}

}  // namespace nhmc
