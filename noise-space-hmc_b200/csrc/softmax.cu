// K10: attention softmax epilogue (unet_ffhq.py:418-425 QKVAttentionLegacy): scale, fp32 softmax over
// the key axis, cast back -- one pass, one warp per row; and its backward
// dlogits = scale * p * (dp - sum(dp * p)).  Rows are T = 64..4096 long (16x16 / 8x8 feature maps here),
// so a row lives in the registers of one warp (cols <= 32*kMaxPerLane) and nothing is re-read.
#include "common.cuh"

namespace nhmc {

constexpr int kMaxPerLane = 32;  // cols <= 1024 in-register; longer rows take the looped kernel

template <typename T>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, float scale, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* src = in + row * cols;
  T* dst = out + row * cols;
  float v[kMaxPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    v[i] = c < cols ? Elem<T>::to_f(src[c]) * scale : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    v[i] = (lane + i * 32 < cols) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = __fdividef(1.0f, sum);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    if (c < cols) dst[c] = Elem<T>::from_f(v[i] * inv);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ dl, float scale, int64_t rows,
                   int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* ps = p + row * cols;
  const T* ds = dp + row * cols;
  T* dst = dl + row * cols;
  float pv[kMaxPerLane], dv[kMaxPerLane];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    pv[i] = c < cols ? Elem<T>::to_f(ps[c]) : 0.f;
    dv[i] = c < cols ? Elem<T>::to_f(ds[c]) : 0.f;
    dot = fmaf(pv[i], dv[i], dot);
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    if (c < cols) dst[c] = Elem<T>::from_f(scale * pv[i] * (dv[i] - dot));
  }
}

// Vector variants (cols % 4 == 0, 4-element aligned rows): a lane owns NCH groups of 4 consecutive elements
// (8-byte loads for 16-bit types, 16-byte for fp32) instead of 32 scalar elements -- the scalar form reached only
// ~2.5 TB/s on the T = 1024 attention maps of the latent U-Net (235 MB of logits per launch).
template <typename T, int NCH>
__global__ void __launch_bounds__(256)
softmax_fwd_vec_kernel(const T* __restrict__ in, T* __restrict__ out, float scale, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* src = in + row * cols;
  T* dst = out + row * cols;
  const int nvec = cols >> 2;
  float4 v[NCH];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 t = load4<T>(src + 4 * c);
      v[i] = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
    } else {
      v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (lane + i * 32 < nvec) {
      v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = warp_sum(sum);
  const float inv = __fdividef(1.0f, sum);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) store4<T>(dst + 4 * c, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256)
softmax_bwd_vec_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ dl, float scale, int64_t rows,
                       int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* ps = p + row * cols;
  const T* ds = dp + row * cols;
  T* dst = dl + row * cols;
  const int nvec = cols >> 2;
  float4 pv[NCH], dv[NCH];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    pv[i] = make_float4(0.f, 0.f, 0.f, 0.f); dv[i] = pv[i];
    if (c < nvec) { pv[i] = load4<T>(ps + 4 * c); dv[i] = load4<T>(ds + 4 * c); }
    dot = fmaf(pv[i].x, dv[i].x, fmaf(pv[i].y, dv[i].y, fmaf(pv[i].z, dv[i].z, fmaf(pv[i].w, dv[i].w, dot))));
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < nvec)
      store4<T>(dst + 4 * c, make_float4(scale * pv[i].x * (dv[i].x - dot), scale * pv[i].y * (dv[i].y - dot),
                                         scale * pv[i].z * (dv[i].z - dot), scale * pv[i].w * (dv[i].w - dot)));
  }
}

// long rows (latent decoder: T = 4096): one CTA per row, three passes that stay in L1/L2
template <typename T>
__global__ void __launch_bounds__(256)
softmax_fwd_long_kernel(const T* __restrict__ in, T* __restrict__ out, float scale, int cols) {
  __shared__ float red[32];
  const T* src = in + (int64_t)blockIdx.x * cols;
  T* dst = out + (int64_t)blockIdx.x * cols;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, Elem<T>::to_f(src[c]) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s[1] = {0.f};
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s[0] += __expf(Elem<T>::to_f(src[c]) * scale - mx);
  block_sum<float, 1>(s, red);
  __shared__ float total;
  if (threadIdx.x == 0) total = s[0];
  __syncthreads();
  const float inv = __fdividef(1.0f, total);
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    dst[c] = Elem<T>::from_f(__expf(Elem<T>::to_f(src[c]) * scale - mx) * inv);
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_long_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ dl, float scale, int cols) {
  __shared__ float red[32];
  const int64_t off = (int64_t)blockIdx.x * cols;
  float s[1] = {0.f};
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s[0] = fmaf(Elem<T>::to_f(p[off + c]), Elem<T>::to_f(dp[off + c]), s[0]);
  block_sum<float, 1>(s, red);
  __shared__ float total;
  if (threadIdx.x == 0) total = s[0];
  __syncthreads();
  const float dot = total;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    dl[off + c] = Elem<T>::from_f(scale * Elem<T>::to_f(p[off + c]) * (Elem<T>::to_f(dp[off + c]) - dot));
}

// long rows held in the registers of ONE CTA (cols <= 256 threads * 4 * NCH): single read, single write
template <typename T, int NCH>
__global__ void __launch_bounds__(256)
softmax_fwd_row_kernel(const T* __restrict__ in, T* __restrict__ out, float scale, int cols) {
  __shared__ float red[32];
  __shared__ float bc;
  const T* src = in + (int64_t)blockIdx.x * cols;
  T* dst = out + (int64_t)blockIdx.x * cols;
  const int nvec = cols >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 v[NCH];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c < nvec) {
      const float4 t = load4<T>(src + 4 * c);
      v[i] = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
    }
    mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red[0]; for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]); bc = m; }
  __syncthreads();
  mx = bc;
  float s[1] = {0.f};
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (threadIdx.x + i * 256 < nvec) {
      v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
      s[0] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  block_sum<float, 1>(s, red);
  if (threadIdx.x == 0) bc = s[0];
  __syncthreads();
  const float inv = __fdividef(1.0f, bc);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec) store4<T>(dst + 4 * c, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256)
softmax_bwd_row_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ dl, float scale, int cols) {
  __shared__ float red[32];
  __shared__ float bc;
  const int64_t off = (int64_t)blockIdx.x * cols;
  const int nvec = cols >> 2;
  float4 pv[NCH], dv[NCH];
  float s[1] = {0.f};
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * 256;
    pv[i] = make_float4(0.f, 0.f, 0.f, 0.f); dv[i] = pv[i];
    if (c < nvec) { pv[i] = load4<T>(p + off + 4 * c); dv[i] = load4<T>(dp + off + 4 * c); }
    s[0] = fmaf(pv[i].x, dv[i].x, fmaf(pv[i].y, dv[i].y, fmaf(pv[i].z, dv[i].z, fmaf(pv[i].w, dv[i].w, s[0]))));
  }
  block_sum<float, 1>(s, red);
  if (threadIdx.x == 0) bc = s[0];
  __syncthreads();
  const float dot = bc;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec)
      store4<T>(dl + off + 4 * c, make_float4(scale * pv[i].x * (dv[i].x - dot), scale * pv[i].y * (dv[i].y - dot),
                                              scale * pv[i].z * (dv[i].z - dot), scale * pv[i].w * (dv[i].w - dot)));
  }
}

static inline bool vec_ok(const void* a, const void* b, const void* c, int cols, size_t elem) {
  auto al = [&](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (4 * elem)) == 0; };
  return cols % 4 == 0 && al(a) && al(b) && al(c);
}

template <typename T>
static int run_fwd(const void* in, void* out, float scale, int64_t rows, int cols, cudaStream_t st) {
  if (cols <= 32 * kMaxPerLane) {
    const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
    if (vec_ok(in, out, nullptr, cols, sizeof(T))) {
#define SM_FWD(N) softmax_fwd_vec_kernel<T, N><<<blocks, 256, 0, st>>>((const T*)in, (T*)out, scale, rows, cols)
      if (cols <= 128) SM_FWD(1); else if (cols <= 256) SM_FWD(2); else if (cols <= 512) SM_FWD(4); else SM_FWD(8);
#undef SM_FWD
    } else {
      softmax_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)in, (T*)out, scale, rows, cols);
    }
  } else if (cols <= 8192 && vec_ok(in, out, nullptr, cols, sizeof(T))) {
    if (cols <= 4096) softmax_fwd_row_kernel<T, 4><<<(unsigned)rows, 256, 0, st>>>((const T*)in, (T*)out, scale, cols);
    else softmax_fwd_row_kernel<T, 8><<<(unsigned)rows, 256, 0, st>>>((const T*)in, (T*)out, scale, cols);
  } else {
    softmax_fwd_long_kernel<T><<<(unsigned)rows, 256, 0, st>>>((const T*)in, (T*)out, scale, cols);
  }
  return NHMC_LAUNCH_CHECK();
}
template <typename T>
static int run_bwd(const void* p, const void* dp, void* dl, float scale, int64_t rows, int cols, cudaStream_t st) {
  if (cols <= 32 * kMaxPerLane) {
    const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
    if (vec_ok(p, dp, dl, cols, sizeof(T))) {
#define SM_BWD(N) softmax_bwd_vec_kernel<T, N><<<blocks, 256, 0, st>>>((const T*)p, (const T*)dp, (T*)dl, scale, rows, cols)
      if (cols <= 128) SM_BWD(1); else if (cols <= 256) SM_BWD(2); else if (cols <= 512) SM_BWD(4); else SM_BWD(8);
#undef SM_BWD
    } else {
      softmax_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)p, (const T*)dp, (T*)dl, scale, rows, cols);
    }
  } else if (cols <= 8192 && vec_ok(p, dp, dl, cols, sizeof(T))) {
    if (cols <= 4096) softmax_bwd_row_kernel<T, 4><<<(unsigned)rows, 256, 0, st>>>((const T*)p, (const T*)dp, (T*)dl, scale, cols);
    else softmax_bwd_row_kernel<T, 8><<<(unsigned)rows, 256, 0, st>>>((const T*)p, (const T*)dp, (T*)dl, scale, cols);
  } else {
    softmax_bwd_long_kernel<T><<<(unsigned)rows, 256, 0, st>>>((const T*)p, (const T*)dp, (T*)dl, scale, cols);
  }
  return NHMC_LAUNCH_CHECK();
}

}  // namespace nhmc

using namespace nhmc;

extern "C" int nhmc_softmax_fwd(int dtype, const void* logits, void* probs, float scale, int64_t rows, int cols,
                                nhmc_stream_t stream) {
  NHMC_CHECK_ARG(logits && probs && rows > 0 && cols > 0, "bad arguments");
  NHMC_CHECK_ARG(rows < (int64_t)1 << 31, "too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NHMC_F32: return run_fwd<float>(logits, probs, scale, rows, cols, st);
    case NHMC_BF16: return run_fwd<__nv_bfloat16>(logits, probs, scale, rows, cols, st);
    case NHMC_F16: return run_fwd<__half>(logits, probs, scale, rows, cols, st);
  }
  set_error("nhmc_softmax_fwd: unknown dtype %d", dtype);
  return NHMC_ERR_ARG;
}

extern "C" int nhmc_softmax_bwd(int dtype, const void* probs, const void* dprobs, void* dlogits, float scale,
                                int64_t rows, int cols, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(probs && dprobs && dlogits && rows > 0 && cols > 0, "bad arguments");
  NHMC_CHECK_ARG(rows < (int64_t)1 << 31, "too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NHMC_F32: return run_bwd<float>(probs, dprobs, dlogits, scale, rows, cols, st);
    case NHMC_BF16: return run_bwd<__nv_bfloat16>(probs, dprobs, dlogits, scale, rows, cols, st);
    case NHMC_F16: return run_bwd<__half>(probs, dprobs, dlogits, scale, rows, cols, st);
  }
  set_error("nhmc_softmax_bwd: unknown dtype %d", dtype);
  return NHMC_ERR_ARG;
}
