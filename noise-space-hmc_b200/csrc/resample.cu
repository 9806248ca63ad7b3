// U-Net resampling / skip-concat epilogues, channels-last (unet_ffhq.py:153-211 Upsample/Downsample without
// conv inside the up/down ResBlocks :307-316, and the skip concat of :729-731).
//   gather4 : out[b,i,j,:] = scale * sum_{di,dj in {0,1}} in[b,2i+di,2j+dj,:]   (avg_pool2d fwd: scale 1/4;
//                                                                               nearest-upsample bwd: scale 1)
//   scatter4: out[b,2i+di,2j+dj,:] = scale * in[b,i,j,:]                         (nearest-upsample fwd: scale 1;
//                                                                               avg_pool2d bwd: scale 1/4)
//   split2  : backward of cat([a, b], channel): two dense channels-last gradients in one pass
// HBM-bound, 16-byte vectors along C, four independent loads (gather) or stores (scatter) per thread.
#include "common.cuh"

namespace nhmc {

template <typename T> struct V16 { static constexpr int N = 16 / sizeof(T); };

__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

template <typename T> __device__ __forceinline__ void up(const uint4& t, float (&v)[16 / sizeof(T)]);
template <> __device__ __forceinline__ void up<float>(const uint4& t, float (&v)[4]) {
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}
template <> __device__ __forceinline__ void up<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void up<__half>(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ uint4 down(const float (&v)[16 / sizeof(T)]);
template <> __device__ __forceinline__ uint4 down<float>(const float (&v)[4]) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
template <> __device__ __forceinline__ uint4 down<__nv_bfloat16>(const float (&v)[8]) {
  uint4 t; uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
  return t;
}
template <> __device__ __forceinline__ uint4 down<__half>(const float (&v)[8]) {
  uint4 t; uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
  return t;
}

// small = [B, h, w, C], big = [B, 2h, 2w, C]; one thread per 16-byte vector of `small`
template <typename T, bool GATHER>
__global__ void __launch_bounds__(256)
pool2_kernel(const T* __restrict__ in, T* __restrict__ out, int h, int w, int cvec, int64_t total, float scale) {
  pdl_wait(); pdl_trigger();
  constexpr int V = V16<T>::N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    int64_t t = idx / cvec;
    const int j = (int)(t % w); t /= w;
    const int i = (int)(t % h);
    const int64_t b = t / h;
    const int64_t C = (int64_t)cvec * V;
    const int64_t big0 = (((b * 2 * h + 2 * i) * 2 * w) + 2 * j) * C + (int64_t)cv * V;
    const int64_t rowb = (int64_t)2 * w * C;
    if (GATHER) {
      const uint4 r0 = ldg16(in + big0), r1 = ldg16(in + big0 + C), r2 = ldg16(in + big0 + rowb), r3 = ldg16(in + big0 + rowb + C);
      float a[V], b1[V], c[V], d[V], o[V];
      up<T>(r0, a); up<T>(r1, b1); up<T>(r2, c); up<T>(r3, d);
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = ((a[k] + b1[k]) + (c[k] + d[k])) * scale;
      stg16(out + idx * V, down<T>(o));
    } else {
      uint4 r = ldg16(in + idx * V);
      if (scale != 1.0f) {
        float a[V];
        up<T>(r, a);
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] *= scale;
        r = down<T>(a);
      }
      stg16(out + big0, r); stg16(out + big0 + C, r); stg16(out + big0 + rowb, r); stg16(out + big0 + rowb + C, r);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
split2_kernel(const T* __restrict__ in, T* __restrict__ o1, T* __restrict__ o2, int64_t rows, int c1v, int c2v) {
  constexpr int V = V16<T>::N;
  const int cv = c1v + c2v;
  const int64_t total = rows * cv;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cv;
    const int k = (int)(idx - r * cv);
    const uint4 v = ldg16(in + idx * V);
    if (k < c1v) stg16(o1 + (r * c1v + k) * V, v);
    else stg16(o2 + (r * c2v + (k - c1v)) * V, v);
  }
}

template <typename T, bool HAS_B, bool HAS_BIAS>
__global__ void __launch_bounds__(256)
bias_add_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, const float* __restrict__ bias, T* out, int64_t total, int cvec) {
  pdl_wait(); pdl_trigger();
  constexpr int V = V16<T>::N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    float va[V], vb[V];
    // plain (coherent) loads: out may alias a
    const uint4 ra = *reinterpret_cast<const uint4*>(a + idx * V);
    up<T>(ra, va);
    if (HAS_B) { up<T>(ldg16(b + idx * V), vb); }
    if (HAS_B && c != nullptr) {
      float vc[V];
      up<T>(ldg16(c + idx * V), vc);
#pragma unroll
      for (int k = 0; k < V; ++k) vb[k] += vc[k];
    }
    const int c0 = (int)(idx % cvec) * V;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float v = va[k];
      if (HAS_B) v += vb[k];
      if (HAS_BIAS) v += __ldg(bias + c0 + k);
      va[k] = v;
    }
    *reinterpret_cast<uint4*>(out + idx * V) = down<T>(va);
  }
}

static inline unsigned grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kSMs * 32;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

template <typename T>
static int run_pool(bool gather, const void* in, void* out, int batch, int h, int w, int C, float scale, bool pdl, cudaStream_t st) {
  constexpr int V = V16<T>::N;
  NHMC_CHECK_ARG(C % V == 0, "channels must be a multiple of the 16-byte vector width");
  const int64_t total = (int64_t)batch * h * w * (C / V);
  const dim3 g(grid_for(total)), b(256);
  if (gather) return cuda_status(launch_pdl(pdl, pool2_kernel<T, true>, g, b, 0, st, (const T*)in, (T*)out, h, w, C / V, total, scale), "pool2 launch");
  return cuda_status(launch_pdl(pdl, pool2_kernel<T, false>, g, b, 0, st, (const T*)in, (T*)out, h, w, C / V, total, scale), "pool2 launch");
}

}  // namespace nhmc

using namespace nhmc;

static int pool_dispatch(bool gather, int dtype, const void* in, void* out, int batch, int h, int w, int C, float scale,
                         unsigned int flags, nhmc_stream_t stream) {
  const bool pdl = !(flags & NHMC_F_NO_PDL);
  NHMC_CHECK_ARG(in && out && batch > 0 && h > 0 && w > 0 && C > 0, "bad arguments");
  NHMC_CHECK_ARG(aligned16(in) && aligned16(out), "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NHMC_F32: return run_pool<float>(gather, in, out, batch, h, w, C, scale, pdl, st);
    case NHMC_BF16: return run_pool<__nv_bfloat16>(gather, in, out, batch, h, w, C, scale, pdl, st);
    case NHMC_F16: return run_pool<__half>(gather, in, out, batch, h, w, C, scale, pdl, st);
  }
  set_error("pool2: unknown dtype %d", dtype);
  return NHMC_ERR_ARG;
}

extern "C" int nhmc_pool2_gather(int dtype, const void* in_big, void* out_small, int batch, int h_small, int w_small,
                                 int channels, float scale, unsigned int flags, nhmc_stream_t stream) {
  return pool_dispatch(true, dtype, in_big, out_small, batch, h_small, w_small, channels, scale, flags, stream);
}
extern "C" int nhmc_pool2_scatter(int dtype, const void* in_small, void* out_big, int batch, int h_small, int w_small,
                                  int channels, float scale, unsigned int flags, nhmc_stream_t stream) {
  return pool_dispatch(false, dtype, in_small, out_big, batch, h_small, w_small, channels, scale, flags, stream);
}

extern "C" int nhmc_split2(int dtype, const void* in, void* out1, void* out2, int64_t rows, int c1, int c2,
                           nhmc_stream_t stream) {
  NHMC_CHECK_ARG(in && out1 && out2 && rows > 0 && c1 > 0 && c2 > 0, "bad arguments");
  NHMC_CHECK_ARG(aligned16(in) && aligned16(out1) && aligned16(out2), "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int V = dtype == NHMC_F32 ? 4 : 8;
  NHMC_CHECK_ARG(c1 % V == 0 && c2 % V == 0, "channel counts must be multiples of the 16-byte vector width");
  const int64_t total = rows * ((c1 + c2) / V);
  switch (dtype) {
    case NHMC_F32: split2_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)in, (float*)out1, (float*)out2, rows, c1 / V, c2 / V); break;
    case NHMC_BF16: split2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out1, (__nv_bfloat16*)out2, rows, c1 / V, c2 / V); break;
    case NHMC_F16: split2_kernel<__half><<<grid_for(total), 256, 0, st>>>((const __half*)in, (__half*)out1, (__half*)out2, rows, c1 / V, c2 / V); break;
    default: set_error("split2: unknown dtype %d", dtype); return NHMC_ERR_ARG;
  }
  return NHMC_LAUNCH_CHECK();
}

template <typename T>
static int run_bias_add(const void* a, const void* b, const void* c, const float* bias, void* out, int64_t rows, int C, bool pdl, cudaStream_t st) {
  constexpr int V = V16<T>::N;
  NHMC_CHECK_ARG(C % V == 0, "channels must be a multiple of the 16-byte vector width");
  const int64_t total = rows * (C / V);
  const unsigned g = grid_for(total);
#define BA(HB, HBI) return cuda_status(launch_pdl(pdl, bias_add_kernel<T, HB, HBI>, dim3(g), dim3(256), 0, st, (const T*)a, (const T*)b, (const T*)c, bias, (T*)out, total, C / V), "bias_add launch")
  if (b) { if (bias) BA(true, true); else BA(true, false); }
  else   { if (bias) BA(false, true); else BA(false, false); }
#undef BA
}

extern "C" int nhmc_bias_add(int dtype, const void* a, const void* b, const void* c, const float* bias, void* out,
                             int64_t rows, int channels, unsigned int flags, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(a && out && rows > 0 && channels > 0, "bad arguments");
  NHMC_CHECK_ARG(!c || b, "a third addend needs the second");
  NHMC_CHECK_ARG(aligned16(a) && aligned16(out) && (!b || aligned16(b)) && (!c || aligned16(c)), "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NHMC_F32: return run_bias_add<float>(a, b, c, bias, out, rows, channels, !(flags & NHMC_F_NO_PDL), st);
    case NHMC_BF16: return run_bias_add<__nv_bfloat16>(a, b, c, bias, out, rows, channels, !(flags & NHMC_F_NO_PDL), st);
    case NHMC_F16: return run_bias_add<__half>(a, b, c, bias, out, rows, channels, !(flags & NHMC_F_NO_PDL), st);
  }
  set_error("bias_add: unknown dtype %d", dtype);
  return NHMC_ERR_ARG;
}
