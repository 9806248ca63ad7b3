// Experiment harness (not product code): launch-shape / unroll variants of the DDIM epilogue forward kernel, timed with
// CUDA events on a 64-chain problem with a cold, clean L2.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
// -I../../noise-space-hmc_b200/csrc exp_ddim.cu -o exp_ddim ; run on the GPU box.  Results -> profiles/r2_stream_variants.md
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ float4 ld_ro(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ float4 load4bf(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <int U, int MINB, int DIV = 0>
__global__ void __launch_bounds__(256, MINB)
ddim_fwd(const float* __restrict__ xt, const __nv_bfloat16* __restrict__ eps, int64_t eps_stride, float* __restrict__ out,
         uint32_t* __restrict__ mask, float c1, float sa, float c2, float ss, int64_t n_groups, int64_t n_per_chain) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * 256) >> 5;
  const float* xb = xt + (int64_t)b * n_per_chain;
  const __nv_bfloat16* eb = eps + (int64_t)b * eps_stride;
  float* ob = out + (int64_t)b * n_per_chain;
  uint32_t* mb = mask + (int64_t)b * n_groups * 4;
  const float rcp = __frcp_rn(sa);
  for (int64_t g0 = warp0; g0 < n_groups; g0 += nwarps * U) {
    float4 xv[U], ev[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t gidx = g0 + (int64_t)u * nwarps;
      const int64_t e0 = gidx * 128 + lane * 4;
      const bool ok = gidx < n_groups;
      xv[u] = ok ? ld_ro(reinterpret_cast<const float4*>(xb + e0)) : make_float4(0, 0, 0, 0);
      ev[u] = ok ? load4bf(eb + e0) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t gidx = g0 + (int64_t)u * nwarps;
      if (gidx >= n_groups) break;
      const int64_t e0 = gidx * 128 + lane * 4;
      float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      const float es[4] = {ev[u].x, ev[u].y, ev[u].z, ev[u].w};
      uint32_t bits[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float num = __fsub_rn(xs[j], __fmul_rn(es[j], c1));
        float x0raw;
        if (DIV == 0) x0raw = __fdiv_rn(num, sa);
        else if (DIV == 1) { const float q0 = __fmul_rn(num, rcp); const float r = __fmaf_rn(-q0, sa, num); x0raw = __fmaf_rn(r, rcp, q0); }
        else x0raw = __fmul_rn(num, rcp);
        const bool in = (x0raw >= -1.0f) && (x0raw <= 1.0f);
        const float x0 = fminf(fmaxf(x0raw, -1.0f), 1.0f);
        xs[j] = __fadd_rn(__fmul_rn(ss, x0), __fmul_rn(c2, es[j]));
        bits[j] = __ballot_sync(0xffffffffu, in);
      }
      st_stream(reinterpret_cast<float4*>(ob + e0), make_float4(xs[0], xs[1], xs[2], xs[3]));
      if (lane < 4) mb[gidx * 4 + lane] = lane == 0 ? bits[0] : lane == 1 ? bits[1] : lane == 2 ? bits[2] : bits[3];
    }
  }
}

// plain copy with the same byte mix (16 B read + 8 B read + 16 B write per lane) as the ceiling for this access pattern
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
mix_copy(const float4* __restrict__ a, const uint2* __restrict__ e, float4* __restrict__ o, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += stride * U) {
    float4 v[U]; uint2 w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * stride; if (i < n) { v[u] = ld_ro(a + i); w[u] = e[i]; } }
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * stride; if (i < n) { v[u].x += __uint_as_float(w[u].x & 1u); st_stream(o + i, v[u]); } }
  }
}

__device__ __forceinline__ float hash01(size_t i) { uint32_t h = (uint32_t)(i * 2654435761u) ^ (uint32_t)(i >> 17); h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; return (h & 0xFFFFFF) / 16777216.0f; }
__global__ void rnd_f32(float* p, size_t n) { for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = 6.f * hash01(i) - 3.f; }
__global__ void rnd_bf16(__nv_bfloat16* p, size_t n) { for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = __float2bfloat16(4.f * hash01(i + 77) - 2.f); }
static float* flushbuf; static size_t flushn = (size_t)256 << 20;
__global__ void fill(float* p, size_t n) { for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = 0.f; }
__global__ void readsum(const float* p, size_t n, float* o) { float s = 0; for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += p[i]; if (s == 12345.f) *o = s; }

template <typename F> static float timeit(F f) {
  std::vector<float> ms;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int r = 0; r < 15; ++r) {
    fill<<<1184, 256>>>(flushbuf, flushn / 4); readsum<<<1184, 256>>>(flushbuf, flushn / 4, flushbuf);
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float t; cudaEventElapsedTime(&t, a, b); if (r >= 3) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}

int main() {
  const int B = 64; const int64_t N = 3 * 256 * 256, ES = 6 * 256 * 256, G = N / 128;
  float *x, *o; __nv_bfloat16* e; uint32_t* m;
  cudaMalloc(&x, B * N * 4); cudaMalloc(&o, B * N * 4); cudaMalloc(&e, B * ES * 2); cudaMalloc(&m, B * G * 16); cudaMalloc(&flushbuf, flushn);
  rnd_f32<<<1184, 256>>>(x, (size_t)B * N); rnd_bf16<<<1184, 256>>>(e, (size_t)B * ES);
  const double bytes = (double)B * N * (4 + 2 + 4) + (double)B * N / 8;
#define RUND(U, MINB, CTAS, DIV)                                                                                       \
  {                                                                                                                    \
    float t = timeit([&] { ddim_fwd<U, MINB, DIV><<<dim3(CTAS, B), 256>>>(x, e, ES, o, m, 0.96f, 0.28f, 0.69f, 0.72f, G, N); }); \
    printf("ddim_fwd U=%d minb=%d ctas/chain=%d div=%d : %.2f us  %.0f GB/s  (%s)\n", U, MINB, CTAS, DIV, t * 1e3, bytes / 1e9 / (t / 1e3), cudaGetErrorString(cudaGetLastError())); \
  }
#define RUN(U, MINB, CTAS) RUND(U, MINB, CTAS, 0)
  RUN(1, 8, 74) RUN(1, 8, 96) RUN(1, 8, 148) RUN(1, 8, 37) RUN(1, 8, 18)
  RUN(2, 8, 74) RUN(2, 8, 37) RUN(2, 8, 18) RUN(2, 6, 37) RUN(2, 4, 37)
  RUN(4, 4, 37) RUN(4, 4, 18) RUN(4, 4, 9) RUN(4, 6, 18)
  RUND(1, 8, 74, 1) RUND(1, 8, 74, 2) RUND(2, 6, 37, 1) RUND(2, 6, 37, 2) RUND(4, 4, 18, 1) RUND(4, 4, 18, 2) RUND(2, 8, 74, 1) RUND(4, 4, 37, 1)
  const int64_t nv = B * N / 4;
#define RUNC(U, MINB, CTAS)                                                                                            \
  {                                                                                                                    \
    float t = timeit([&] { mix_copy<U, MINB><<<CTAS, 256>>>((const float4*)x, (const uint2*)e, (float4*)o, nv); });        \
    printf("mix_copy U=%d minb=%d ctas=%d : %.2f us  %.0f GB/s\n", U, MINB, CTAS, t * 1e3, (double)nv * 40 / 1e9 / (t / 1e3));  \
  }
  RUNC(1, 8, 1184) RUNC(1, 8, 4736) RUNC(2, 8, 1184) RUNC(4, 8, 1184) RUNC(4, 4, 592) RUNC(8, 4, 592) RUNC(4, 8, 2368)
  return 0;
}
