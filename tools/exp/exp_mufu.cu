// How many warps per scheduler does it take to saturate MUFU.EX2?  (sizing the softmax warps of attn.cu)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/exp/exp_mufu.cu -o /tmp/exp_mufu && /tmp/exp_mufu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MIX>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[32];
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MIX) v[i] = fmaf(v[i], 0.999f, -0.01f);
      v[i] = ex2(v[i]);
      if (MIX) v[i] = v[i] - 1.5f;
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mix = 0; mix < 2; ++mix)
    for (int wps = 1; wps <= 8; wps *= 2) {   // warps per scheduler
      const int threads = 128 * wps;
      if (mix) k<1><<<148, threads>>>(out, cyc, iters); else k<0><<<148, threads>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      const double per = (double)c[0] / ((double)iters * 32 * wps);
      printf("mix=%d warps/scheduler=%d : %.2f clk per MUFU warp-instruction per scheduler (%s)\n", mix, wps, per, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
