// Shared device/host helpers for the nhmc kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nhmc.h"

namespace nhmc {

constexpr int kSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kThreads = 256;      // 8 warps; 8 CTAs/SM resident for the streaming kernels

void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define NHMC_CHECK_ARG(cond, msg)            \
  do {                                       \
    if (!(cond)) {                           \
      nhmc::set_error("%s: %s", __func__, msg); \
      return NHMC_ERR_ARG;                   \
    }                                        \
  } while (0)

#define NHMC_LAUNCH_CHECK() nhmc::cuda_status(cudaGetLastError(), __func__)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- programmatic dependent launch (PDL, sm_90+) -------------------------------------------------------------
// The U-Net epilogue kernels (GroupNorm statistics / apply / one-CTA kernel, pooling, residual add) often follow one
// another directly.  Launched with the programmatic-stream-serialization attribute, a kernel's CTAs are scheduled as
// soon as the previous kernel's CTAs have all STARTED and executed pdl_trigger(); they then block in pdl_wait() until
// that kernel has completed and flushed its writes.  Launch latency and CTA scheduling of the dependent overlap the
// tail of its predecessor; captured into a CUDA graph the edge becomes a programmatic dependency.  Kernels whose CTAs
// wait on each other (the resident GroupNorm with its global barrier) never trigger early.  Per call: NHMC_F_NO_PDL in
// the entry point's flag word launches plainly.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

template <typename... KA, typename... A>
static inline cudaError_t launch_pdl(bool pdl, void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}

// ---- streaming loads/stores: data is touched once per kernel, keep it out of L1 ----------
// ld_stream: plain global load (safe when the same thread rewrites the location later);
// ld_ro: read-only path, only for buffers the kernel never writes.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_ro(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w));
}

// ---- element conversion for the score-net tensors ---------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// 4 consecutive elements of type T <-> float4 (8-byte vector access for 16-bit types)
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return ld_stream(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&raw.x);
  __half2 b = *reinterpret_cast<__half2*>(&raw.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float4& v);
template <> __device__ __forceinline__ void store4<float>(float* p, const float4& v) {
  st_stream(reinterpret_cast<float4*>(p), v);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float4& v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// ---- reductions: warp shuffle, then one block-level pass through shared memory ------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of `v` over the block; valid in thread 0 (and in all of warp 0). NV values per thread.
template <typename T, int NV> __device__ __forceinline__ void block_sum(T (&v)[NV], T* smem /* [NV*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // smem may still be in use by a previous call
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      T t = lane < nwarp ? smem[i * 32 + lane] : T(0);
      v[i] = warp_sum(t);
    }
  }
}

// Deterministic cross-block finish: every block of chain `b` writes its partial, the block that
// takes the last ticket adds the partials in index order.  `ticket` wraps back to 0 by itself.
// Returns true in thread 0 of the finishing block with the totals in v.
template <int NV>
__device__ __forceinline__ bool grid_finish(double (&v)[NV], double* partials /* [nblk][NV] for this chain */,
                                            unsigned int* ticket, int blk, int nblk) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) partials[blk * NV + i] = v[i];
    __threadfence();
    unsigned int t = atomicInc(ticket, (unsigned int)(nblk - 1));
    is_last = (t == (unsigned int)(nblk - 1));
  }
  __syncthreads();
  if (!is_last) return false;
  if (threadIdx.x == 0) {
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.0;
    for (int k = 0; k < nblk; ++k) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += __ldcg(&partials[k * NV + i]);
    }
    return true;
  }
  return false;
}

// Deterministic cross-block finish with the whole last block: thread t adds partials t, t + 256, ... in fp64, then one
// fixed-shape block reduction -- the result depends on nblk only, never on arrival order, and the tail costs one
// block-reduce instead of a serial walk over nblk L2 round trips.
template <int NV>
__device__ __forceinline__ bool grid_finish_par(double (&v)[NV], double* partials, unsigned int* ticket, int blk, int nblk,
                                                double* red /* [NV*32] */) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) partials[blk * NV + i] = v[i];
    __threadfence();
    is_last = atomicInc(ticket, (unsigned int)(nblk - 1)) == (unsigned int)(nblk - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.0;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldcg(&partials[k * NV + i]);
  }
  block_sum<double, NV>(v, red);
  return threadIdx.x == 0;
}

// CTAs per chain for a streaming kernel that ENDS in a cross-block reduction: one resident wave (8 CTAs of 256 threads per
// SM over all chains), so that the block reduce + ticket of every CTA is amortised over >= ~10 vectors per thread.
static inline int reduce_cta_count(int64_t items, int64_t per_cta, int n_chains) {
  int64_t want = (items + per_cta - 1) / per_cta;
  int64_t cap = (int64_t)kSMs * 8 / (n_chains > 0 ? n_chains : 1);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

static inline int blocks_per_chain(int64_t n_vec, int n_chains, int vec_per_thread) {
  // enough CTAs to cover the chain once with `vec_per_thread` vectors per thread, but keep the
  // whole grid within ~4 resident waves of the 148 SMs (8 CTAs of 256 threads each per SM).
  int64_t want = (n_vec + (int64_t)kThreads * vec_per_thread - 1) / ((int64_t)kThreads * vec_per_thread);
  int64_t cap = (int64_t)kSMs * 8 * 4 / (n_chains > 0 ? n_chains : 1);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

// CTAs per chain when one CTA handles `per_cta` work items per pass (same residency cap as above).
static inline int cta_count(int64_t items, int64_t per_cta, int n_chains) {
  int64_t want = (items + per_cta - 1) / per_cta;
  int64_t cap = (int64_t)kSMs * 8 * 4 / (n_chains > 0 ? n_chains : 1);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

constexpr int kMaxBlocksPerChain = 256;  // bound for the partial-sum workspace

struct RedWs {  // layout of the reduction part of an operator workspace
  double* partials;       // [B][kMaxBlocksPerChain]
  unsigned int* tickets;  // [B]
};
static inline size_t red_ws_bytes(int n_chains) {
  return (size_t)n_chains * kMaxBlocksPerChain * sizeof(double) + (size_t)n_chains * sizeof(unsigned int) + 64;
}
static inline RedWs red_ws(void* ws, int n_chains) {
  RedWs r;
  r.partials = reinterpret_cast<double*>(ws);
  r.tickets = reinterpret_cast<unsigned int*>(r.partials + (size_t)n_chains * kMaxBlocksPerChain);
  return r;
}

// offset of the operator-specific scratch behind the reduction area
static inline size_t red_ws_aligned(int n_chains) { return (red_ws_bytes(n_chains) + 255) & ~(size_t)255; }

}  // namespace nhmc
