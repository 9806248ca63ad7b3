// tcgen05 batched GEMM with 3xTF32 operand splitting (see gemm_tc.cuh) + the host-side tensor-map plumbing (tc5.cuh).
//
// One CTA = one 128 x 128 tile of C, the whole K loop.  Warp-specialised, as the hardware wants it:
//   warp 0   (one lane)  TMA producer : per 32-wide K block one 4-D box {32 K, 128 rows, 1 matrix, 2 planes} per operand
//                                        -> shared memory, 128-byte swizzle, arrives on full[stage] (expect_tx)
//   warp 1   (one lane)  MMA issuer   : tcgen05.mma.kind::tf32 M128 N128 K8, 3 products per K slice (hi*lo, lo*hi, hi*hi),
//                                        fp32 accumulators in TMEM (128 columns); tcgen05.commit frees the stage / publishes
//   warps 2-5            epilogue     : tcgen05.ld (thread = accumulator row) -> elementwise epilogue -> global memory
// Epilogues are those of the SIMT kernel this replaces (deblur.cu, round 1): spectral gain, residual + sum r^2 (fp64,
// deterministic last-block finish), final clip mask, and the hi/lo split of the result for the next product of the chain.
#include "gemm_tc.cuh"
#include "tc5.cuh"

namespace nhmc {
namespace tc5 {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NHMC_ERR_UNSUPPORTED; }
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return NHMC_ERR_CUDA; }
  return NHMC_OK;
}

}  // namespace tc5

using namespace tc5;

namespace {

constexpr int BM = 128, BK = 32;                    // BK fp32 = one 128-byte swizzle row
constexpr int PLANE_A = BM * BK * 4;
constexpr int TC_THREADS = 192;
// Tile width BN: 128 (3 stages of 64 KiB) while that keeps the launch within one wave of the chip, else 256 (2 stages of
// 96 KiB): half the CTAs, each A tile reused over twice the columns, the widest MMA shape.
template <int BN> struct TcTile {
  static constexpr int PLANE_B = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * PLANE_A + 2 * PLANE_B;
  static constexpr int STAGES = BN == 128 ? 3 : 2;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers etc. */;
};

struct TcKernelArgs {
  int M, N, K, channels;
  int a_shared, b_shared;
  int epi, out_transposed, no_clamp;
  float* out_hi; float* out_lo; int64_t out_mat_stride;
  const float* gain; const float* y; const float* xhat;
  double* loss;
  RedWs ws;
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int EPI, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcKernelArgs a) {
  constexpr int PLANE_B = TcTile<BN>::PLANE_B, STAGE_BYTES = TcTile<BN>::STAGE_BYTES, STAGES = TcTile<BN>::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
  double* red = reinterpret_cast<double*>(tmem_slot + 2);             // [4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = blockIdx.z, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int num_kb = (a.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
  // tail of the previous product of the chain; its results are visible from here on
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        unsigned char* st = tiles + s * STAGE_BYTES;
        mbar_arrive_expect_tx(full + s, STAGE_BYTES);
        tma_load_4d(st, &mapA, full + s, kb * BK, m0, a.a_shared ? 0 : mat, 0);
        tma_load_4d(st + 2 * PLANE_A, &mapB, full + s, kb * BK, n0, a.b_shared ? 0 : mat, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(2 /* tf32 */, BM, BN);
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sa = smem_addr(tiles + s * STAGE_BYTES);
        const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + PLANE_A);
        const uint64_t b_hi = umma_desc_sw128(sa + 2 * PLANE_A), b_lo = umma_desc_sw128(sa + 2 * PLANE_A + PLANE_B);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t off = (uint64_t)(ks * 32 >> 4);                 // 8 tf32 = 32 bytes along the swizzled row
          umma_tf32(tmem, a_hi + off, b_lo + off, idesc, acc);
          umma_tf32(tmem, a_lo + off, b_hi + off, idesc, 1u);
          umma_tf32(tmem, a_hi + off, b_hi + off, idesc, 1u);
          acc = 1u;
        }
        umma_commit(empty + s);                                          // stage reusable once these MMAs have read it
      }
      umma_commit(accum_full);
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes [32*(w%4), +32) = rows of the tile ----
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < a.M;
    const int c = mat % a.channels;
    double lsum = 0.0;
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int64_t MN = (int64_t)a.M * a.N;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      tmem_ld_wait();
      const int nb = n0 + ch * 32;
      if (m_ok && nb < a.N) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int64_t row = (int64_t)m * a.N + nb;
      if (EPI == TC_EPI_GAIN) {
        const float4* gp = reinterpret_cast<const float4*>(a.gain + (int64_t)c * MN + row);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (nb + j4 * 4 < a.N) {
            const float4 g = __ldg(gp + j4);
            v[j4 * 4] *= g.x; v[j4 * 4 + 1] *= g.y; v[j4 * 4 + 2] *= g.z; v[j4 * 4 + 3] *= g.w;
          }
        }
      }
      if (EPI == TC_EPI_SPEC) {
        // spectral-domain residual of the deblurring operators (deblur.cu): r = y_spec - gain * v, loss += r^2, out = gain * r
        const float4* gp = reinterpret_cast<const float4*>(a.gain + (int64_t)c * MN + row);
        const float4* yp = reinterpret_cast<const float4*>(a.y + (int64_t)mat * MN + row);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (nb + j4 * 4 < a.N) {
            const float4 g = __ldg(gp + j4), yv = __ldg(yp + j4);
            const float g4[4] = {g.x, g.y, g.z, g.w}, y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float rr = __fsub_rn(y4[i], __fmul_rn(g4[i], v[j4 * 4 + i]));
              lsum += (double)rr * (double)rr;
              v[j4 * 4 + i] = __fmul_rn(g4[i], rr);
            }
          }
        }
      }
      if (EPI == TC_EPI_RESID) {
        const float4* yp = reinterpret_cast<const float4*>(a.y + (int64_t)mat * MN + row);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (nb + j4 * 4 < a.N) {
            const float4 yv = __ldg(yp + j4);
            const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float rr = __fsub_rn(y4[i], v[j4 * 4 + i]);
              v[j4 * 4 + i] = rr;
              lsum += (double)rr * (double)rr;
            }
          }
        }
      }
      if (EPI == TC_EPI_GRAD) {
        const float4* xp = reinterpret_cast<const float4*>(a.xhat + (int64_t)mat * MN + row);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (nb + j4 * 4 < a.N) {
            const float4 xv = __ldg(xp + j4);
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              v[j4 * 4 + i] = -(2.0f * v[j4 * 4 + i]) * ((a.no_clamp || (x4[i] >= -1.0f && x4[i] <= 1.0f)) ? 1.0f : 0.0f);
          }
        }
      }
      float* oh = a.out_hi + (int64_t)mat * a.out_mat_stride;
      float* ol = a.out_lo ? a.out_lo + (int64_t)mat * a.out_mat_stride : nullptr;
      if (a.out_transposed) {
        // C^T[n][m]: consecutive lanes = consecutive m -> each store instruction is one coalesced 128-byte line
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (nb + j < a.N) {
            const int64_t o = (int64_t)(nb + j) * a.M + m;
            if (ol) { const float h = tf32_hi(v[j]); oh[o] = h; ol[o] = v[j] - h; } else oh[o] = v[j];
          }
        }
      } else {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          if (nb + j4 * 4 < a.N) {
            float4 h = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (ol) {
              const float4 f = h;
              h = make_float4(tf32_hi(f.x), tf32_hi(f.y), tf32_hi(f.z), tf32_hi(f.w));
              *reinterpret_cast<float4*>(ol + row + j4 * 4) = make_float4(f.x - h.x, f.y - h.y, f.z - h.z, f.w - h.w);
            }
            *reinterpret_cast<float4*>(oh + row + j4 * 4) = h;
          }
        }
      }
      }
    }
    if (EPI == TC_EPI_RESID || EPI == TC_EPI_SPEC) {
      lsum = warp_sum(lsum);
      if (lane == 0) red[q] = lsum;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<BN>(tmem); }
  if (EPI == TC_EPI_RESID || EPI == TC_EPI_SPEC) {
    double v[1] = {red[0] + red[1] + red[2] + red[3]};
    const int chain = mat / a.channels, c = mat % a.channels;
    const int tiles_c = gridDim.x * gridDim.y;
    const int blk = (c * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (grid_finish<1>(v, a.ws.partials + (int64_t)chain * kMaxBlocksPerChain, a.ws.tickets + chain, blk, tiles_c * a.channels))
      a.loss[chain] = v[0];
  }
}

__global__ void tc_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n4, int clamp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    if (clamp) {
      v.x = fminf(fmaxf(v.x, -1.0f), 1.0f); v.y = fminf(fmaxf(v.y, -1.0f), 1.0f);
      v.z = fminf(fmaxf(v.z, -1.0f), 1.0f); v.w = fminf(fmaxf(v.w, -1.0f), 1.0f);
    }
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}

__global__ void tc_split_table_kernel(const float* __restrict__ t, float* __restrict__ hi, float* __restrict__ lo, int rows, int cols,
                                      int transpose) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? t[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    float v; int64_t o; bool ok;
    if (transpose) {   // out[c][r]
      const int c = c0 + i, r = r0 + threadIdx.x;
      v = tile[threadIdx.x][i]; o = (int64_t)c * rows + r; ok = r < rows && c < cols;
    } else {
      const int r = r0 + i, c = c0 + threadIdx.x;
      v = tile[i][threadIdx.x]; o = (int64_t)r * cols + c; ok = r < rows && c < cols;
    }
    if (ok) { const float h = tf32_hi(v); hi[o] = h; lo[o] = v - h; }
  }
}

int operand_map(CUtensorMap* map, const TcOperand& op, int mats, int box_rows) {
  const bool shared = op.mat_stride == 0;
  const uint64_t dims[4] = {(uint64_t)op.cols, (uint64_t)op.rows, (uint64_t)(shared ? 1 : mats), 2};
  const uint64_t strides[3] = {(uint64_t)op.cols * 4, (uint64_t)(shared ? (int64_t)op.rows * op.cols : op.mat_stride) * 4,
                               (uint64_t)op.plane_stride * 4};
  const uint32_t box[4] = {(uint32_t)BK, (uint32_t)box_rows, 1, 2};
  return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, op.base, dims, strides, box);
}

template <int EPI, int BN>
int launch_epi_bn(const CUtensorMap& ma, const CUtensorMap& mb, const TcKernelArgs& k, int mats, cudaStream_t st) {
  auto kern = tc_gemm_kernel<EPI, BN>;
  static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcTile<BN>::SMEM);
  if (attr != cudaSuccess) return cuda_status(attr, "tc_gemm smem attribute");
  const dim3 grid((k.N + BN - 1) / BN, (k.M + BM - 1) / BM, mats);
  return cuda_status(launch_pdl(true, kern, grid, dim3(TC_THREADS), (size_t)TcTile<BN>::SMEM, st, ma, mb, k), "tc_gemm launch");
}

static inline int pick_bn(int M, int N, int mats) {
  const int64_t ctas128 = (int64_t)((N + 127) / 128) * ((M + BM - 1) / BM) * mats;
  return (N > 128 && ctas128 > kSMs) ? 256 : 128;
}

template <int EPI>
int launch_epi(const TcGemm& g, const TcKernelArgs& k, cudaStream_t st) {
  const int bn = pick_bn(k.M, k.N, g.mats);
  CUtensorMap ma, mb;
  if (int e = operand_map(&ma, g.A, g.mats, BM)) return e;
  if (int e = operand_map(&mb, g.B, g.mats, bn)) return e;
  if (bn == 256) return launch_epi_bn<EPI, 256>(ma, mb, k, g.mats, st);
  return launch_epi_bn<EPI, 128>(ma, mb, k, g.mats, st);
}

}  // namespace

bool tc_gemm_supported(int M, int N, int K) {
  return M > 0 && N > 0 && K > 0 && M % 4 == 0 && N % 4 == 0 && K % 4 == 0 && tc5::encode_tiled_fn() != nullptr;
}

int tc_gemm_launch(const TcGemm& g, cudaStream_t st) {
  const int M = g.A.rows, N = g.B.rows, K = g.A.cols;
  NHMC_CHECK_ARG(g.B.cols == K && tc_gemm_supported(M, N, K), "tc_gemm: unsupported shape");
  NHMC_CHECK_ARG(aligned16(g.A.base) && aligned16(g.B.base) && aligned16(g.out_hi) && (!g.out_lo || aligned16(g.out_lo)),
                 "tc_gemm: operands must be 16-byte aligned");
  NHMC_CHECK_ARG(g.A.plane_stride % 4 == 0 && g.B.plane_stride % 4 == 0 && g.A.mat_stride % 4 == 0 && g.B.mat_stride % 4 == 0 &&
                 g.out_mat_stride % 4 == 0, "tc_gemm: strides must be multiples of 16 bytes");
  TcKernelArgs k{};
  k.M = M; k.N = N; k.K = K; k.channels = g.channels > 0 ? g.channels : 1;
  k.a_shared = g.A.mat_stride == 0; k.b_shared = g.B.mat_stride == 0;
  k.epi = g.epi; k.out_transposed = g.out_transposed; k.no_clamp = g.no_clamp;
  k.out_hi = g.out_hi; k.out_lo = g.out_lo; k.out_mat_stride = g.out_mat_stride;
  k.gain = g.gain; k.y = g.y; k.xhat = g.xhat; k.loss = g.loss; k.ws = g.ws;
  switch (g.epi) {
    case TC_EPI_NONE: return launch_epi<TC_EPI_NONE>(g, k, st);
    case TC_EPI_GAIN: NHMC_CHECK_ARG(g.gain, "gain required"); return launch_epi<TC_EPI_GAIN>(g, k, st);
    case TC_EPI_RESID:
      NHMC_CHECK_ARG(g.y && g.loss && ((N + 127) / 128) * ((M + BM - 1) / BM) * k.channels <= kMaxBlocksPerChain, "residual epilogue arguments");
      return launch_epi<TC_EPI_RESID>(g, k, st);
    case TC_EPI_GRAD: NHMC_CHECK_ARG(g.xhat, "xhat required"); return launch_epi<TC_EPI_GRAD>(g, k, st);
    case TC_EPI_SPEC:
      NHMC_CHECK_ARG(g.y && g.gain && g.loss && ((N + 127) / 128) * ((M + BM - 1) / BM) * k.channels <= kMaxBlocksPerChain, "spectral residual epilogue arguments");
      return launch_epi<TC_EPI_SPEC>(g, k, st);
  }
  set_error("tc_gemm: unknown epilogue %d", g.epi);
  return NHMC_ERR_ARG;
}

int tc_split_launch(const float* x, float* hi, float* lo, int64_t n, int clamp, cudaStream_t st) {
  NHMC_CHECK_ARG(n % 4 == 0 && aligned16(x) && aligned16(hi) && aligned16(lo), "split needs 16-byte aligned buffers and n % 4 == 0");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  tc_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, hi, lo, n4, clamp);
  return NHMC_LAUNCH_CHECK();
}

int tc_split_table_launch(const float* t, float* hi, float* lo, int rows, int cols, int transpose, cudaStream_t st) {
  tc_split_table_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, st>>>(t, hi, lo, rows, cols, transpose);
  return NHMC_LAUNCH_CHECK();
}

}  // namespace nhmc
