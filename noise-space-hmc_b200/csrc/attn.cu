// K10b: fused attention forward  out = softmax(scale * q k^T) v  for the score-network passes that keep no gradient
// state (the latent U-Net of hmc_latent: three forwards per leapfrog step, no VJP; QKVAttentionLegacy,
// ldm/modules/diffusionmodules/openaimodel.py:545-575 == guided_diffusion/unet_ffhq.py:408-425).  The unfused path
// writes the [B*heads, T, T] logits, reads and rewrites them in the softmax kernel and reads them again in the P V
// product: 4 passes over 235 MB per block at T = 1024 with 14 heads of 32 channels.  Here they never leave the SM.
//
// One CTA = 128 query rows of one (sample, head); tcgen05, warp-specialised:
//   warp 0 (one lane)  TMA producer : Q tile once; K tiles through a 3-deep, V tiles through a 2-deep ring
//   warp 1 (one lane)  MMA issuer   : S = Q K_j^T in two 64-key halves (M128 N64, K = head dim) -> TMEM columns [0,64) / [64,128)
//                                     O_half += P V per half (M128 N = head dim, K = 64)       -> TMEM columns [128, 128 + 2 head dim)
//   warps 2-5, 6-9     softmax      : two warpgroups, one per key half (lower / upper 64 keys of every tile), each with its
//                                     own running maximum, sum and accumulator (O_lo / O_hi, TMEM columns [128, 128 + 2 head dim)),
//                                     merged once in the epilogue.  Thread = query row (tcgen05.ld 32x32b: no shuffles).  ONE
//                                     pass over the keys: p = 2^(c s - c m_used) where m_used follows the running maximum
//                                     lazily -- only when that has grown by more than 2^8 are O (tcgen05.ld -> scale ->
//                                     tcgen05.st) and l = sum p rescaled, a warp-uniform and rare event -- so every
//                                     exponential is evaluated once and Q K^T is computed once.  P goes as bf16 into shared
//                                     memory in the 128-byte-swizzled K-major layout the second MMA reads.  A TMEM half goes
//                                     back to the MMA issuer as soon as its values have been read for the second time
//                                     (maximum, then exponentials, 32 columns at a time: 90 registers, two CTAs per SM).
// The exponentials are the bound of this kernel: MUFU.EX2 issues one warp instruction per 8 clocks per scheduler, 16 results
// per clock per SM (tools/exp/exp_mufu.cu, profiles/r2_attn_ncu_summary.md).
// Layouts (chosen so that every MMA operand is K-major and the producing / consuming GEMMs need no transpose):
//   qk  [B, T, heads, 2, ch]  token-major, q then k of a head adjacent (the qkv projection's own channel order)
//   v   [B, heads*ch, T]      channel-major
//   out [B, T, heads*ch]      token-major
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc5.cuh"

namespace nhmc {
using namespace tc5;

namespace {

constexpr int AT_BM = 128, AT_BN = 128, AT_THREADS = 320;   // producer warp, MMA warp, two softmax warpgroups
constexpr int AT_QBYTES = AT_BM * 128;            // 128 rows x one 128-byte swizzle row (ch = 32: upper half zero-filled by TMA)
constexpr int AT_KBYTES = AT_BN * 128;
constexpr int AT_PBYTES = 2 * AT_BM * 128;        // two 64-key chunks of [128 rows][128 bytes]
#ifndef NHMC_ATTN_KSTAGES
#define NHMC_ATTN_KSTAGES 3
#endif
constexpr int AT_KS = NHMC_ATTN_KSTAGES;

template <int CH> struct AttnSmem {
  static constexpr int VCHUNK = CH * 128;         // [CH rows][64 keys] bf16
  static constexpr int VBYTES = 2 * VCHUNK;
  static constexpr int KS = CH == 32 ? AT_KS : 2;
  static constexpr int OFF_K = AT_QBYTES, OFF_V = OFF_K + KS * AT_KBYTES, OFF_P = OFF_V + 2 * VBYTES, OFF_BAR = OFF_P + AT_PBYTES;
  // no alignment slack: two CTAs of the ch = 32 kernel must fit one SM (2 x (112.25 KiB + 1 KiB reserved) <= 228 KiB); the
  // dynamic shared memory window of a kernel without static shared memory starts 1024-byte aligned, checked at run time
  static constexpr int TOTAL = OFF_BAR + 256;
};

struct AttnArgs {
  void* out;
  int T, heads;
  float c;          // scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool BF16> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Barrier indices.  S is produced in 64-key halves (TMEM columns [0,64) / [64,128)), each with its own full / empty pair; P has
// one 64-key shared-memory chunk per half with a full barrier (softmax -> MMA) and a "P V done" barrier (MMA commit -> softmax);
// O accumulates in TMEM columns [128, 128 + CH).
enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 5, B_VFULL = 9, B_VEMPTY = 11, B_SFULL = 13, B_SEMPTY = 15, B_PFULL = 17, B_PVDONE = 19,
       B_OFULL = 21, B_MERGE = 22, B_COUNT = 23 };
constexpr float AT_RESCALE_LOG2 = 8.0f;   // O and l are rescaled only when a row maximum grew by more than 2^8: p <= 256 in between

template <int CH, bool BF16>
__global__ void __launch_bounds__(AT_THREADS, CH == 32 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  using L = AttnSmem<CH>;
  constexpr int KS = L::KS;
  static_assert(KS <= 4, "barrier table holds four K stages");
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int nkv = (a.T + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    if (smem_addr(sm) & 1023u) asm volatile("trap;");             // swizzled tiles need 1024-byte aligned bases
    prefetch_tensormap(&mapQK);
    prefetch_tensormap(&mapV);
    for (int i = 0; i < B_COUNT; ++i) {                            // S "empty" and P "full": all 128 softmax threads arrive
      const bool by_softmax = (i >= B_SEMPTY && i < B_SEMPTY + 2) || (i >= B_PFULL && i < B_PFULL + 2) || i == B_MERGE;
      mbar_init(bars + i, by_softmax ? 128 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 128;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bars + B_QFULL, AT_QBYTES);
      tma_load_4d(sm, &mapQK, bars + B_QFULL, 0, 2 * h, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % KS, vs = j & 1;
        mbar_wait(bars + B_KEMPTY + ks, ((j / KS) & 1) ^ 1);
        mbar_arrive_expect_tx(bars + B_KFULL + ks, AT_KBYTES);
        tma_load_4d(sm + L::OFF_K + ks * AT_KBYTES, &mapQK, bars + B_KFULL + ks, 0, 2 * h + 1, j * AT_BN, b);
        mbar_wait(bars + B_VEMPTY + vs, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bars + B_VFULL + vs, L::VBYTES);
        unsigned char* vd = sm + L::OFF_V + vs * L::VBYTES;
        tma_load_3d(vd, &mapV, bars + B_VFULL + vs, j * AT_BN, h * CH, b);
        tma_load_3d(vd + L::VCHUNK, &mapV, bars + B_VFULL + vs, j * AT_BN + 64, h * CH, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(BF16 ? 1 : 0, AT_BM, 64);
      constexpr uint32_t idesc_o = umma_idesc(BF16 ? 1 : 0, AT_BM, CH);
      const uint64_t dq = umma_desc_sw128(smem_addr(sm));
      const uint32_t p_addr = smem_addr(sm + L::OFF_P);
      // O += P(j, hf) V(j, hf): 64 keys = four K16 slices
      auto pv = [&](int j, int hf) {
        const int vs = j & 1;
        if (hf == 0) mbar_wait(bars + B_VFULL + vs, (j >> 1) & 1);
        mbar_wait(bars + B_PFULL + hf, j & 1);
        tc_fence_after();
        const uint64_t dp = umma_desc_sw128(p_addr + hf * (AT_BM * 128));
        const uint64_t dv = umma_desc_sw128(smem_addr(sm + L::OFF_V + vs * L::VBYTES) + hf * L::VCHUNK);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)                            // 16 keys = 32 bytes along the swizzled row
          umma_f16(tmem_o + hf * CH, dp + (uint64_t)(kk * 2), dv + (uint64_t)(kk * 2), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(bars + B_PVDONE + hf);
        if (hf == 1) umma_commit(bars + B_VEMPTY + vs);
      };
      mbar_wait(bars + B_QFULL, 0);
      // issue order = the order in which the softmax warps free / fill things: S(j, lo), P V(j-1, lo), S(j, hi), P V(j-1, hi)
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % KS;
        mbar_wait(bars + B_KFULL + ks, (j / KS) & 1);
        const uint32_t k_addr = smem_addr(sm + L::OFF_K + ks * AT_KBYTES);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(bars + B_SEMPTY + hf, (j & 1) ^ 1);
          tc_fence_after();
          const uint64_t dk = umma_desc_sw128(k_addr + hf * (64 * 128));
#pragma unroll
          for (int kk = 0; kk < CH / 16; ++kk)
            umma_f16(tmem + hf * 64, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(bars + B_SFULL + hf);
          if (hf == 1) umma_commit(bars + B_KEMPTY + ks);
          if (j > 0) pv(j - 1, hf);
        }
      }
      pv(nkv - 1, 0);
      pv(nkv - 1, 1);
      umma_commit(bars + B_OFULL);
    }
  } else {
    // Two softmax warpgroups: warps 2-5 take the lower 64 keys of every tile, warps 6-9 the upper 64, each with its own running
    // maximum, sum and accumulator (O_lo / O_hi in TMEM) -- two independent streams, merged once at the end.  Four softmax
    // warps per scheduler (two CTAs per SM) instead of two: the FFMA -> MUFU -> FADD / pack / store chain of one warp no
    // longer leaves the special-function pipe idle.  Thread = query row within its warpgroup.
    const int hf = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tmem_s = tmem + lane_base + (uint32_t)(hf * 64);
    const uint32_t tmem_oh = tmem_o + lane_base + (uint32_t)(hf * CH);
    const bool ragged = (a.T % AT_BN) != 0;
    // m_used is the maximum the stored exponentials refer to: it follows the true running maximum lazily (only when that has
    // grown by more than 2^AT_RESCALE_LOG2), and then O (tensor memory) and l are rescaled -- a warp-uniform, rare event.
    float m_used = -INFINITY;
    float l4[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned char* pc = sm + L::OFF_P + hf * (AT_BM * 128) + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < nkv; ++j) {
      const int valid = a.T - j * AT_BN - hf * 64;                 // keys of this half that exist (<= 0: none)
      const bool mask = ragged && j == nkv - 1;
      mbar_wait(bars + B_SFULL + hf, j & 1);
      tc_fence_after();
      // pass A over the 64 logits of the unit: the maximum (32 at a time: registers)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t r[32];
        tmem_ld32(tmem_s + (uint32_t)(g * 32), r);
        tmem_ld_wait();
        if (mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (g * 32 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        }
      }
      const float m_new = fmaxf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), m_used);
      const bool grow = (m_new - m_used) * a.c > AT_RESCALE_LOG2;        // also true from -inf
      if (j == 0) {
        m_used = m_new;                                            // nothing accumulated yet
      } else if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1);              // every P V into this accumulator has landed
        tc_fence_after();
        const float f = grow ? ex2((m_used - m_new) * a.c) : 1.0f; // m_used = -inf -> 0 (all keys so far were masked)
        if (grow) m_used = m_new;
#pragma unroll
        for (int g = 0; g < CH / 32; ++g) {
          uint32_t o[32];
          tmem_ld32(tmem_oh + (uint32_t)(g * 32), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st32(tmem_oh + (uint32_t)(g * 32), o);
        }
        tmem_st_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) l4[i] *= f;
      }
      const float mc = m_used * a.c;
      // the P V that read this chunk last (previous tile) must be done before it is overwritten
      if (j > 0) mbar_wait(bars + B_PVDONE + hf, (j - 1) & 1);
      // pass B: exponentials, 32 at a time; the TMEM half goes back to the MMA issuer after the second load
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t r[32];
        tmem_ld32(tmem_s + (uint32_t)(g * 32), r);
        tmem_ld_wait();
        if (g == 1) { tc_fence_before(); mbar_arrive(bars + B_SEMPTY + hf); }
        float p[32];
        if (mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = g * 32 + i < valid ? ex2(fmaf(__uint_as_float(r[i]), a.c, -mc)) : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = ex2(fmaf(__uint_as_float(r[i]), a.c, -mc));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) l4[i & 3] += p[i];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int chunk = g * 4 + c4;                           // 16-byte chunk of this row's 128 bytes
          uint4 w;
          w.x = pack2<BF16>(p[c4 * 8], p[c4 * 8 + 1]);
          w.y = pack2<BF16>(p[c4 * 8 + 2], p[c4 * 8 + 3]);
          w.z = pack2<BF16>(p[c4 * 8 + 4], p[c4 * 8 + 5]);
          w.w = pack2<BF16>(p[c4 * 8 + 6], p[c4 * 8 + 7]);
          *reinterpret_cast<uint4*>(pc + ((chunk ^ sw) << 4)) = w;
        }
      }
      tc_fence_before();                                          // orders a rescale's tcgen05.st before the MMA that follows the arrive
      fence_proxy_async();                                        // generic-proxy stores -> visible to the MMA's async reads
      mbar_arrive(bars + B_PFULL + hf);
    }
    // ---- merge the two halves and write  (O_lo f_lo + O_hi f_hi) / (l_lo f_lo + l_hi f_hi)  -> out[b, t, h*CH + c] ----
    // The upper warpgroup hands (m, l) of its rows over through the Q tile's shared memory: every S product has been read by
    // now (the last one by this very warpgroup), so nothing reads Q any more.
    const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
    float2* xch = reinterpret_cast<float2*>(sm);
    if (hf == 1) {
      xch[row] = make_float2(m_used, l);
      __threadfence_block();
      mbar_arrive(bars + B_MERGE);
    } else {
      mbar_wait(bars + B_OFULL, 0);
      tc_fence_after();
      mbar_wait(bars + B_MERGE, 0);
      const float2 other = xch[row];
      const float m_all = fmaxf(m_used, other.x);
      const float f_lo = ex2((m_used - m_all) * a.c);
      const float f_hi = other.x == -INFINITY ? 0.f : ex2((other.x - m_all) * a.c);
      const float inv = __fdividef(1.0f, l * f_lo + other.y * f_hi);
      const float s_lo = f_lo * inv, s_hi = f_hi * inv;
      const int t = q0 + row;
      const int C = a.heads * CH;
      unsigned char* orow = reinterpret_cast<unsigned char*>(a.out) + (((int64_t)b * a.T + t) * C + (int64_t)h * CH) * 2;
#pragma unroll
      for (int g = 0; g < CH / 32; ++g) {
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem_o + lane_base + (uint32_t)(g * 32), r0);
        tmem_ld32(tmem_o + lane_base + (uint32_t)(CH + g * 32), r1);
        tmem_ld_wait();
        if (t < a.T) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r0[c4 * 8 + i]) * s_lo + __uint_as_float(r1[c4 * 8 + i]) * s_hi;
            uint4 w;
            w.x = pack2<BF16>(v[0], v[1]);
            w.y = pack2<BF16>(v[2], v[3]);
            w.z = pack2<BF16>(v[4], v[5]);
            w.w = pack2<BF16>(v[6], v[7]);
            *reinterpret_cast<uint4*>(orow + g * 64 + c4 * 16) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int CH, bool BF16>
int attn_launch(const void* qk, const void* v, void* out, int batch, int T, int heads, float scale, unsigned flags, cudaStream_t st) {
  const CUtensorMapDataType dt = BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap mqk, mv;
  {
    const uint64_t dims[4] = {(uint64_t)CH, (uint64_t)(2 * heads), (uint64_t)T, (uint64_t)batch};
    const uint64_t strides[3] = {(uint64_t)CH * 2, (uint64_t)heads * 2 * CH * 2, (uint64_t)T * heads * 2 * CH * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)AT_BM, 1};          // CH = 32: columns 32..63 are out of bounds -> zero
    if (int e = make_tensor_map(&mqk, dt, 4, qk, dims, strides, box)) return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)T, (uint64_t)(heads * CH), (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)T * 2, (uint64_t)heads * CH * T * 2};
    const uint32_t box[3] = {64, (uint32_t)CH, 1};
    if (int e = make_tensor_map(&mv, dt, 3, v, dims, strides, box)) return e;
  }
  auto kern = attn_fwd_kernel<CH, BF16>;
  static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<CH>::TOTAL);
  if (attr != cudaSuccess) return cuda_status(attr, "attn_fwd smem attribute");
  AttnArgs a{out, T, heads, scale * 1.4426950408889634f};
  const dim3 grid((T + AT_BM - 1) / AT_BM, batch * heads);
  return cuda_status(launch_pdl(!(flags & NHMC_F_NO_PDL), kern, grid, dim3(AT_THREADS), (size_t)AttnSmem<CH>::TOTAL, st, mqk, mv, a),
                     "attn_fwd launch");
}

}  // namespace
}  // namespace nhmc

extern "C" int nhmc_attn_fwd(int dtype, const void* qk, const void* v, void* out, int batch, int tokens, int heads, int head_dim,
                             float scale, unsigned int flags, nhmc_stream_t stream) {
  using namespace nhmc;
  NHMC_CHECK_ARG(qk && v && out && batch > 0 && tokens > 0 && heads > 0, "bad arguments");
  NHMC_CHECK_ARG(aligned16(qk) && aligned16(v) && aligned16(out), "operands must be 16-byte aligned");
  NHMC_CHECK_ARG(tokens % 8 == 0, "tokens must be a multiple of 8 (16-byte rows of v)");
  NHMC_CHECK_ARG((int64_t)batch * heads <= 65535, "batch * heads exceeds the grid limit");
  if ((dtype != NHMC_BF16 && dtype != NHMC_F16) || (head_dim != 32 && head_dim != 64)) {
    set_error("nhmc_attn_fwd: bf16 / f16 with head_dim 32 or 64 only (dtype %d, head_dim %d)", dtype, head_dim);
    return NHMC_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NHMC_BF16)
    return head_dim == 32 ? attn_launch<32, true>(qk, v, out, batch, tokens, heads, scale, flags, st)
                          : attn_launch<64, true>(qk, v, out, batch, tokens, heads, scale, flags, st);
  return head_dim == 32 ? attn_launch<32, false>(qk, v, out, batch, tokens, heads, scale, flags, st)
                        : attn_launch<64, false>(qk, v, out, batch, tokens, heads, scale, flags, st);
}
