// K4/K5: measurement operators fused with residual, loss reduction and adjoint.
// Replaces obs_functions/Hfuncs.py {Inpainting :119-154, SuperResolution :180-234, HDR :406-445}
// plus `loss = sum((y_0 - H(xt.clip(-1,1)))**2)` and its autograd (main_sampling.py:709-711).
// Deblurring2D and PhaseRetrieval live in deblur.cu / phase.cu and are dispatched from here.
//
// One pass per gradient evaluation: read x_hat (N) + y (M), write g (N) [+ clamped image];
// loss accumulated in fp64 in-register, warp-shuffle + one block reduce + deterministic
// last-block finish (common.cuh).
#include "common.cuh"

namespace nhmc {

// defined in deblur.cu / phase.cu
int deblur_forward(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int deblur_adjoint(const nhmc_op*, const float*, float*, int, void*, cudaStream_t, bool pinv);
int deblur_loss_grad(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t);
size_t deblur_workspace_bytes(const nhmc_op*, int);
int sepmat_forward(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int sepmat_adjoint(const nhmc_op*, const float*, float*, int, void*, cudaStream_t, bool pinv);
int sepmat_loss_grad(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t);
size_t sepmat_workspace_bytes(const nhmc_op*, int);
size_t spectral_tables_bytes(const nhmc_op*);
int spectral_prepare_tables(const nhmc_op*, float*, cudaStream_t);
int deblur_spectral_measurement(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int deblur_loss_grad_spec(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t);
int cs_forward(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int cs_adjoint(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int cs_loss_grad(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t);
size_t cs_workspace_bytes(const nhmc_op*, int);
int phase_forward(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int phase_pinv(const nhmc_op*, const float*, float*, int, void*, cudaStream_t);
int phase_loss_grad(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t);
size_t phase_workspace_bytes(const nhmc_op*, int);
size_t phase_tables_bytes(const nhmc_op*);
int phase_prepare_tables(const nhmc_op*, void*, cudaStream_t);

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }
__device__ __forceinline__ float inr(float v, float lo, float hi) { return (v >= lo && v <= hi) ? 1.0f : 0.0f; }
// final clip of main_sampling.py:709 and its gradient mask; nc = 1 disables both (latent path)
__device__ __forceinline__ float clampx(float v, int nc) { return nc ? v : clamp1(v); }
__device__ __forceinline__ float maskx(float v, int nc) { return nc ? 1.0f : inr(v, -1.0f, 1.0f); }

// =====================================================================================
// Inpainting
// =====================================================================================
// y[b, 3k+c] = x[b, c, kept[k]]
__global__ void inpaint_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const int32_t* __restrict__ kept,
                                   int n_kept, int C, int64_t P) {
  const int b = blockIdx.y;
  const int64_t M = (int64_t)n_kept * C;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(j / C), c = (int)(j - (int64_t)k * C);
    y[(int64_t)b * M + j] = x[((int64_t)b * C + c) * P + kept[k]];
  }
}

// out[b, c, pix] = r[b, 3*rank(pix)+c] if kept else 0   (H^T = H^+ for a row-selection matrix)
__global__ void inpaint_adj_kernel(const float* __restrict__ r, float* __restrict__ out, const int32_t* __restrict__ p2k,
                                   int n_kept, int C, int64_t P) {
  const int b = blockIdx.y;
  const int64_t N = (int64_t)C * P, M = (int64_t)n_kept * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / P);
    const int64_t pix = e - (int64_t)c * P;
    const int k = p2k[pix];
    out[(int64_t)b * N + e] = k >= 0 ? r[(int64_t)b * M + (int64_t)k * C + c] : 0.0f;
  }
}

__global__ void __launch_bounds__(kThreads)
inpaint_loss_grad_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                         float* __restrict__ xc_out, const int32_t* __restrict__ p2k, int n_kept, int C, int64_t P,
                         double* loss, RedWs ws, int nc) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t N = (int64_t)C * P, M = (int64_t)n_kept * C;
  const int64_t n_vec = N / 4;  // P % 4 == 0 is checked by the host, so a float4 never straddles channels
  const float* xb = xh + (int64_t)b * N;
  const float* yb = y + (int64_t)b * M;
  double acc = 0.0;
  for (int64_t v = (int64_t)blk * kThreads + threadIdx.x; v < n_vec; v += (int64_t)nblk * kThreads) {
    const int64_t e0 = v * 4;
    const int c = (int)(e0 / P);
    const int64_t pix0 = e0 - (int64_t)c * P;
    const float4 xv = ld_ro(reinterpret_cast<const float4*>(xb + e0));
    const int4 kv = __ldg(reinterpret_cast<const int4*>(p2k + pix0));
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const int ks[4] = {kv.x, kv.y, kv.z, kv.w};
    float gs[4], cs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      cs[j] = clampx(xs[j], nc);
      gs[j] = 0.0f;
      if (ks[j] >= 0) {
        const float r = __fsub_rn(__ldg(yb + (int64_t)ks[j] * C + c), cs[j]);
        acc += (double)r * (double)r;
        gs[j] = -(2.0f * r) * maskx(xs[j], nc);
      }
    }
    st_stream(reinterpret_cast<float4*>(g + (int64_t)b * N + e0), make_float4(gs[0], gs[1], gs[2], gs[3]));
    if (xc_out) st_stream(reinterpret_cast<float4*>(xc_out + (int64_t)b * N + e0), make_float4(cs[0], cs[1], cs[2], cs[3]));
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish_par<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk, red)) loss[b] = vsum[0];
}

// =====================================================================================
// Super-resolution (r x r block average)
// =====================================================================================
__global__ void sr_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int D, int r) {
  const int b = blockIdx.y, yd = D / r;
  const int64_t M = (int64_t)C * yd * yd;
  const float inv = 1.0f / (float)(r * r);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(j / ((int64_t)yd * yd));
    const int rem = (int)(j - (int64_t)c * yd * yd);
    const int ci = rem / yd, cj = rem - ci * yd;
    const float* src = x + (((int64_t)b * C + c) * D + (int64_t)ci * r) * D + (int64_t)cj * r;
    float s = 0.f;
    for (int u = 0; u < r; ++u) {
      float rs = 0.f;
      for (int w = 0; w < r; ++w) rs += src[(int64_t)u * D + w];
      s += rs;
    }
    y[(int64_t)b * M + j] = s * inv;
  }
}

// out = upsample(r) * mul   (mul = 1/r^2 for H^T, 1 for H^+)
__global__ void sr_up_kernel(const float* __restrict__ rr, float* __restrict__ out, int C, int D, int r, float mul) {
  const int b = blockIdx.y, yd = D / r;
  const int64_t N = (int64_t)C * D * D, M = (int64_t)C * yd * yd;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / ((int64_t)D * D));
    const int rem = (int)(e - (int64_t)c * D * D);
    const int i = rem / D, j = rem - i * D;
    out[(int64_t)b * N + e] = rr[(int64_t)b * M + ((int64_t)c * yd + i / r) * yd + j / r] * mul;
  }
}

// One CTA owns a strip of TR rows x D columns of one (chain, channel): stage raw x in shared memory,
// reduce r x r cells in a fixed order, then write the gradient of every pixel of the strip.
__global__ void __launch_bounds__(kThreads)
sr_loss_grad_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                    float* __restrict__ xc_out, int C, int D, int r, int TR, double* loss, RedWs ws, int nc) {
  extern __shared__ float smem[];
  float* tile = smem;                    // [TR][D] raw x_hat
  float* rowsum = tile + (size_t)TR * D; // [TR][D/r]
  float* cellres = rowsum + (size_t)TR * (D / r);  // [TR/r][D/r] residual per cell
  const int b = blockIdx.y, strip = blockIdx.x, nstrips = gridDim.x;
  const int strips_per_ch = D / TR;
  const int c = strip / strips_per_ch, row0 = (strip - c * strips_per_ch) * TR;
  const int yd = D / r, cpr = D / r;  // cells per row
  const int64_t N = (int64_t)C * D * D, M = (int64_t)C * yd * yd;
  const float* src = xh + (int64_t)b * N + ((int64_t)c * D + row0) * D;
  const int n_vec = TR * D / 4;
  for (int v = threadIdx.x; v < n_vec; v += kThreads)
    reinterpret_cast<float4*>(tile)[v] = ld_ro(reinterpret_cast<const float4*>(src) + v);
  __syncthreads();
  for (int i = threadIdx.x; i < TR * cpr; i += kThreads) {
    const int row = i / cpr, cc = i - row * cpr;
    const float* p = tile + (size_t)row * D + (size_t)cc * r;
    float s = 0.f;
    for (int w = 0; w < r; ++w) s += clampx(p[w], nc);
    rowsum[i] = s;
  }
  __syncthreads();
  double acc = 0.0;
  const float inv = 1.0f / (float)(r * r);
  for (int i = threadIdx.x; i < (TR / r) * cpr; i += kThreads) {
    const int cr = i / cpr, cc = i - cr * cpr;
    float s = 0.f;
    for (int u = 0; u < r; ++u) s += rowsum[(cr * r + u) * cpr + cc];
    const float res = __fsub_rn(__ldg(y + (int64_t)b * M + ((int64_t)c * yd + (row0 / r + cr)) * yd + cc), s * inv);
    cellres[i] = res;
    acc += (double)res * (double)res;
  }
  __syncthreads();
  float* gdst = g + (int64_t)b * N + ((int64_t)c * D + row0) * D;
  float* cdst = xc_out ? xc_out + (int64_t)b * N + ((int64_t)c * D + row0) * D : nullptr;
  for (int v = threadIdx.x; v < n_vec; v += kThreads) {
    const int e = v * 4, row = e / D, col = e - row * D;
    const float4 xv = reinterpret_cast<const float4*>(tile)[v];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    float gs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float res = cellres[(row / r) * cpr + (col + j) / r];
      gs[j] = -(2.0f * res) * inv * maskx(xs[j], nc);
    }
    st_stream(reinterpret_cast<float4*>(gdst) + v, make_float4(gs[0], gs[1], gs[2], gs[3]));
    if (cdst) st_stream(reinterpret_cast<float4*>(cdst) + v, make_float4(clampx(xs[0], nc), clampx(xs[1], nc), clampx(xs[2], nc), clampx(xs[3], nc)));
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, strip, nstrips)) loss[b] = vsum[0];
}

// r in {4, 8, 16}: no shared memory, no block barrier.  A cell is covered by LPC = (r/4) x RS lanes: r/4 float4 across a row,
// RS groups of r/RS = 4 rows; a lane keeps its 4 float4 in registers (independent 16-byte loads in flight), the cell sum is
// 4 row sums per lane + log2(LPC) shuffles, and the same lane writes the gradient (and the clamped copy) of the r x 4 pixels it loaded.  A warp covers
// 32 / LPC neighbouring cells of one cell row: every load / store instruction touches whole 128-byte lines.
template <int R, int RS>
__global__ void __launch_bounds__(kThreads)
sr_loss_grad_reg_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                        float* __restrict__ xc_out, int C, int D, double* loss, RedWs ws, int nc) {
  constexpr int LPC = (R / 4) * RS;                // lanes per cell: R/4 across a row x RS groups of R/RS rows
  constexpr int NR = R / RS;                       // rows a lane keeps in registers
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int yd = D / R;
  const int64_t N = (int64_t)C * D * D, M = (int64_t)C * yd * yd;
  const int64_t n_items = M * LPC;                 // one item = (cell, lane-in-cell)
  const float inv = 1.0f / (float)(R * R);
  double acc = 0.0;
  const float* xb = xh + (int64_t)b * N;
  for (int64_t it = (int64_t)blk * kThreads + threadIdx.x; it < ((n_items + 31) / 32) * 32; it += (int64_t)nblk * kThreads) {
    const bool ok = it < n_items;                  // whole warps stay in the loop for the shuffles
    const int64_t cell = ok ? it / LPC : 0;
    const int q = (int)(it % LPC);
    const int qc = q % (R / 4), qr = q / (R / 4);  // column float4 and row group of this lane inside the cell
    const int c = (int)(cell / ((int64_t)yd * yd));
    const int rem = (int)(cell - (int64_t)c * yd * yd);
    const int ci = rem / yd, cj = rem - ci * yd;
    const int64_t off = ((int64_t)c * D + (int64_t)ci * R + qr * NR) * D + (int64_t)cj * R + qc * 4;
    float4 v[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) v[u] = ok ? ld_ro(reinterpret_cast<const float4*>(xb + off + (int64_t)u * D)) : make_float4(0, 0, 0, 0);
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      float rs = clampx(v[u].x, nc) + clampx(v[u].y, nc);
      rs += clampx(v[u].z, nc); rs += clampx(v[u].w, nc);
      s += rs;
    }
#pragma unroll
    for (int o = LPC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float res = ok ? __fsub_rn(__ldg(y + (int64_t)b * M + cell), s * inv) : 0.f;
    if (ok && q == 0) acc += (double)res * (double)res;
    if (ok) {
      const float gv = -(2.0f * res) * inv;
      float* gd = g + (int64_t)b * N + off;
      float* cd = xc_out ? xc_out + (int64_t)b * N + off : nullptr;
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        st_stream(reinterpret_cast<float4*>(gd + (int64_t)u * D),
                  make_float4(gv * maskx(v[u].x, nc), gv * maskx(v[u].y, nc), gv * maskx(v[u].z, nc), gv * maskx(v[u].w, nc)));
        if (cd) st_stream(reinterpret_cast<float4*>(cd + (int64_t)u * D),
                          make_float4(clampx(v[u].x, nc), clampx(v[u].y, nc), clampx(v[u].z, nc), clampx(v[u].w, nc)));
      }
    }
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish_par<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk, red)) loss[b] = vsum[0];
}

// =====================================================================================
// HDR
// =====================================================================================
__global__ void hdr_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = clamp1(x[e] * 2.0f);   // x / 0.5 (Hfuncs.py:423) is exactly 2 x in binary floating point
}
__global__ void copy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) y[e] = x[e];
}

__global__ void __launch_bounds__(kThreads)
hdr_loss_grad_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                     float* __restrict__ xc_out, int64_t N, double* loss, RedWs ws, int nc) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t n_vec = N / 4;
  const float4* xb = reinterpret_cast<const float4*>(xh + (int64_t)b * N);
  const float4* yb = reinterpret_cast<const float4*>(y + (int64_t)b * N);
  double acc = 0.0;
  constexpr int U = 4;
  const int64_t stride = (int64_t)nblk * kThreads;
  for (int64_t v0 = (int64_t)blk * kThreads + threadIdx.x; v0 < n_vec; v0 += stride * U) {
    float4 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      xv[u] = yv[u] = make_float4(0, 0, 0, 0);
      if (v < n_vec) { xv[u] = ld_ro(xb + v); yv[u] = ld_ro(yb + v); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= n_vec) continue;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
      float gs[4], cs[4];
      double a4 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cs[j] = clampx(xs[j], nc);
        const float h = cs[j] * 2.0f;      // x / 0.5 (Hfuncs.py:423) == 2 x exactly: no IEEE division sequence needed
        const float r = __fsub_rn(ys[j], clamp1(h));
        a4 += (double)r * (double)r;
        // -2 r * d clip(h)/dh * dh/dx * d clamp(x)/dx
        gs[j] = -(2.0f * r) * inr(h, -1.0f, 1.0f) * 2.0f * maskx(xs[j], nc);
      }
      acc += a4;
      st_stream(reinterpret_cast<float4*>(g + (int64_t)b * N) + v, make_float4(gs[0], gs[1], gs[2], gs[3]));
      if (xc_out) st_stream(reinterpret_cast<float4*>(xc_out + (int64_t)b * N) + v, make_float4(cs[0], cs[1], cs[2], cs[3]));
    }
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish_par<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk, red)) loss[b] = vsum[0];
}

// =====================================================================================
// Denoising (H = identity, Hfuncs.py:157-177): same fused pass as HDR without the x/0.5 clip
// =====================================================================================
__global__ void __launch_bounds__(kThreads)
denoise_loss_grad_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                     float* __restrict__ xc_out, int64_t N, double* loss, RedWs ws, int nc) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t n_vec = N / 4;
  const float4* xb = reinterpret_cast<const float4*>(xh + (int64_t)b * N);
  const float4* yb = reinterpret_cast<const float4*>(y + (int64_t)b * N);
  double acc = 0.0;
  constexpr int U = 4;
  const int64_t stride = (int64_t)nblk * kThreads;
  for (int64_t v0 = (int64_t)blk * kThreads + threadIdx.x; v0 < n_vec; v0 += stride * U) {
    float4 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      xv[u] = yv[u] = make_float4(0, 0, 0, 0);
      if (v < n_vec) { xv[u] = ld_ro(xb + v); yv[u] = ld_ro(yb + v); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= n_vec) continue;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
      float gs[4], cs[4];
      double a4 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cs[j] = clampx(xs[j], nc);
        const float r = __fsub_rn(ys[j], cs[j]);
        a4 += (double)r * (double)r;
        gs[j] = -(2.0f * r) * maskx(xs[j], nc);
      }
      acc += a4;
      st_stream(reinterpret_cast<float4*>(g + (int64_t)b * N) + v, make_float4(gs[0], gs[1], gs[2], gs[3]));
      if (xc_out) st_stream(reinterpret_cast<float4*>(xc_out + (int64_t)b * N) + v, make_float4(cs[0], cs[1], cs[2], cs[3]));
    }
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish_par<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk, red)) loss[b] = vsum[0];
}

// =====================================================================================
// Colorization (Hfuncs.py:655-695): y[b, pix] = sum_c w_c x[b, c, pix]; H^T r = a_c r; H^+ y = p_c y
// (w, a, p: products of the 1x3 SVD factors, built by the caller with the reference's torch.svd call)
// =====================================================================================
struct Coef3 { float v[3]; };
__global__ void color_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t P, Coef3 w) {
  const int b = blockIdx.y;
  const float* xb = x + (int64_t)b * 3 * P;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
    y[(int64_t)b * P + p] = fmaf(w.v[2], xb[2 * P + p], fmaf(w.v[1], xb[P + p], w.v[0] * xb[p]));
}
__global__ void color_adj_kernel(const float* __restrict__ r, float* __restrict__ out, int64_t P, Coef3 a) {
  const int b = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const float v = r[(int64_t)b * P + p];
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((int64_t)b * 3 + c) * P + p] = a.v[c] * v;
  }
}
__global__ void __launch_bounds__(kThreads)
color_loss_grad_kernel(const float* __restrict__ xh, const float* __restrict__ y, float* __restrict__ g,
                       float* __restrict__ xc_out, int64_t P, Coef3 w, Coef3 a, double* loss, RedWs ws, int nc) {
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t n_vec = P / 4;
  const float* xb = xh + (int64_t)b * 3 * P;
  double acc = 0.0;
  for (int64_t v = (int64_t)blk * kThreads + threadIdx.x; v < n_vec; v += (int64_t)nblk * kThreads) {
    float xs[3][4], cs[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 t = ld_ro(reinterpret_cast<const float4*>(xb + c * P) + v);
      xs[c][0] = t.x; xs[c][1] = t.y; xs[c][2] = t.z; xs[c][3] = t.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) cs[c][j] = clampx(xs[c][j], nc);
    }
    const float4 yv = ld_ro(reinterpret_cast<const float4*>(y + (int64_t)b * P) + v);
    const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
    float rs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yh = fmaf(w.v[2], cs[2][j], fmaf(w.v[1], cs[1][j], w.v[0] * cs[0][j]));
      rs[j] = __fsub_rn(ys[j], yh);
      acc += (double)rs[j] * (double)rs[j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gs[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gs[j] = -(2.0f * (a.v[c] * rs[j])) * maskx(xs[c][j], nc);
      st_stream(reinterpret_cast<float4*>(g + ((int64_t)b * 3 + c) * P) + v, make_float4(gs[0], gs[1], gs[2], gs[3]));
      if (xc_out) st_stream(reinterpret_cast<float4*>(xc_out + ((int64_t)b * 3 + c) * P) + v, make_float4(cs[c][0], cs[c][1], cs[c][2], cs[c][3]));
    }
  }
  __shared__ double red[32];
  double vsum[1] = {acc};
  block_sum<double, 1>(vsum, red);
  if (grid_finish_par<1>(vsum, ws.partials + (int64_t)b * kMaxBlocksPerChain, ws.tickets + b, blk, nblk, red)) loss[b] = vsum[0];
}
static inline Coef3 coef3(const nhmc_op* op, int first) { Coef3 c; for (int i = 0; i < 3; ++i) c.v[i] = op->coef[first + i]; return c; }

static inline int sr_tile_rows(int D, int r) {
  int TR = r;
  while (TR < 16 && TR * 2 <= D) TR *= 2;  // ~1024 float4 per CTA at D = 256
  return TR;
}

static int check_op(const nhmc_op* op) {
  NHMC_CHECK_ARG(op, "null operator");
  NHMC_CHECK_ARG(op->channels > 0 && op->img_dim > 0, "bad operator geometry");
  switch (op->kind) {
    case NHMC_OP_INPAINT:
      NHMC_CHECK_ARG(op->pix_to_kept && op->n_kept >= 0 && (op->n_kept == 0 || op->kept_pix), "inpaint tables missing");
      break;
    case NHMC_OP_SR:
      NHMC_CHECK_ARG(op->ratio > 0 && op->img_dim % op->ratio == 0, "ratio must divide img_dim");
      break;
    case NHMC_OP_HDR:
      break;
    case NHMC_OP_DEBLUR2D:
      NHMC_CHECK_ARG(op->U1 && op->V1 && op->U2 && op->V2 && op->gain, "deblur tables missing");
      break;
    case NHMC_OP_PHASE:
      NHMC_CHECK_ARG(op->pad >= 0, "bad pad");
      break;
    case NHMC_OP_SEPMAT:
      NHMC_CHECK_ARG(op->ratio > 0 && op->img_dim % op->ratio == 0 && op->U1, "separable operator needs ratio | img_dim and the L matrix");
      break;
    case NHMC_OP_CS:
      NHMC_CHECK_ARG(op->n_kept >= 0 && (op->n_kept == 0 || op->kept_pix) && (int64_t)op->n_kept <= (int64_t)op->img_dim * op->img_dim, "compressive sensing needs the pixel permutation (kept_pix) and n_kept <= D*D");
      break;
    case NHMC_OP_DENOISE:
      break;
    case NHMC_OP_COLOR:
      NHMC_CHECK_ARG(op->channels == 3, "colorization is defined for 3 channels");
      break;
    default:
      set_error("unknown operator kind %d", op->kind);
      return NHMC_ERR_ARG;
  }
  return NHMC_OK;
}

}  // namespace nhmc

using namespace nhmc;

extern "C" int64_t nhmc_op_meas_size(const nhmc_op* op) {
  if (!op) return -1;
  const int64_t C = op->channels, D = op->img_dim;
  switch (op->kind) {
    case NHMC_OP_INPAINT: return (int64_t)op->n_kept * C;
    case NHMC_OP_SR: return op->ratio > 0 ? C * (D / op->ratio) * (D / op->ratio) : -1;
    case NHMC_OP_HDR: return C * D * D;
    case NHMC_OP_DEBLUR2D: return C * D * D;
    case NHMC_OP_PHASE: return C * (D + 2 * op->pad) * (D + 2 * op->pad);
    case NHMC_OP_DENOISE: return C * D * D;
    case NHMC_OP_COLOR: return D * D;
    case NHMC_OP_CS: return (int64_t)op->n_kept * C;
    case NHMC_OP_SEPMAT: return op->ratio > 0 ? C * (D / op->ratio) * (D / op->ratio) : -1;
    default: return -1;
  }
}

extern "C" size_t nhmc_op_tables_bytes(const nhmc_op* op) {
  if (!op) return 0;
  if (op->kind == NHMC_OP_PHASE) return phase_tables_bytes(op);
  if (op->kind != NHMC_OP_DEBLUR2D && op->kind != NHMC_OP_SEPMAT) return 0;
  return spectral_tables_bytes(op);
}

extern "C" int nhmc_op_prepare_tables(const nhmc_op* op, void* tables, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(op && tables && aligned16(tables), "operator and a 16-byte aligned table buffer required");
  NHMC_CHECK_ARG(nhmc_op_tables_bytes(op) > 0, "this operator has no derived tables");
  if (op->kind == NHMC_OP_PHASE) return phase_prepare_tables(op, tables, (cudaStream_t)stream);
  return spectral_prepare_tables(op, reinterpret_cast<float*>(tables), (cudaStream_t)stream);
}

extern "C" size_t nhmc_op_workspace_bytes(const nhmc_op* op, int n_chains) {
  if (!op || n_chains < 1) return 0;
  size_t base = red_ws_bytes(n_chains);
  base = (base + 255) & ~(size_t)255;
  if (op->kind == NHMC_OP_DEBLUR2D) return base + deblur_workspace_bytes(op, n_chains);
  if (op->kind == NHMC_OP_SEPMAT) return base + sepmat_workspace_bytes(op, n_chains);
  if (op->kind == NHMC_OP_CS) return base + cs_workspace_bytes(op, n_chains);
  if (op->kind == NHMC_OP_PHASE) return base + phase_workspace_bytes(op, n_chains);
  return base;
}

extern "C" int nhmc_op_forward(const nhmc_op* op, const float* x, float* y, int n_chains, void* workspace,
                               nhmc_stream_t stream) {
  if (int e = check_op(op)) return e;
  NHMC_CHECK_ARG(x && n_chains > 0, "bad arguments");
  if (nhmc_op_meas_size(op) == 0) return NHMC_OK;  // nothing is measured (all pixels dropped)
  NHMC_CHECK_ARG(y, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = op->channels, D = op->img_dim;
  const int64_t P = (int64_t)D * D;
  switch (op->kind) {
    case NHMC_OP_INPAINT: {
      const int64_t M = (int64_t)op->n_kept * C;
      if (M == 0) return NHMC_OK;
      inpaint_fwd_kernel<<<dim3(cta_count(M, kThreads, n_chains), n_chains), kThreads, 0, st>>>(x, y, op->kept_pix, op->n_kept, C, P);
      break;
    }
    case NHMC_OP_SR: {
      const int64_t M = nhmc_op_meas_size(op);
      sr_fwd_kernel<<<dim3(cta_count(M, kThreads, n_chains), n_chains), kThreads, 0, st>>>(x, y, C, D, op->ratio);
      break;
    }
    case NHMC_OP_HDR: {
      const int64_t total = (int64_t)n_chains * C * P;
      hdr_fwd_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x, y, total);
      break;
    }
    case NHMC_OP_DENOISE: {
      const int64_t total = (int64_t)n_chains * C * P;
      copy_kernel<<<cta_count(total, kThreads * 4, 1), kThreads, 0, st>>>(x, y, total);
      break;
    }
    case NHMC_OP_COLOR:
      color_fwd_kernel<<<dim3(cta_count(P, kThreads, n_chains), n_chains), kThreads, 0, st>>>(x, y, P, coef3(op, 0));
      break;
    case NHMC_OP_DEBLUR2D: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return deblur_forward(op, x, y, n_chains, ws, st);
    }
    case NHMC_OP_PHASE: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return phase_forward(op, x, y, n_chains, ws, st);
    }
    case NHMC_OP_SEPMAT: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return sepmat_forward(op, x, y, n_chains, ws, st);
    }
    case NHMC_OP_CS: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return cs_forward(op, x, y, n_chains, ws, st);
    }
  }
  return NHMC_LAUNCH_CHECK();
}

static int adjoint_or_pinv(const nhmc_op* op, const float* r, float* out, int n_chains, void* workspace,
                           cudaStream_t st, bool pinv) {
  if (int e = check_op(op)) return e;
  NHMC_CHECK_ARG(r && out && n_chains > 0, "bad arguments");
  const int C = op->channels, D = op->img_dim;
  const int64_t P = (int64_t)D * D, N = (int64_t)C * P;
  switch (op->kind) {
    case NHMC_OP_INPAINT:
      inpaint_adj_kernel<<<dim3(cta_count(N, kThreads * 4, n_chains), n_chains), kThreads, 0, st>>>(r, out, op->pix_to_kept, op->n_kept, C, P);
      break;
    case NHMC_OP_SR:
      sr_up_kernel<<<dim3(cta_count(N, kThreads * 4, n_chains), n_chains), kThreads, 0, st>>>(
          r, out, C, D, op->ratio, pinv ? 1.0f : 1.0f / (float)(op->ratio * op->ratio));
      break;
    case NHMC_OP_HDR:
      if (!pinv) { set_error("hdr is non-linear: no adjoint; use nhmc_op_loss_grad"); return NHMC_ERR_UNSUPPORTED; }
      copy_kernel<<<cta_count((int64_t)n_chains * N, kThreads * 4, 1), kThreads, 0, st>>>(r, out, (int64_t)n_chains * N);
      break;
    case NHMC_OP_DENOISE:
      copy_kernel<<<cta_count((int64_t)n_chains * N, kThreads * 4, 1), kThreads, 0, st>>>(r, out, (int64_t)n_chains * N);
      break;
    case NHMC_OP_COLOR:
      color_adj_kernel<<<dim3(cta_count(P, kThreads, n_chains), n_chains), kThreads, 0, st>>>(r, out, P, coef3(op, pinv ? 6 : 3));
      break;
    case NHMC_OP_DEBLUR2D: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      NHMC_CHECK_ARG(!pinv || op->gain_pinv, "gain_pinv table missing");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return deblur_adjoint(op, r, out, n_chains, ws, st, pinv);
    }
    case NHMC_OP_CS: {   // all singular values are 1: H^+ = H^T
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return cs_adjoint(op, r, out, n_chains, ws, st);
    }
    case NHMC_OP_SEPMAT: {
      NHMC_CHECK_ARG(workspace, "workspace required");
      NHMC_CHECK_ARG(!pinv || op->V1, "pseudo-inverse matrix P missing");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return sepmat_adjoint(op, r, out, n_chains, ws, st, pinv);
    }
    case NHMC_OP_PHASE: {
      if (!pinv) { set_error("phase retrieval is non-linear: no adjoint; use nhmc_op_loss_grad"); return NHMC_ERR_UNSUPPORTED; }
      NHMC_CHECK_ARG(workspace, "workspace required");
      char* ws = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
      return phase_pinv(op, r, out, n_chains, ws, st);
    }
  }
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_op_adjoint(const nhmc_op* op, const float* r, float* out, int n_chains, void* workspace,
                               nhmc_stream_t stream) {
  return adjoint_or_pinv(op, r, out, n_chains, workspace, (cudaStream_t)stream, false);
}
extern "C" int nhmc_op_pinv(const nhmc_op* op, const float* y, float* out, int n_chains, void* workspace,
                            nhmc_stream_t stream) {
  return adjoint_or_pinv(op, y, out, n_chains, workspace, (cudaStream_t)stream, true);
}

extern "C" int nhmc_op_loss_grad(const nhmc_op* op, const float* x_hat, const float* y, float* g, float* xc_out,
                                 double* loss, int n_chains, void* workspace, nhmc_stream_t stream) {
  if (int e = check_op(op)) return e;
  NHMC_CHECK_ARG(x_hat && g && loss && workspace && n_chains > 0, "bad arguments");
  NHMC_CHECK_ARG(y || nhmc_op_meas_size(op) == 0, "null measurement");
  NHMC_CHECK_ARG(aligned16(x_hat) && aligned16(g) && (!xc_out || aligned16(xc_out)), "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = op->channels, D = op->img_dim;
  const int64_t P = (int64_t)D * D, N = (int64_t)C * P;
  const RedWs ws = red_ws(workspace, n_chains);
  char* extra = (char*)workspace + ((red_ws_bytes(n_chains) + 255) & ~(size_t)255);
  switch (op->kind) {
    case NHMC_OP_INPAINT: {
      NHMC_CHECK_ARG(P % 4 == 0 && aligned16(op->pix_to_kept), "img_dim^2 must be a multiple of 4");
      dim3 grid(cta_count(N / 4, kThreads * 2, n_chains), n_chains);   // gather-bound: many small CTAs hide the dependent y loads best
      inpaint_loss_grad_kernel<<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, op->pix_to_kept, op->n_kept, C, P, loss, ws, op->no_clamp);
      break;
    }
    case NHMC_OP_SR: {
      const int r = op->ratio;
      if ((r == 4 || r == 8 || r == 16) && D % r == 0 && aligned16(x_hat) && aligned16(g)) {
        const int rs = r == 16 ? 4 : (r == 8 ? 2 : 1);                       // a lane keeps 4 rows: ~60 registers, full occupancy
        const int64_t items = (int64_t)C * (D / r) * (D / r) * (r / 4) * rs;
        dim3 grid(reduce_cta_count(items, kThreads, n_chains), n_chains);
        if (r == 4) sr_loss_grad_reg_kernel<4, 1><<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, C, D, loss, ws, op->no_clamp);
        else if (r == 8) sr_loss_grad_reg_kernel<8, 2><<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, C, D, loss, ws, op->no_clamp);
        else sr_loss_grad_reg_kernel<16, 4><<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, C, D, loss, ws, op->no_clamp);
        break;
      }
      const int TR = sr_tile_rows(D, r);
      NHMC_CHECK_ARG(D % 4 == 0 && D % TR == 0, "img_dim must be a multiple of 4 and of the strip height");
      const int strips = C * (D / TR);
      NHMC_CHECK_ARG(strips <= kMaxBlocksPerChain, "image too large for the SR strip reduction");
      const size_t smem = ((size_t)TR * D + (size_t)TR * (D / r) + (size_t)(TR / r) * (D / r)) * sizeof(float);
      NHMC_CHECK_ARG(smem <= 200 * 1024, "SR strip does not fit in shared memory");
      if (smem > 48 * 1024) {
        if (int e = cuda_status(cudaFuncSetAttribute(sr_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sr smem")) return e;
      }
      sr_loss_grad_kernel<<<dim3(strips, n_chains), kThreads, smem, st>>>(x_hat, y, g, xc_out, C, D, r, TR, loss, ws, op->no_clamp);
      break;
    }
    case NHMC_OP_HDR: {
      NHMC_CHECK_ARG(N % 4 == 0 && aligned16(y), "N must be a multiple of 4");
      dim3 grid(reduce_cta_count(N / 4, kThreads * 4, n_chains), n_chains);
      hdr_loss_grad_kernel<<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, N, loss, ws, op->no_clamp);
      break;
    }
    case NHMC_OP_DENOISE: {
      NHMC_CHECK_ARG(N % 4 == 0 && aligned16(y), "N must be a multiple of 4");
      dim3 grid(reduce_cta_count(N / 4, kThreads * 4, n_chains), n_chains);
      denoise_loss_grad_kernel<<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, N, loss, ws, op->no_clamp);
      break;
    }
    case NHMC_OP_COLOR: {
      NHMC_CHECK_ARG(P % 4 == 0 && aligned16(y), "img_dim^2 must be a multiple of 4");
      dim3 grid(reduce_cta_count(P / 4, kThreads, n_chains), n_chains);
      color_loss_grad_kernel<<<grid, kThreads, 0, st>>>(x_hat, y, g, xc_out, P, coef3(op, 0), coef3(op, 3), loss, ws, op->no_clamp);
      break;
    }
    case NHMC_OP_DEBLUR2D:
      return deblur_loss_grad(op, x_hat, y, g, xc_out, loss, n_chains, workspace, st);
    case NHMC_OP_SEPMAT:
      return sepmat_loss_grad(op, x_hat, y, g, xc_out, loss, n_chains, workspace, st);
    case NHMC_OP_CS:
      return cs_loss_grad(op, x_hat, y, g, xc_out, loss, n_chains, workspace, st);
    case NHMC_OP_PHASE:
      (void)extra;
      return phase_loss_grad(op, x_hat, y, g, xc_out, loss, n_chains, workspace, st);
  }
  return NHMC_LAUNCH_CHECK();
}

extern "C" int nhmc_op_prepare_measurement(const nhmc_op* op, const float* y, float* y_prepared, int n_chains, void* workspace,
                                           nhmc_stream_t stream) {
  NHMC_CHECK_ARG(op && y && y_prepared && workspace && n_chains > 0, "bad arguments");
  if (op->kind != NHMC_OP_DEBLUR2D || nhmc_op_tables_bytes(op) == 0) {
    set_error("this operator evaluates against the measurement as it is (no prepared form)");
    return NHMC_ERR_UNSUPPORTED;
  }
  NHMC_CHECK_ARG(aligned16(y) && aligned16(y_prepared), "measurements must be 16-byte aligned");
  return deblur_spectral_measurement(op, y, y_prepared, n_chains, (char*)workspace + red_ws_aligned(n_chains), (cudaStream_t)stream);
}

extern "C" int nhmc_op_loss_grad_prepared(const nhmc_op* op, const float* x_hat, const float* y_prepared, float* g, float* xc_out,
                                          double* loss, int n_chains, void* workspace, nhmc_stream_t stream) {
  NHMC_CHECK_ARG(op && x_hat && y_prepared && g && loss && workspace && n_chains > 0, "bad arguments");
  NHMC_CHECK_ARG(op->kind == NHMC_OP_DEBLUR2D && nhmc_op_tables_bytes(op) > 0, "only the spectral deblurring operators have a prepared measurement");
  NHMC_CHECK_ARG(aligned16(x_hat) && aligned16(y_prepared) && aligned16(g) && (!xc_out || aligned16(xc_out)), "buffers must be 16-byte aligned");
  return deblur_loss_grad_spec(op, x_hat, y_prepared, g, xc_out, loss, n_chains, workspace, (cudaStream_t)stream);
}
