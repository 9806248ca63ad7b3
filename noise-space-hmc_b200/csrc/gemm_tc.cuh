// Batched fp32-grade GEMM on the 5th-generation tensor cores (tcgen05, "3xTF32"):
//   C[mat] (M x N) = A[mat] (M x K) * B[mat]^T (N x K)      both operands K-major ("TN"), fp32 accumulators in TMEM
// Every operand is a PAIR of fp32 planes (hi, lo) with hi = x truncated to tf32 and lo = x - hi (exact), so that
//   A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (dropped term ~2^-22 relative: fp32-grade, what the 1e-3
// image tolerance of the deblurring operators needs; a single TF32 pass is not enough).
// Used by the spectral operators (deblur.cu): Hfuncs.py:448-523 (Deblurring2D), :236-316 (Deblurring), :527-607 (SRConv).
#pragma once
#include "common.cuh"

namespace nhmc {

enum { TC_EPI_NONE = 0, TC_EPI_GAIN = 1, TC_EPI_RESID = 2, TC_EPI_GRAD = 3, TC_EPI_SPEC = 4 };

// A plane pair in global memory: plane p (0 = hi, 1 = lo) of matrix `mat` starts at base + p*plane_stride + mat*mat_stride
// (elements); rows are dense (row stride = cols).  mat_stride 0 = one table shared by all matrices.
struct TcOperand {
  const float* base;
  int rows, cols;            // rows x K, K contiguous
  int64_t mat_stride, plane_stride;
};

struct TcGemm {
  TcOperand A, B;            // A: M x K, B: N x K
  int mats, channels;
  int epi;
  int out_transposed;        // store C^T (N x M) instead of C (M x N)
  float* out_hi;             // plane pair of the result (the next product's operand) ...
  float* out_lo;             // ... or nullptr: out_hi receives the plain fp32 result
  int64_t out_mat_stride;
  const float* gain;         // [channels][M][N]                (TC_EPI_GAIN)   indexed in C's logical (m, n)
  const float* y;            // [mats][M][N] measurement        (TC_EPI_RESID): v = y - v, loss += v^2
                             //                                 (TC_EPI_SPEC):  r = y - gain v, loss += r^2, v = gain r   (spectral-domain residual)
  const float* xhat;         // [mats][M][N] unclamped image    (TC_EPI_GRAD):  v = -2 v 1[-1 <= xhat <= 1]
  double* loss;              // [chains]
  int no_clamp;
  RedWs ws;
};

bool tc_gemm_supported(int M, int N, int K);
int tc_gemm_launch(const TcGemm& g, cudaStream_t st);

// hi/lo split of `n` fp32 values (optionally clamped to [-1, 1] first): hi = x & ~0x1fff, lo = x - hi
int tc_split_launch(const float* x, float* hi, float* lo, int64_t n, int clamp, cudaStream_t st);
// hi/lo split of a row-major [rows x cols] table, optionally transposed (out is [cols x rows])
int tc_split_table_launch(const float* t, float* hi, float* lo, int rows, int cols, int transpose, cudaStream_t st);

}  // namespace nhmc
