// tcgen05 / TMEM / TMA tensor-core GEMMs (TF32 inputs, fp32 accumulate) for sm_100a.
// Each launcher returns the number of kernels launched, or -1 when the shape is outside what the
// kernel covers (the caller then uses the fp32 CUDA-core path of gemm_simt.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gatx {
// Optional epilogue of the input-gradient GEMM (replaces EB:879-893 and the segment sum of EB:682-691): the tile is
// dL/dHout of the layer below; with Hout of that layer at hand the epilogue writes g_pre = g * LReLU'(Hout) instead of g
// and cdot[row][head] = g . Hout - g_pre . bias (= sum over the row's edges of alpha * galpha), so the separate node-wise
// pass over the gradient (12 N F bytes) disappears.  Needs 256-column tiles (the persistent kernel), head_dim % 32 == 0.
struct GemmFuse {
  const float* Hout = nullptr;  // [M][ld] layer output (LReLU(h) concatenated over heads); nullptr = plain GEMM
  int ld = 0;                   // row pitch of Hout (= N)
  float* cdot = nullptr;        // [M][heads]
  const float* bias = nullptr;  // [N] or nullptr
  float act_slope = 0.01f;
  int head_dim = 0, heads = 0;
};
// C[M][N] (ldc) (+)= A[M][K] (lda) * B[N][K]^T (ldb); all row-major with K contiguous.
int launch_gemm_tc_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                      int K, bool accumulate, cudaStream_t st);
// C[m][n] (+)= A0 B0^T + A1 B1^T (second pair optional: K1 = 0); columns n >= n_split are written to C1
// (at column n - n_split), so P_l and P_r come out of one pass over X.
int launch_gemm_tc_tn2(const float* A0, int64_t lda0, const float* B0, int64_t ldb0, int K0, const float* A1,
                       int64_t lda1, const float* B1, int64_t ldb1, int K1, float* C0, float* C1, int n_split,
                       int64_t ldc, int M, int N, bool accumulate, cudaStream_t st, const GemmFuse* fuse = nullptr);
// C[M][N] (ldc) += A[K][M]^T (lda) * B[K][N] (ldb): contraction over the (huge) node dimension K.
int launch_gemm_tc_atb(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                       int64_t K, float* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace gatx
