// fp32 CUDA-core GEMM with arbitrary operand strides and deterministic split-K.
// This is the exact-arithmetic path (GATX_GEMM_FP32_SIMT) and the fallback for shapes the
// tcgen05 kernels (gemm_tc.cu) do not cover.  It evaluates, once per node, the contractions the
// reference recomputes per edge: W_l x / W_r x (EB:303-316, EB:415-420), gW = gP^T X (EB:771-782)
// and gX = gP W (EB:859-869).
#include "common.cuh"

namespace gatx {

constexpr int TM = 128, TN = 64, TK = 16, GT = 256;

// C[m][n] = sum_{k in split} A(m,k) * B(n,k);  A(m,k) = A[m*sAm + k*sAk], B(n,k) = B[n*sBn + k*sBk]
__global__ void __launch_bounds__(GT)
gemm_simt_kernel(const float* __restrict__ A, int64_t sAm, int64_t sAk, const float* __restrict__ B,
                 int64_t sBn, int64_t sBk, float* __restrict__ C, int64_t ldc, int M, int N, int64_t K,
                 int64_t k_per_split, int accumulate, float* __restrict__ ws) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = kbeg + k_per_split < K ? kbeg + k_per_split : K;
  const int ty = tid / 16, tx = tid % 16;  // 16 x 16 threads, 8 x 4 outputs each
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kmajor = (sAk == 1), b_kmajor = (sBk == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int r = 0; r < TM * TK / GT; ++r) {
      const int id = tid + r * GT;
      const int m = a_kmajor ? id / TK : id % TM, k = a_kmajor ? id % TK : id / TM;
      const int gm = m0 + m;
      const int64_t gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? __ldg(A + (int64_t)gm * sAm + gk * sAk) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < TN * TK / GT; ++r) {
      const int id = tid + r * GT;
      const int n = b_kmajor ? id / TK : id % TN, k = b_kmajor ? id % TK : id / TN;
      const int gn = n0 + n;
      const int64_t gk = k0 + k;
      Bs[k][n] = (gn < N && gk < kend) ? __ldg(B + (int64_t)gn * sBn + gk * sBk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
  float* out = split ? ws + (int64_t)blockIdx.z * M * N : C;
  const int64_t ldo = split ? N : ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* p = out + (int64_t)gm * ldo + gn;
      *p = (!split && accumulate) ? *p + acc[i][j] : acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                     float* __restrict__ C, int64_t ldc, int accumulate) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * total + i];  // fixed order: deterministic
    float* p = C + (i / N) * ldc + (i % N);
    *p = accumulate ? *p + s : s;
  }
}

int launch_gemm_simt(const float* A, int64_t sAm, int64_t sAk, const float* B, int64_t sBn, int64_t sBk,
                     float* C, int64_t ldc, int M, int N, int64_t K, bool accumulate, float* splitk_ws,
                     size_t splitk_ws_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int gm = (M + TM - 1) / TM, gn = (N + TN - 1) / TN;
  int splits = 1;
  if (splitk_ws && (int64_t)gm * gn < 2 * kNumSMs && K > 4096) {
    splits = (int)((4 * kNumSMs + (int64_t)gm * gn - 1) / ((int64_t)gm * gn));
    const int64_t max_by_k = (K + 1023) / 1024;
    if (splits > max_by_k) splits = (int)max_by_k;
    const size_t per = sizeof(float) * (size_t)M * N;
    if ((size_t)splits * per > splitk_ws_bytes) splits = (int)(splitk_ws_bytes / per);
    if (splits < 1) splits = 1;
  }
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + TK - 1) / TK * TK;
  splits = (int)((K + kps - 1) / kps);
  if (splits < 1) splits = 1;
  dim3 grid(gm, gn, splits);
  gemm_simt_kernel<<<grid, GT, 0, st>>>(A, sAm, sAk, B, sBn, sBk, C, ldc, M, N, K, kps, accumulate ? 1 : 0,
                                        splitk_ws);
  if (splits == 1) return 1;
  const int64_t total = (int64_t)M * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(splitk_ws, splits, M, N, C, ldc, accumulate ? 1 : 0);
  return 2;
}

// Wcat[r][k]      = W[r][k]       r in [0,F)   (W_l, applied to the source)
// Wcat[F + r][k]  = W[r][I + k]                (W_r, applied to the destination); k >= I zero
// WcatT[k][c]     = Wcat[c][k]    [I][2F]
__global__ void pack_weights_kernel(const float* __restrict__ W, int F, int I, float* __restrict__ Wcat,
                                    int ldk, float* __restrict__ WcatT) {
  const int64_t total = (int64_t)2 * F * ldk;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / ldk), k = (int)(i % ldk);
    float v = 0.f;
    if (k < I) v = c < F ? W[(int64_t)c * 2 * I + k] : W[(int64_t)(c - F) * 2 * I + I + k];
    Wcat[i] = v;
    if (k < I && WcatT) WcatT[(int64_t)k * 2 * F + c] = v;
  }
}
int launch_pack_weights(const float* W, int F, int I, float* Wcat, int ldk, float* WcatT, cudaStream_t st) {
  const int64_t total = (int64_t)2 * F * ldk;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  pack_weights_kernel<<<blocks, 256, 0, st>>>(W, F, I, Wcat, ldk, WcatT);
  return 1;
}


__global__ void split_tf32_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, int cols, int cols_pad,
                                  float* __restrict__ hi, int64_t ld_hi, float* __restrict__ lo, int64_t ld_lo) {
  const int64_t total = rows * cols_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols_pad;
    const int c = (int)(i - r * cols_pad);
    float h = 0.f, l = 0.f;
    if (c < cols) {
      const float x = X[r * ldx + c];
      h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      l = x - h;
    }
    hi[r * ld_hi + c] = h;
    lo[r * ld_lo + c] = l;
  }
}
int launch_split_tf32(const float* X, int64_t ldx, int64_t rows, int cols, int cols_pad, float* out_hi, int64_t ld_hi,
                      float* out_lo, int64_t ld_lo, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t total = rows * cols_pad;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  split_tf32_kernel<<<(int)blocks, 256, 0, st>>>(X, ldx, rows, cols, cols_pad, out_hi, ld_hi, out_lo, ld_lo);
  return 1;
}

}  // namespace gatx
