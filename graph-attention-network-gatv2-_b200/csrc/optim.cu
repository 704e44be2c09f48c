// Gradient clipping + Adam / SGD + gradient reset as two launches over one flat parameter buffer.
// Replaces reduce_sum_squares / scale_grads / clip_grad_norm (EB:146-177, 250-278: per group a
// cudaMalloc, two blocking memcpys and atomics on the partial sums), adam_update_kernel (EB:896-916),
// sgd_update_kernel (EB:919-923) and the cudaMemsets of EB:1631-1633.  Clip groups are the
// reference's: {W of all layers}, {a of all layers}, {W_o} (EB:1561-1566), threshold 5.0.
#include "common.cuh"

namespace gatx {

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ grads, OptimGroups grp, float* __restrict__ partials) {
  __shared__ float red[8];
  for (int g = 0; g < 3; ++g) {
    float s = 0.f;
    for (int64_t i = grp.begin[g] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < grp.end[g];
         i += (int64_t)gridDim.x * blockDim.x) {
      const float v = grads[i];
      s = fmaf(v, v, s);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      partials[g * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
optimizer_kernel(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m,
                 float* __restrict__ v, int64_t n, OptimGroups grp, int clip, int optimizer, float lr, float b1,
                 float b2, int t, const float* __restrict__ partials, int n_partials) {
  __shared__ float scale_s[3];
  if (threadIdx.x < 3) {
    float scale = 1.0f;
    if (clip) {
      double ss = 0.0;
      for (int b = 0; b < n_partials; ++b) ss += (double)partials[threadIdx.x * n_partials + b];  // fixed order
      const float norm = sqrtf((float)ss);                        // EB:268
      if (norm > 5.0f) scale = 5.0f / (norm + 1e-9f);            // EB:270-272
      if (!(scale < 1.0f)) scale = 1.0f;                          // EB:275
    }
    scale_s[threadIdx.x] = scale;
  }
  __syncthreads();
  const float c1 = 1.0f - powf(b1, (float)t), c2 = 1.0f - powf(b2, (float)t);  // EB:908, EB:911
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = i < grp.end[0] ? 0 : (i < grp.end[1] ? 1 : 2);
    const float sc = scale_s[g];
    float gr = grads[i];
    if (sc < 1.0f) gr *= sc;
    if (optimizer == 1) {
      const float mi = b1 * m[i] + (1.0f - b1) * gr;
      const float vi = b2 * v[i] + (1.0f - b2) * (gr * gr);
      m[i] = mi;
      v[i] = vi;
      params[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + 1e-8f);  // EB:914
    } else {
      params[i] -= lr * gr;  // EB:922
    }
    grads[i] = 0.f;  // EB:1631-1633
  }
}

int launch_optimizer(float* params, float* grads, float* m, float* v, int64_t n, OptimGroups grp, bool clip,
                     int optimizer, float lr, float b1, float b2, int t, float* norm_partials, cudaStream_t st) {
  int launches = 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kOptimBlocks) blocks = kOptimBlocks;
  if (blocks < 1) blocks = 1;
  if (clip) {
    grad_sumsq_kernel<<<blocks, 256, 0, st>>>(grads, grp, norm_partials);
    ++launches;
  }
  optimizer_kernel<<<blocks, 256, 0, st>>>(params, grads, m, v, n, grp, clip ? 1 : 0, optimizer, lr, b1, b2, t,
                                           norm_partials, blocks);
  return launches + 1;
}

}  // namespace gatx
