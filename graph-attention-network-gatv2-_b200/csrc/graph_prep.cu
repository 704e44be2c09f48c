// Graph preparation (one-off, integer, bit-exact against the oracle) and parameter init.
//   csr -> coo + in-degree      replaces csr_to_coo_kernel (EB:67-84) and the degree loop (EB:89-99)
//   stable source-major transpose (csc_ptr / csc_dst / csc_eid): new, feeds the deterministic
//     backward scatter that replaces the reference's global float atomics (EB:786, EB:868-869)
//   Philox Xavier-uniform init   replaces setup_states_kernel / xavier_init_kernel_curand (EB:181-248)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gatx {

__global__ void csr_to_coo_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                  int* __restrict__ src, int* __restrict__ dst, int* __restrict__ deg,
                                  int n_rows) {
  // one warp per row: coalesced over the row's edges (the reference walks a row per thread)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int beg = row_ptr[warp], end = row_ptr[warp + 1];
  if (lane == 0 && deg) deg[warp] = end - beg;
  for (int e = beg + lane; e < end; e += 32) {
    src[e] = col_idx[e];
    dst[e] = warp;
  }
}

int launch_csr_to_coo(const int* row_ptr, const int* col_idx, int* src, int* dst, int* deg, int n_rows,
                      cudaStream_t st) {
  if (n_rows <= 0) return 0;
  const int threads = 256, rows_per_block = threads / 32;
  csr_to_coo_kernel<<<(n_rows + rows_per_block - 1) / rows_per_block, threads, 0, st>>>(row_ptr, col_idx, src,
                                                                                      dst, deg, n_rows);
  return 1;
}

__global__ void iota_kernel(int* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (int)i;
}
__global__ void count_sources_kernel(const int* __restrict__ col_idx, int64_t E, int* __restrict__ counts) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&counts[col_idx[i]], 1);  // integer: order-independent, bit-exact
}
__global__ void gather_int_kernel(const int* __restrict__ table, const int* __restrict__ idx, int64_t n,
                                  int* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = table[idx[i]];
}

int build_csc(const int* col_idx, const int* coo_dst, int64_t E, int n_src, int* csc_ptr, int* csc_dst,
              int* csc_eid, cudaStream_t st) {
  int launches = 0;
  if (cudaMemsetAsync(csc_ptr, 0, sizeof(int) * (size_t)(n_src + 1), st) != cudaSuccess) return -1;
  if (E == 0) return 0;
  int *keys_out = nullptr, *iota = nullptr, *counts = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, scan_bytes = 0;
  int end_bit = 1;
  while ((1ll << end_bit) < (long long)n_src) ++end_bit;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, col_idx, keys_out, iota, csc_eid, (int)E, 0, end_bit, st);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, counts, csc_ptr + 1, n_src, st);
  if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
  bool ok = cudaMalloc(&keys_out, sizeof(int) * (size_t)E) == cudaSuccess &&
            cudaMalloc(&iota, sizeof(int) * (size_t)E) == cudaSuccess &&
            cudaMalloc(&counts, sizeof(int) * (size_t)n_src) == cudaSuccess &&
            cudaMalloc(&tmp, tmp_bytes) == cudaSuccess;
  if (ok) {
    const int threads = 256;
    const int blocks = (int)((E + threads - 1) / threads < 148 * 16 ? (E + threads - 1) / threads : 148 * 16);
    iota_kernel<<<blocks, threads, 0, st>>>(iota, E);
    cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)n_src, st);
    count_sources_kernel<<<blocks, threads, 0, st>>>(col_idx, E, counts);
    // LSD radix sort is stable: equal sources keep ascending CSR position
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, col_idx, keys_out, iota, csc_eid, (int)E, 0, end_bit, st);
    cub::DeviceScan::InclusiveSum(tmp, scan_bytes, counts, csc_ptr + 1, n_src, st);
    gather_int_kernel<<<blocks, threads, 0, st>>>(coo_dst, csc_eid, E, csc_dst);
    launches = 6;
    ok = cudaStreamSynchronize(st) == cudaSuccess;
  }
  cudaFree(keys_out);
  cudaFree(iota);
  cudaFree(counts);
  cudaFree(tmp);
  return ok ? launches : -1;
}

// ---- Philox4x32-10 counter-based generator -------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__global__ void philox_uniform_kernel(float* __restrict__ out, int64_t n, float limit, uint64_t seed,
                                      uint64_t stream_id) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t i = q * 4 + t;
      if (i < n) {
        // (0,1] like curand_uniform, then the reference's affine map rnd*2*limit - limit (EB:217-219)
        const float u = ((float)(c[t] >> 8) + 1.0f) * (1.0f / 16777216.0f);
        out[i] = u * 2.0f * limit - limit;
      }
    }
  }
}
int launch_philox_uniform(float* out, int64_t n, float limit, uint64_t seed, uint64_t stream_id,
                          cudaStream_t st) {
  if (n <= 0) return 0;
  const int threads = 256;
  int64_t blocks = (n / 4 + threads) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  philox_uniform_kernel<<<(int)blocks, threads, 0, st>>>(out, n, limit, seed, stream_id);
  return 1;
}

// Inverted dropout (extension).  Thread per (row, 4-column group): element (n, 4q + t) is kept iff word t
// of Philox4x32-10(counter = {q, global row, layer, step}, key = seed) >= thresh = floor(p * 2^32); kept elements are
// scaled by 1 / (1 - p).  The backward pass re-generates the same words to mask the gradient (Y == X: in place).
__global__ void dropout_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int n_rows,
                               int cols, int row0, uint32_t thresh, float scale, uint64_t seed, uint32_t layer,
                               uint32_t step) {
  const int quads = (cols + 3) / 4;
  const int64_t total = (int64_t)n_rows * quads;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / quads), q = (int)(i - (int64_t)n * quads);
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(row0 + n), layer, step};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const float* x = X + (int64_t)n * ldx + 4 * q;
    float* y = Y + (int64_t)n * ldy + 4 * q;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (4 * q + t < cols) y[t] = c[t] >= thresh ? x[t] * scale : 0.f;
  }
}
int launch_dropout(const float* X, int64_t ldx, float* Y, int64_t ldy, int n_rows, int cols, int row0, float p,
                   uint64_t seed, int layer, int64_t step, cudaStream_t st) {
  if (n_rows <= 0 || cols <= 0) return 0;
  const int64_t total = (int64_t)n_rows * ((cols + 3) / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  dropout_kernel<<<(int)blocks, 256, 0, st>>>(X, ldx, Y, ldy, n_rows, cols, row0, thresh, 1.0f / (1.0f - p), seed,
                                             (uint32_t)layer, (uint32_t)step);
  return 1;
}

__global__ void attn_dropout_scale_kernel(float* __restrict__ out, int64_t E, int H, int64_t edge0, uint32_t thresh,
                                          float scale, uint64_t seed, uint32_t layer, uint32_t step) {
  const int quads = (H + 3) / 4;
  const int64_t total = E * quads;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / quads;
    const int q = (int)(i - e * quads);
    uint32_t c[4] = {(uint32_t)(edge0 + e), (uint32_t)q, 0x80000000u | layer, step};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (4 * q + t < H) out[e * H + 4 * q + t] = c[t] >= thresh ? scale : 0.f;
  }
}
int launch_attn_dropout_scale(float* out, int64_t E, int H, int64_t edge0, float p, uint64_t seed, int layer, int64_t step,
                              cudaStream_t st) {
  if (E <= 0 || H <= 0) return 0;
  const int64_t total = E * ((H + 3) / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  attn_dropout_scale_kernel<<<(int)blocks, 256, 0, st>>>(out, E, H, edge0, thresh, 1.0f / (1.0f - p), seed, (uint32_t)layer,
                                                        (uint32_t)step);
  return 1;
}

__global__ void mark_hot_kernel(const int* __restrict__ idx, const int* __restrict__ ptr, int64_t E, int thr_wide,
                                int thr_narrow, int* __restrict__ out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int v = idx[e];
    const int deg = ptr[v + 1] - ptr[v];
    uint32_t o = (uint32_t)v;
    if (deg >= thr_wide) o |= 0x80000000u;
    if (deg >= thr_narrow) o |= 0x40000000u;
    out[e] = (int)o;
  }
}
int launch_mark_hot(const int* idx, const int* ptr, int64_t E, int thr_wide, int thr_narrow, int* out, cudaStream_t st) {
  if (E <= 0) return 0;
  int64_t blocks = (E + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  mark_hot_kernel<<<(int)blocks, 256, 0, st>>>(idx, ptr, E, thr_wide, thr_narrow, out);
  return 1;
}

}  // namespace gatx
