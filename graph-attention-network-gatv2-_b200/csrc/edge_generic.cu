// Generic edge kernels: any number of heads (<= 32) and any per-head dimension with F = H*D <= 1024.
// They keep the drop-in surface complete (the reference accepts arbitrary --heads / --outdims, EB:954-987) for
// shapes the vectorised kernels (edge_kernels.cu, edge_stream.cu) do not cover: D not a multiple of 4, D > 128,
// odd chunk counts.  Same math, same buffers, same determinism (no atomics); scalar coalesced loads, one warp
// per destination (forward, backward pass 1) or source (pass 2) row, lane l owns elements l, l+32, ...
// The per-edge record is {alpha[H], ge[H]}; pass 2 re-derives sign(s) from P_l[src] + P_r[dst].
#include "common.cuh"

namespace gatx {
namespace {

constexpr int kGW = 8;  // warps per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// per-head sums of x[t] (element k = lane + 32 t belongs to head k / D) into out[0..H) (shared, per warp)
template <int T>
__device__ __forceinline__ void head_sums(const float (&x)[T], const int (&hd)[T], int H, float* out, int lane) {
  for (int h = 0; h < H; ++h) {
    float p = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) p += (hd[t] == h) ? x[t] : 0.f;
    p = warp_sum(p);
    if (lane == 0) out[h] = p;
  }
  __syncwarp();
}

template <int T>
__global__ void __launch_bounds__(kGW * 32)
edge_fwd_generic_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                        const float* __restrict__ Pl, const float* __restrict__ Pr, const float* __restrict__ a, int H,
                        int D, Slopes sl, const float* __restrict__ bias, float* __restrict__ Hout, float* __restrict__ hpre, float* __restrict__ score,
                        float* __restrict__ mx, float* __restrict__ sinv, const float* __restrict__ ascale) {
  __shared__ float sc_s[kGW][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, F = H * D;
  const int row = blockIdx.x * kGW + warp;
  if (row >= n_rows) return;
  float av[T], pr[T], acc[T], m[T], s[T];
  int hd[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    const bool ok = k < F;
    hd[t] = ok ? k / D : -1;
    av[t] = ok ? __ldg(a + k) : 0.f;
    pr[t] = ok ? __ldg(Pr + (int64_t)row * F + k) : 0.f;
    acc[t] = 0.f;
    m[t] = -1e9f;  // EB:336
    s[t] = 0.f;
  }
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  for (int e = beg; e < end; ++e) {
    const int src = __ldg(col_idx + e);
    float v[T], part[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int k = lane + 32 * t;
      v[t] = k < F ? __ldg(Pl + (int64_t)src * F + k) : 0.f;
      part[t] = av[t] * lrelu(v[t] + pr[t], sl.attn);  // EB:303-320
    }
    head_sums<T>(part, hd, H, sc_s[warp], lane);
    if (lane < H) score[(int64_t)e * H + lane] = sc_s[warp][lane];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (hd[t] >= 0) {
        const float p = sc_s[warp][hd[t]];
        const float mn = fmaxf(m[t], p);
        const float corr = __expf(m[t] - mn), w = __expf(p - mn);
        s[t] = s[t] * corr + w;
        const float wd = ascale ? w * __ldg(ascale + (int64_t)e * H + hd[t]) : w;  // attention dropout: aggregate only
        acc[t] = acc[t] * corr + wd * v[t];
        m[t] = mn;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    if (k < F) {
      const float inv = 1.0f / (s[t] + 1e-8f);  // EB:379
      const float h = acc[t] * inv + (bias ? bias[k] : 0.f);
      if (hpre) hpre[(int64_t)row * F + k] = h;
      Hout[(int64_t)row * F + k] = lrelu(h, sl.act);
      if (k % D == 0) {
        mx[(int64_t)row * H + hd[t]] = m[t];
        sinv[(int64_t)row * H + hd[t]] = inv;
      }
    }
  }
}

// pass 1: destination-major.  rec[e] = {alpha[H], ge[H]} (RW words per edge)
template <int T>
__global__ void __launch_bounds__(kGW * 32)
edge_bwd_dst_generic_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                            const float* __restrict__ Pl, const float* __restrict__ Pr, const float* __restrict__ a,
                            int H, int D, Slopes sl, const float* __restrict__ bias, const float* __restrict__ Hout,
                            float* __restrict__ gH,
                            const float* __restrict__ score, const float* __restrict__ mx,
                            const float* __restrict__ sinv, float* __restrict__ gPr, float* __restrict__ rec, int RW,
                            float* __restrict__ ga_partials, float* __restrict__ galpha_dbg,
                            const float* __restrict__ ascale) {
  __shared__ float c_s[kGW][32], g_s[kGW][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, F = H * D;
  float av[T], ga[T];
  int hd[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    hd[t] = k < F ? k / D : -1;
    av[t] = k < F ? __ldg(a + k) : 0.f;
    ga[t] = 0.f;
  }
  const int total_warps = gridDim.x * kGW;
  for (int row = blockIdx.x * kGW + warp; row < n_rows; row += total_warps) {
    float gh[T], pr[T], gpr[T], cdot[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int k = lane + 32 * t;
      const float g = k < F ? gH[(int64_t)row * F + k] : 0.f;
      const float ho = k < F ? __ldg(Hout + (int64_t)row * F + k) : 0.f;
      pr[t] = k < F ? __ldg(Pr + (int64_t)row * F + k) : 0.f;
      gh[t] = g * lrelu_grad(ho, sl.act);      // EB:879-893 / EB:599
      // sum over the segment of alpha*galpha = g_pre . h = gH . Hout  (- g_pre . bias when the aggregate carries one)
      cdot[t] = g * ho - ((bias && k < F) ? gh[t] * __ldg(bias + k) : 0.f);
      gpr[t] = 0.f;
      if (k < F) gH[(int64_t)row * F + k] = gh[t];
    }
    head_sums<T>(cdot, hd, H, c_s[warp], lane);
    const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
    for (int e = beg; e < end; ++e) {
      const int src = __ldg(col_idx + e);
      float v[T], gp[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int k = lane + 32 * t;
        v[t] = k < F ? __ldg(Pl + (int64_t)src * F + k) : 0.f;
        gp[t] = gh[t] * v[t];  // EB:636-646
      }
      head_sums<T>(gp, hd, H, g_s[warp], lane);
      float ge_l = 0.f;
      if (lane < H) {
        const float dsc = ascale ? __ldg(ascale + (int64_t)e * H + lane) : 1.f;  // attention dropout (1 when off)
        const float galpha = g_s[warp][lane] * dsc;
        const float alpha = __expf(__ldg(score + (int64_t)e * H + lane) - __ldg(mx + (int64_t)row * H + lane)) *
                            __ldg(sinv + (int64_t)row * H + lane);  // EB:378-379
        ge_l = alpha * (galpha - c_s[warp][lane]);                   // EB:689-690 in closed form
        rec[(int64_t)e * RW + lane] = alpha * dsc;                   // pass 2 aggregates g_h with alpha * dsc
        rec[(int64_t)e * RW + H + lane] = ge_l;
        if (galpha_dbg) galpha_dbg[(int64_t)e * H + lane] = galpha;
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float ge = __shfl_sync(0xffffffffu, ge_l, hd[t] >= 0 ? hd[t] : 0);
        if (hd[t] >= 0) {
          const float sx = v[t] + pr[t];
          ga[t] += ge * lrelu(sx, sl.attn);                    // EB:769
          gpr[t] += ge * av[t] * lrelu_grad(sx, sl.attn);      // EB:774-781
        }
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int k = lane + 32 * t;
      if (k < F) gPr[(int64_t)row * F + k] = gpr[t];
    }
    __syncwarp();
  }
  const int gw = blockIdx.x * kGW + warp;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    if (k < F) ga_partials[(int64_t)gw * F + k] = ga[t];
  }
}

// pass 2: source-major over the stable transpose
template <int T>
__global__ void __launch_bounds__(kGW * 32)
edge_bwd_src_generic_kernel(int n_src, const int* __restrict__ csc_ptr, const int* __restrict__ csc_dst,
                            const int* __restrict__ csc_eid, const float* __restrict__ Pl,
                            const float* __restrict__ Pr, const float* __restrict__ a, int H, int D,
                            Slopes sl, const float* __restrict__ gh, const float* __restrict__ rec, int RW,
                            float* __restrict__ gPl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, F = H * D;
  const int row = blockIdx.x * kGW + warp;
  if (row >= n_src) return;
  float av[T], pl[T], acc[T];
  int hd[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    hd[t] = k < F ? k / D : -1;
    av[t] = k < F ? __ldg(a + k) : 0.f;
    pl[t] = k < F ? __ldg(Pl + (int64_t)row * F + k) : 0.f;
    acc[t] = 0.f;
  }
  const int beg = __ldg(csc_ptr + row), end = __ldg(csc_ptr + row + 1);
  for (int q = beg; q < end; ++q) {
    const int d = __ldg(csc_dst + q), e = __ldg(csc_eid + q);
    float al = 0.f, ge = 0.f;
    if (lane < H) {
      al = __ldg(rec + (int64_t)e * RW + lane);
      ge = __ldg(rec + (int64_t)e * RW + H + lane);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int k = lane + 32 * t;
      const int h = hd[t] >= 0 ? hd[t] : 0;
      const float alh = __shfl_sync(0xffffffffu, al, h), geh = __shfl_sync(0xffffffffu, ge, h);
      if (k < F) {
        const float g = __ldg(gh + (int64_t)d * F + k);
        const float sx = pl[t] + __ldg(Pr + (int64_t)d * F + k);
        acc[t] += alh * g + geh * av[t] * lrelu_grad(sx, sl.attn);  // EB:865-866
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int k = lane + 32 * t;
    if (k < F) gPl[(int64_t)row * F + k] = acc[t];
  }
}

__global__ void unpack_rec_generic_kernel(const float* __restrict__ rec, int64_t E, int H, int RW,
                                          float* __restrict__ alpha, float* __restrict__ ge) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E * H; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / H;
    const int h = (int)(i % H);
    if (alpha) alpha[i] = rec[e * RW + h];
    if (ge) ge[i] = rec[e * RW + H + h];
  }
}

int pick_t(int F) {
  for (int t : {1, 2, 4, 8, 16, 32})
    if (F <= 32 * t) return t;
  return -1;
}

#define GENERIC_DISPATCH(tv, ...)                              \
  do {                                                         \
    switch (tv) {                                              \
      case 1: { constexpr int T = 1; __VA_ARGS__; } break;     \
      case 2: { constexpr int T = 2; __VA_ARGS__; } break;     \
      case 4: { constexpr int T = 4; __VA_ARGS__; } break;     \
      case 8: { constexpr int T = 8; __VA_ARGS__; } break;     \
      case 16: { constexpr int T = 16; __VA_ARGS__; } break;   \
      default: { constexpr int T = 32; __VA_ARGS__; } break;   \
    }                                                          \
  } while (0)

}  // namespace

bool edge_generic_supported(int H, int D) { return H >= 1 && H <= 32 && D >= 1 && (int64_t)H * D <= 1024; }
int edge_generic_rec_words(int H) { return (2 * H + 3) / 4 * 4; }
constexpr int kGenericBwdBlocks = kNumSMs;
int edge_generic_partials() { return kGenericBwdBlocks * kGW; }

int launch_edge_forward_generic(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                                float* Hout, float* hpre, float* score, float* mx, float* sinv, cudaStream_t st) {
  const int tv = pick_t(H * D);
  if (!edge_generic_supported(H, D) || tv < 0) return -1;
  if (g.n_rows <= 0) return 0;
  GENERIC_DISPATCH(tv, edge_fwd_generic_kernel<T><<<(g.n_rows + kGW - 1) / kGW, kGW * 32, 0, st>>>(
                           g.n_rows, g.row_ptr, g.col_idx, Pl, Pr, a, H, D, g.slopes, g.bias, Hout, hpre, score, mx, sinv, g.ascale));
  return 1;
}

int launch_edge_backward_generic(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                                 const float* Hout, float* gH, const float* score, const float* mx, const float* sinv,
                                 float* gPr, float* gPl, uint32_t* rec, float* ga_partials, int* n_partials,
                                 float* galpha_dbg, cudaStream_t st) {
  const int tv = pick_t(H * D);
  if (!edge_generic_supported(H, D) || tv < 0) return -1;
  *n_partials = 0;
  if (g.n_rows <= 0) return 0;
  const int RW = edge_generic_rec_words(H);
  int blocks = (g.n_rows + kGW - 1) / kGW;
  if (blocks > kGenericBwdBlocks) blocks = kGenericBwdBlocks;
  GENERIC_DISPATCH(tv, {
    edge_bwd_dst_generic_kernel<T><<<blocks, kGW * 32, 0, st>>>(g.n_rows, g.row_ptr, g.col_idx, Pl, Pr, a, H, D, g.slopes, g.bias,
                                                                Hout, gH,
                                                                score, mx, sinv, gPr, reinterpret_cast<float*>(rec), RW,
                                                                ga_partials, galpha_dbg, g.ascale);
    edge_bwd_src_generic_kernel<T><<<(g.n_src + kGW - 1) / kGW, kGW * 32, 0, st>>>(
        g.n_src, g.csc_ptr, g.csc_dst, g.csc_eid, Pl, Pr, a, H, D, g.slopes, gH, reinterpret_cast<const float*>(rec), RW,
        gPl);
  });
  *n_partials = blocks * kGW;
  return 2;
}

int launch_unpack_rec_generic(const uint32_t* rec, int64_t E, int H, float* alpha, float* ge, cudaStream_t st) {
  if (E <= 0) return 0;
  int64_t blocks = (E * H + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  unpack_rec_generic_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(rec), E, H,
                                                         edge_generic_rec_words(H), alpha, ge);
  return 1;
}

}  // namespace gatx
