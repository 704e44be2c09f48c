// Fused edge passes of one GATv2 layer on projected features (the HBM-bound hot kernels).
//
// Forward  (replaces gatv2_edge_score_kernel EB:279-324, compute_max_sum_attn_score EB:326-359,
//           compute_attn_coeff EB:362-384, aggregate_kernel EB:386-424, postActivationLayerOutput
//           EB:426-459): one pass over a destination-sorted CSR row gathers P_l[src] with 128-bit
//           loads, forms e = a^T LReLU(P_l[src] + P_r[dst]) with warp shuffles, keeps an online
//           (running-max) softmax in registers and writes LReLU(sum alpha P_l[src]).  No atomics:
//           every output element has exactly one writer, so results are run-to-run deterministic.
// Backward (replaces kernel_grad_atten_coeff EB:612-651, compute_grad_attn_score_kernel EB:654-696 and
//           the per-edge parts of compute_grad_parameters_kernel EB:698-798 /
//           compute_features_input_gradients EB:801-874 / compute_preActivation_..._gradient EB:879-893):
//   pass 1, destination-major: galpha = g_h[dst].P_l[src]; ge = alpha (galpha - sum_seg alpha galpha) where
//           the segment sum equals gH[dst].Hout[dst] (so one pass suffices instead of the reference's
//           O(deg^2) loop); ga += ge LReLU(s); gP_r[dst] = sum ge a LReLU'(s); per-edge record
//           {sign(s) bits, alpha, ge} for pass 2.
//   pass 2, source-major over the stable transpose: gP_l[src] = sum alpha g_h[dst] + ge a LReLU'(s).
//
// Thread mapping: a feature row of F = H*D floats is F/4 float4 chunks.  LPR = min(32, F/4) lanes
// cover a row (lane li owns chunks li, li+LPR, ... -> every load instruction reads one contiguous
// 16*LPR-byte span), NV = F/(4*LPR) chunks per lane, and G = 32/LPR edges of the same row are
// processed side by side by sub-groups of the warp.  A head is D/4 consecutive chunks
// (requires D % 4 == 0, D <= 128, D/4 a power of two) so per-head dot products are xor-shuffle
// reductions over D/4 lanes.
#include "common.cuh"

namespace gatx {

namespace {

constexpr int kWarps = 8;        // warps per CTA for the row-per-warp kernels
constexpr int kHeavyWarps = 16;  // warps per CTA for the CTA-per-heavy-row kernels

struct Shape {
  int H, D, F, lph, lg_lph;  // lph = D/4 chunks (lanes) per head
  Slopes sl;                 // LeakyReLU slopes (set by the launchers from EdgeGraph::slopes)
  const float* bias;         // [F] or nullptr (EdgeGraph::bias)
  const float* ascale;       // [E][H] attention-dropout scale or nullptr (EdgeGraph::ascale)
};

__device__ __forceinline__ float head_reduce(float p, int lph) {
  for (int off = lph >> 1; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
  return p;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ------------------------------------------------------------------------------------ forward
template <int NV>
struct FwdState {
  float m[NV], s[NV];
  float4 acc[NV];
};

template <int NV, int LPR>
__device__ __forceinline__ void fwd_init(FwdState<NV>& st) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    st.m[j] = -1e9f;  // EB:336
    st.s[j] = 0.f;
    st.acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NV>
__device__ __forceinline__ void fwd_merge(FwdState<NV>& st, int j, float m2, float s2, float4 a2) {
  const float mn = fmaxf(st.m[j], m2);
  const float c1 = __expf(st.m[j] - mn), c2 = __expf(m2 - mn);
  st.s[j] = st.s[j] * c1 + s2 * c2;
  st.acc[j].x = st.acc[j].x * c1 + a2.x * c2;
  st.acc[j].y = st.acc[j].y * c1 + a2.y * c2;
  st.acc[j].z = st.acc[j].z * c1 + a2.z * c2;
  st.acc[j].w = st.acc[j].w * c1 + a2.w * c2;
  st.m[j] = mn;
}

// Processes edges first, first+step, ... (< end) of one row; `first` already includes this
// warp's offset, the sub-group offset is added here.  Two edges per sub-group are in flight.
template <int NV, int LPR>
__device__ __forceinline__ void fwd_range(FwdState<NV>& st, int first, int end, int step,
                                          const int* __restrict__ col_idx, const float* __restrict__ Pl,
                                          const Shape sh, const float4 (&pr)[NV], const float4 (&av)[NV],
                                          float* __restrict__ score, int li, int sub) {
  constexpr int G = 32 / LPR;
  const bool head_lane = (li & (sh.lph - 1)) == 0;
  for (int base = first; base < end; base += 2 * step) {
    const int e0 = base + sub, e1 = base + step + sub;
    const bool ok0 = e0 < end, ok1 = e1 < end;
    const int s0 = ok0 ? __ldg(col_idx + e0) : 0, s1 = ok1 ? __ldg(col_idx + e1) : 0;
    float4 v0[NV], v1[NV];
    const float* r0 = Pl + (int64_t)s0 * sh.F + 4 * li;
    const float* r1 = Pl + (int64_t)s1 * sh.F + 4 * li;
#pragma unroll
    for (int j = 0; j < NV; ++j) v0[j] = ok0 ? ldg4_stream(r0 + 4 * j * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NV; ++j) v1[j] = ok1 ? ldg4_stream(r1 + 4 * j * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float4(&v)[NV] = u == 0 ? v0 : v1;
      const bool ok = u == 0 ? ok0 : ok1;
      const int e = u == 0 ? e0 : e1;
      if (u == 1 && base + step >= end) break;  // warp-uniform
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float sa = sh.sl.attn;
        float p = av[j].x * lrelu(v[j].x + pr[j].x, sa) + av[j].y * lrelu(v[j].y + pr[j].y, sa) +
                  av[j].z * lrelu(v[j].z + pr[j].z, sa) + av[j].w * lrelu(v[j].w + pr[j].w, sa);
        p = head_reduce(p, sh.lph);
        if (ok) {
          if (head_lane) score[(int64_t)e * sh.H + ((li + j * LPR) >> sh.lg_lph)] = p;
          const float mn = fmaxf(st.m[j], p);
          const float corr = __expf(st.m[j] - mn), w = __expf(p - mn);
          st.s[j] = st.s[j] * corr + w;
          // attention dropout scales the aggregated term only; the softmax denominator keeps every edge
          const float wd = sh.ascale ? w * __ldg(sh.ascale + (int64_t)e * sh.H + ((li + j * LPR) >> sh.lg_lph)) : w;
          st.acc[j].x = st.acc[j].x * corr + wd * v[j].x;
          st.acc[j].y = st.acc[j].y * corr + wd * v[j].y;
          st.acc[j].z = st.acc[j].z * corr + wd * v[j].z;
          st.acc[j].w = st.acc[j].w * corr + wd * v[j].w;
          st.m[j] = mn;
        }
      }
    }
  }
  // merge the G sub-groups of the warp
  if (G > 1) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float m2 = __shfl_xor_sync(0xffffffffu, st.m[j], off);
        const float s2 = __shfl_xor_sync(0xffffffffu, st.s[j], off);
        float4 a2;
        a2.x = __shfl_xor_sync(0xffffffffu, st.acc[j].x, off);
        a2.y = __shfl_xor_sync(0xffffffffu, st.acc[j].y, off);
        a2.z = __shfl_xor_sync(0xffffffffu, st.acc[j].z, off);
        a2.w = __shfl_xor_sync(0xffffffffu, st.acc[j].w, off);
        fwd_merge<NV>(st, j, m2, s2, a2);
      }
    }
  }
}

template <int NV, int LPR>
__device__ __forceinline__ void fwd_finalize(const FwdState<NV>& st, int row, const Shape sh,
                                             float* __restrict__ Hout, float* __restrict__ hpre,
                                             float* __restrict__ mx, float* __restrict__ sinv, int li) {
  const bool head_lane = (li & (sh.lph - 1)) == 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float inv = 1.0f / (st.s[j] + 1e-8f);  // EB:379
    float4 h = make_float4(st.acc[j].x * inv, st.acc[j].y * inv, st.acc[j].z * inv, st.acc[j].w * inv);
    if (sh.bias) {
      const float4 b = ldg4(sh.bias + 4 * (li + j * LPR));
      h = make_float4(h.x + b.x, h.y + b.y, h.z + b.z, h.w + b.w);
    }
    const int64_t off = (int64_t)row * sh.F + 4 * (li + j * LPR);
    if (hpre) st4(hpre + off, h);
    const float sc = sh.sl.act;
    st4(Hout + off, make_float4(lrelu(h.x, sc), lrelu(h.y, sc), lrelu(h.z, sc), lrelu(h.w, sc)));
    if (head_lane) {
      const int hd = (li + j * LPR) >> sh.lg_lph;
      mx[(int64_t)row * sh.H + hd] = st.m[j];
      sinv[(int64_t)row * sh.H + hd] = inv;
    }
  }
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kWarps * 32)
edge_fwd_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                const float* __restrict__ Pl, const float* __restrict__ Pr, const float* __restrict__ a,
                Shape sh, float* __restrict__ Hout, float* __restrict__ hpre, float* __restrict__ score,
                float* __restrict__ mx, float* __restrict__ sinv) {
  constexpr int G = 32 / LPR;
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  if (end - beg > kHeavyDeg) return;  // CTA-per-row kernel
  float4 pr[NV], av[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    pr[j] = ldg4(Pr + (int64_t)row * sh.F + 4 * (li + j * LPR));
    av[j] = ldg4(a + 4 * (li + j * LPR));
  }
  FwdState<NV> st;
  fwd_init<NV, LPR>(st);
  fwd_range<NV, LPR>(st, beg, end, G, col_idx, Pl, sh, pr, av, score, li, sub);
  if (sub == 0) fwd_finalize<NV, LPR>(st, row, sh, Hout, hpre, mx, sinv, li);
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kHeavyWarps * 32)
edge_fwd_heavy_kernel(const int* __restrict__ heavy_rows, const int* __restrict__ row_ptr,
                      const int* __restrict__ col_idx, const float* __restrict__ Pl,
                      const float* __restrict__ Pr, const float* __restrict__ a, Shape sh,
                      float* __restrict__ Hout, float* __restrict__ hpre, float* __restrict__ score,
                      float* __restrict__ mx, float* __restrict__ sinv) {
  constexpr int G = 32 / LPR, CH = NV * LPR;
  extern __shared__ __align__(16) float smem[];
  float* sm_m = smem;                                                       // [W][CH]
  float* sm_s = sm_m + kHeavyWarps * CH;                                    // [W][CH]
  float4* sm_a = reinterpret_cast<float4*>(sm_s + kHeavyWarps * CH);        // [W][CH]
  const int row = heavy_rows[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  float4 pr[NV], av[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    pr[j] = ldg4(Pr + (int64_t)row * sh.F + 4 * (li + j * LPR));
    av[j] = ldg4(a + 4 * (li + j * LPR));
  }
  FwdState<NV> st;
  fwd_init<NV, LPR>(st);
  fwd_range<NV, LPR>(st, beg + w * G, end, kHeavyWarps * G, col_idx, Pl, sh, pr, av, score, li, sub);
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = li + j * LPR;
      sm_m[w * CH + c] = st.m[j];
      sm_s[w * CH + c] = st.s[j];
      sm_a[w * CH + c] = st.acc[j];
    }
  }
  __syncthreads();
  if (w == 0 && sub == 0) {
    for (int ww = 1; ww < kHeavyWarps; ++ww)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = li + j * LPR;
        fwd_merge<NV>(st, j, sm_m[ww * CH + c], sm_s[ww * CH + c], sm_a[ww * CH + c]);
      }
    fwd_finalize<NV, LPR>(st, row, sh, Hout, hpre, mx, sinv, li);
  }
}

// ------------------------------------------------------------------------- backward, pass 1
// per-edge record: [4*NV mask words][H alpha][H ge], padded to a multiple of 4 words
__host__ __device__ inline int rec_words(int H, int NV) { return (4 * NV + 2 * H + 3) / 4 * 4; }

template <int NV>
struct Bwd1Row {
  float4 gh[NV], pr[NV], gpr[NV];
  float c[NV], m[NV], inv[NV];
};

template <int NV, int LPR>
__device__ __forceinline__ void bwd1_load_row(Bwd1Row<NV>& r, int row, const Shape sh,
                                              const float* __restrict__ Pr, const float* __restrict__ Hout,
                                              const float* __restrict__ gH, const float* __restrict__ mx,
                                              const float* __restrict__ sinv, int li) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int64_t off = (int64_t)row * sh.F + 4 * (li + j * LPR);
    const float4 g = *reinterpret_cast<const float4*>(gH + off);  // rewritten in place: no nc path
    const float4 ho = ldg4(Hout + off);
    r.pr[j] = ldg4(Pr + off);
    // EB:879-893 / EB:599: gradient through the activation, LReLU'(h) has the sign of LReLU(h)
    const float sc = sh.sl.act;
    r.gh[j] = make_float4(g.x * lrelu_grad(ho.x, sc), g.y * lrelu_grad(ho.y, sc), g.z * lrelu_grad(ho.z, sc),
                          g.w * lrelu_grad(ho.w, sc));
    // sum_seg alpha*galpha = g_pre . (h - bias) = gH . Hout - g_pre . bias  because LReLU'(h) * h = LReLU(h)
    float cd = dot4(g, ho);
    if (sh.bias) cd -= dot4(r.gh[j], ldg4(sh.bias + 4 * (li + j * LPR)));
    r.c[j] = head_reduce(cd, sh.lph);
    const int hd = (li + j * LPR) >> sh.lg_lph;
    r.m[j] = __ldg(mx + (int64_t)row * sh.H + hd);
    r.inv[j] = __ldg(sinv + (int64_t)row * sh.H + hd);
    r.gpr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NV, int LPR>
__device__ __forceinline__ void bwd1_range(Bwd1Row<NV>& r, float4 (&ga)[NV], int first, int end, int step,
                                           const int* __restrict__ col_idx, const float* __restrict__ Pl,
                                           const Shape sh, const float4 (&av)[NV],
                                           const float* __restrict__ score, uint32_t* __restrict__ rec,
                                           float* __restrict__ galpha_dbg, int li, int sub) {
  const int RW = rec_words(sh.H, NV);
  const bool head_lane = (li & (sh.lph - 1)) == 0;
  const uint32_t sub_mask = LPR == 32 ? 0xffffffffu : ((1u << LPR) - 1u);
  for (int base = first; base < end; base += 2 * step) {
    const int e0 = base + sub, e1 = base + step + sub;
    const bool ok0 = e0 < end, ok1 = e1 < end;
    const int s0 = ok0 ? __ldg(col_idx + e0) : 0, s1 = ok1 ? __ldg(col_idx + e1) : 0;
    float4 v0[NV], v1[NV];
    float sc0[NV], sc1[NV];
    const float* r0 = Pl + (int64_t)s0 * sh.F + 4 * li;
    const float* r1 = Pl + (int64_t)s1 * sh.F + 4 * li;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v0[j] = ok0 ? ldg4_stream(r0 + 4 * j * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
      sc0[j] = ok0 ? __ldg(score + (int64_t)e0 * sh.H + ((li + j * LPR) >> sh.lg_lph)) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v1[j] = ok1 ? ldg4_stream(r1 + 4 * j * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
      sc1[j] = ok1 ? __ldg(score + (int64_t)e1 * sh.H + ((li + j * LPR) >> sh.lg_lph)) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && base + step >= end) break;  // warp-uniform
      const float4(&v)[NV] = u == 0 ? v0 : v1;
      const float(&sc)[NV] = u == 0 ? sc0 : sc1;
      const bool ok = u == 0 ? ok0 : ok1;
      const int e = u == 0 ? e0 : e1;
      uint32_t* re = rec + (int64_t)e * RW;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float galpha = head_reduce(dot4(r.gh[j], v[j]), sh.lph);
        float alpha = __expf(sc[j] - r.m[j]) * r.inv[j];  // EB:378-379
        const float dsc = (sh.ascale && ok) ? __ldg(sh.ascale + (int64_t)e * sh.H + ((li + j * LPR) >> sh.lg_lph)) : 1.f;
        galpha *= dsc;                                          // h = sum alpha * dsc * P_l
        const float ge = ok ? alpha * (galpha - r.c[j]) : 0.f;  // EB:689-690 in closed form
        alpha *= dsc;                                           // pass 2 aggregates g_h with alpha * dsc
        const float sx = v[j].x + r.pr[j].x, sy = v[j].y + r.pr[j].y, sz = v[j].z + r.pr[j].z,
                    sw = v[j].w + r.pr[j].w;
        // ga += ge * LReLU(s)  (EB:769)
        const float sa = sh.sl.attn;
        ga[j].x += ge * lrelu(sx, sa); ga[j].y += ge * lrelu(sy, sa);
        ga[j].z += ge * lrelu(sz, sa); ga[j].w += ge * lrelu(sw, sa);
        // m = ge * a * LReLU'(s)  (EB:774-775) accumulated for the destination
        r.gpr[j].x += ge * av[j].x * lrelu_grad(sx, sa); r.gpr[j].y += ge * av[j].y * lrelu_grad(sy, sa);
        r.gpr[j].z += ge * av[j].z * lrelu_grad(sz, sa); r.gpr[j].w += ge * av[j].w * lrelu_grad(sw, sa);
        const uint32_t bx = __ballot_sync(0xffffffffu, sx > 0.f), by = __ballot_sync(0xffffffffu, sy > 0.f),
                       bz = __ballot_sync(0xffffffffu, sz > 0.f), bw = __ballot_sync(0xffffffffu, sw > 0.f);
        if (ok) {
          if (li == 0) {
            const int sh_ = sub * LPR;
            *reinterpret_cast<uint4*>(re + 4 * j) = make_uint4((bx >> sh_) & sub_mask, (by >> sh_) & sub_mask,
                                                               (bz >> sh_) & sub_mask, (bw >> sh_) & sub_mask);
          }
          if (head_lane) {
            const int hd = (li + j * LPR) >> sh.lg_lph;
            re[4 * NV + hd] = __float_as_uint(alpha);
            re[4 * NV + sh.H + hd] = __float_as_uint(ge);
            if (galpha_dbg) galpha_dbg[(int64_t)e * sh.H + hd] = galpha;
          }
        }
      }
    }
  }
}

template <int NV, int LPR>
__device__ __forceinline__ void subgroup_sum(float4 (&x)[NV]) {
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      x[j].x += __shfl_xor_sync(0xffffffffu, x[j].x, off);
      x[j].y += __shfl_xor_sync(0xffffffffu, x[j].y, off);
      x[j].z += __shfl_xor_sync(0xffffffffu, x[j].z, off);
      x[j].w += __shfl_xor_sync(0xffffffffu, x[j].w, off);
    }
}

// block-level deterministic sum of per-warp float4[NV] accumulators -> out[CH*4] (lane layout li/j)
template <int NV, int LPR, int W>
__device__ __forceinline__ void block_sum_write(float4 (&x)[NV], float4* sm /*[W][CH]*/, float* __restrict__ out,
                                                int w, int li, int sub) {
  constexpr int CH = NV * LPR;
  subgroup_sum<NV, LPR>(x);
  if (sub == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) sm[w * CH + li + j * LPR] = x[j];
  __syncthreads();
  if (w == 0 && sub == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 t = sm[li + j * LPR];
      for (int ww = 1; ww < W; ++ww) {
        const float4 o = sm[ww * CH + li + j * LPR];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      st4(out + 4 * (li + j * LPR), t);
    }
  }
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kWarps * 32)
edge_bwd_dst_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                    const float* __restrict__ Pl, const float* __restrict__ Pr, const float* __restrict__ a,
                    Shape sh, const float* __restrict__ Hout, float* __restrict__ gH,
                    const float* __restrict__ score, const float* __restrict__ mx,
                    const float* __restrict__ sinv, float* __restrict__ gPr, uint32_t* __restrict__ rec,
                    float* __restrict__ ga_partials, float* __restrict__ galpha_dbg) {
  constexpr int G = 32 / LPR, CH = NV * LPR;
  __shared__ __align__(16) float4 sm[kWarps * CH];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  float4 av[NV], ga[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    av[j] = ldg4(a + 4 * (li + j * LPR));
    ga[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int total_warps = gridDim.x * kWarps;
  // rows are dealt to CTAs in contiguous groups of kWarps so that neighbouring warps stream
  // neighbouring CSR rows
  for (int row = blockIdx.x * kWarps + w; row < n_rows; row += total_warps) {
    const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
    if (end - beg > kHeavyDeg) continue;
    Bwd1Row<NV> r;
    bwd1_load_row<NV, LPR>(r, row, sh, Pr, Hout, gH, mx, sinv, li);
    bwd1_range<NV, LPR>(r, ga, beg, end, G, col_idx, Pl, sh, av, score, rec, galpha_dbg, li, sub);
    subgroup_sum<NV, LPR>(r.gpr);
    if (sub == 0)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t off = (int64_t)row * sh.F + 4 * (li + j * LPR);
        st4(gPr + off, r.gpr[j]);
        st4(gH + off, r.gh[j]);  // pre-activation gradient, gathered by pass 2
      }
  }
  block_sum_write<NV, LPR, kWarps>(ga, sm, ga_partials + (int64_t)blockIdx.x * sh.F, w, li, sub);
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kHeavyWarps * 32)
edge_bwd_dst_heavy_kernel(const int* __restrict__ heavy_rows, const int* __restrict__ row_ptr,
                          const int* __restrict__ col_idx, const float* __restrict__ Pl,
                          const float* __restrict__ Pr, const float* __restrict__ a, Shape sh,
                          const float* __restrict__ Hout, float* __restrict__ gH,
                          const float* __restrict__ score, const float* __restrict__ mx,
                          const float* __restrict__ sinv, float* __restrict__ gPr, uint32_t* __restrict__ rec,
                          float* __restrict__ ga_partials, float* __restrict__ galpha_dbg) {
  constexpr int G = 32 / LPR, CH = NV * LPR;
  extern __shared__ __align__(16) float smem[];
  float4* sm = reinterpret_cast<float4*>(smem);  // [kHeavyWarps][CH]
  const int row = heavy_rows[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  float4 av[NV], ga[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    av[j] = ldg4(a + 4 * (li + j * LPR));
    ga[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  Bwd1Row<NV> r;
  bwd1_load_row<NV, LPR>(r, row, sh, Pr, Hout, gH, mx, sinv, li);
  __syncthreads();  // every warp has read gH[row] before warp 0 overwrites it
  if (w == 0 && sub == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) st4(gH + (int64_t)row * sh.F + 4 * (li + j * LPR), r.gh[j]);
  bwd1_range<NV, LPR>(r, ga, beg + w * G, end, kHeavyWarps * G, col_idx, Pl, sh, av, score, rec, galpha_dbg, li,
                      sub);
  block_sum_write<NV, LPR, kHeavyWarps>(r.gpr, sm, gPr + (int64_t)row * sh.F, w, li, sub);
  __syncthreads();
  block_sum_write<NV, LPR, kHeavyWarps>(ga, sm, ga_partials + (int64_t)blockIdx.x * sh.F, w, li, sub);
}

// ------------------------------------------------------------------------- backward, pass 2
template <int NV, int LPR>
__device__ __forceinline__ void bwd2_range(float4 (&acc)[NV], int first, int end, int step,
                                           const int* __restrict__ csc_dst, const int* __restrict__ csc_eid,
                                           const float* __restrict__ gH, const uint32_t* __restrict__ rec,
                                           const Shape sh, const float4 (&av)[NV], int li, int sub) {
  const int RW = rec_words(sh.H, NV);
  for (int base = first; base < end; base += 2 * step) {
    const int q0 = base + sub, q1 = base + step + sub;
    const bool ok0 = q0 < end, ok1 = q1 < end;
    const int d0 = ok0 ? __ldg(csc_dst + q0) : 0, d1 = ok1 ? __ldg(csc_dst + q1) : 0;
    const int e0 = ok0 ? __ldg(csc_eid + q0) : 0, e1 = ok1 ? __ldg(csc_eid + q1) : 0;
    float4 g0[NV], g1[NV];
    uint4 k0[NV], k1[NV];
    float al0[NV], al1[NV], ge0[NV], ge1[NV];
    const float* r0 = gH + (int64_t)d0 * sh.F + 4 * li;
    const float* r1 = gH + (int64_t)d1 * sh.F + 4 * li;
    const uint32_t* c0 = rec + (int64_t)e0 * RW;
    const uint32_t* c1 = rec + (int64_t)e1 * RW;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int hd = (li + j * LPR) >> sh.lg_lph;
      if (ok0) {
        g0[j] = ldg4_stream(r0 + 4 * j * LPR);
        k0[j] = __ldg(reinterpret_cast<const uint4*>(c0 + 4 * j));
        al0[j] = __uint_as_float(__ldg(c0 + 4 * NV + hd));
        ge0[j] = __uint_as_float(__ldg(c0 + 4 * NV + sh.H + hd));
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int hd = (li + j * LPR) >> sh.lg_lph;
      if (ok1) {
        g1[j] = ldg4_stream(r1 + 4 * j * LPR);
        k1[j] = __ldg(reinterpret_cast<const uint4*>(c1 + 4 * j));
        al1[j] = __uint_as_float(__ldg(c1 + 4 * NV + hd));
        ge1[j] = __uint_as_float(__ldg(c1 + 4 * NV + sh.H + hd));
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (ok0) {
        const float al = al0[j], ge = ge0[j];
        // EB:865-866: g_h[dst] * alpha + ge * a * LReLU'(s), LReLU'(s) from the recorded sign bit
        acc[j].x += al * g0[j].x + ge * av[j].x * (((k0[j].x >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].y += al * g0[j].y + ge * av[j].y * (((k0[j].y >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].z += al * g0[j].z + ge * av[j].z * (((k0[j].z >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].w += al * g0[j].w + ge * av[j].w * (((k0[j].w >> li) & 1u) ? 1.f : sh.sl.attn);
      }
      if (ok1) {
        const float al = al1[j], ge = ge1[j];
        acc[j].x += al * g1[j].x + ge * av[j].x * (((k1[j].x >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].y += al * g1[j].y + ge * av[j].y * (((k1[j].y >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].z += al * g1[j].z + ge * av[j].z * (((k1[j].z >> li) & 1u) ? 1.f : sh.sl.attn);
        acc[j].w += al * g1[j].w + ge * av[j].w * (((k1[j].w >> li) & 1u) ? 1.f : sh.sl.attn);
      }
    }
  }
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kWarps * 32)
edge_bwd_src_kernel(int n_src, const int* __restrict__ csc_ptr, const int* __restrict__ csc_dst,
                    const int* __restrict__ csc_eid, const float* __restrict__ a, Shape sh,
                    const float* __restrict__ gH, const uint32_t* __restrict__ rec, float* __restrict__ gPl) {
  constexpr int G = 32 / LPR;
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_src) return;
  const int lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  const int beg = __ldg(csc_ptr + row), end = __ldg(csc_ptr + row + 1);
  if (end - beg > kHeavyDeg) return;
  float4 av[NV], acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    av[j] = ldg4(a + 4 * (li + j * LPR));
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  bwd2_range<NV, LPR>(acc, beg, end, G, csc_dst, csc_eid, gH, rec, sh, av, li, sub);
  subgroup_sum<NV, LPR>(acc);
  if (sub == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) st4(gPl + (int64_t)row * sh.F + 4 * (li + j * LPR), acc[j]);
}

template <int NV, int LPR>
__global__ void __launch_bounds__(kHeavyWarps * 32)
edge_bwd_src_heavy_kernel(const int* __restrict__ heavy_srcs, const int* __restrict__ csc_ptr,
                          const int* __restrict__ csc_dst, const int* __restrict__ csc_eid,
                          const float* __restrict__ a, Shape sh, const float* __restrict__ gH,
                          const uint32_t* __restrict__ rec, float* __restrict__ gPl) {
  constexpr int G = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  float4* sm = reinterpret_cast<float4*>(smem);
  const int row = heavy_srcs[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, li = lane % LPR, sub = lane / LPR;
  const int beg = __ldg(csc_ptr + row), end = __ldg(csc_ptr + row + 1);
  float4 av[NV], acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    av[j] = ldg4(a + 4 * (li + j * LPR));
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  bwd2_range<NV, LPR>(acc, beg + w * G, end, kHeavyWarps * G, csc_dst, csc_eid, gH, rec, sh, av, li, sub);
  block_sum_write<NV, LPR, kHeavyWarps>(acc, sm, gPl + (int64_t)row * sh.F, w, li, sub);
}

// ------------------------------------------------------------------------------- small kernels
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_partials, int n,
                                       float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < n_partials; ++b) s += partials[(int64_t)b * n + i];
  out[i] = accumulate ? out[i] + s : s;
}

// column sums of a row-major [n_rows][cols] matrix: block b sums rows b, b + gridDim.x, ... into partials[b][cols]
__global__ void colsum_partial_kernel(const float* __restrict__ M, int n_rows, int cols, float* __restrict__ partials) {
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) s += M[(int64_t)r * cols + c];
    partials[(int64_t)blockIdx.x * cols + c] = s;
  }
}

__global__ void unpack_rec_kernel(const uint32_t* __restrict__ rec, int64_t E, int H, int NV, int RW,
                                  float* __restrict__ alpha, float* __restrict__ ge) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E * H; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / H;
    const int h = (int)(i % H);
    if (alpha) alpha[i] = __uint_as_float(rec[e * RW + 4 * NV + h]);
    if (ge) ge[i] = __uint_as_float(rec[e * RW + 4 * NV + H + h]);
  }
}

__global__ void alpha_from_score_kernel(const float* __restrict__ score, const int* __restrict__ coo_dst,
                                        const float* __restrict__ mx, const float* __restrict__ sinv, int64_t E,
                                        int H, float* __restrict__ alpha) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E * H; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = coo_dst[i / H];
    const int h = (int)(i % H);
    alpha[i] = __expf(score[i] - mx[d * H + h]) * sinv[d * H + h];  // EB:378-379
  }
}

// last layer with H > 1 (extension, EB:440-449 semantics): Hout[n][k] = mean_h Hfull[n][h][k]
__global__ void head_mean_kernel(const float* __restrict__ Hfull, int N, int H, int D, float* __restrict__ Hout) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)N * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / D;
    const int k = (int)(i % D);
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += Hfull[(n * H + h) * D + k];
    Hout[i] = s / (float)H;
  }
}
__global__ void head_bcast_grad_kernel(const float* __restrict__ gHout, int N, int H, int D,
                                       float* __restrict__ gHfull) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)N * H * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / ((int64_t)H * D);
    const int k = (int)(i % D);
    gHfull[i] = gHout[n * D + k] * (1.0f / (float)H);  // EB:597-602
  }
}

bool make_shape(int H, int D, Shape* sh, int* nv, int* lpr) {
  if (H < 1 || D < 4 || D % 4 || D > 128 || H > 32) return false;
  const int lph = D / 4;
  if (lph & (lph - 1)) return false;
  const int F = H * D, CH = F / 4;
  int LPR, NV;
  if (CH < 32) {
    if (CH & (CH - 1)) return false;
    LPR = CH;
    NV = 1;
  } else {
    if (CH % 32) return false;
    LPR = 32;
    NV = CH / 32;
    if (NV != 1 && NV != 2 && NV != 4 && NV != 8) return false;
  }
  int lg = 0;
  while ((1 << lg) < lph) ++lg;
  if (sh) *sh = Shape{H, D, F, lph, lg, Slopes{kSlope, kSlope}, nullptr};
  if (nv) *nv = NV;
  if (lpr) *lpr = LPR;
  return true;
}

// dispatch over the (NV, LPR) instantiations
#define GATX_DISPATCH(NVv, LPRv, ...)                               \
  do {                                                              \
    if (LPRv == 32) {                                               \
      if (NVv == 1) { constexpr int NV = 1, LPR = 32; __VA_ARGS__; }       \
      else if (NVv == 2) { constexpr int NV = 2, LPR = 32; __VA_ARGS__; }  \
      else if (NVv == 4) { constexpr int NV = 4, LPR = 32; __VA_ARGS__; }  \
      else { constexpr int NV = 8, LPR = 32; __VA_ARGS__; }                \
    } else if (LPRv == 16) { constexpr int NV = 1, LPR = 16; __VA_ARGS__; } \
    else if (LPRv == 8) { constexpr int NV = 1, LPR = 8; __VA_ARGS__; }    \
    else if (LPRv == 4) { constexpr int NV = 1, LPR = 4; __VA_ARGS__; }    \
    else if (LPRv == 2) { constexpr int NV = 1, LPR = 2; __VA_ARGS__; }    \
    else { constexpr int NV = 1, LPR = 1; __VA_ARGS__; }                   \
  } while (0)

template <typename K>
void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

bool edge_shape_supported(int H, int D) { return make_shape(H, D, nullptr, nullptr, nullptr); }
int edge_rec_words(int H, int D) {
  int nv, lpr;
  if (!make_shape(H, D, nullptr, &nv, &lpr)) return -1;
  return rec_words(H, nv);
}

int launch_edge_forward(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                        float* Hout, float* hpre, float* score, float* mx, float* sinv, cudaStream_t st) {
  Shape sh;
  int nv, lpr, launches = 0;
  if (!make_shape(H, D, &sh, &nv, &lpr)) return -1;
  sh.sl = g.slopes;
  sh.bias = g.bias;
  sh.ascale = g.ascale;
  if (g.n_rows <= 0) return 0;
  GATX_DISPATCH(nv, lpr, {
    edge_fwd_kernel<NV, LPR><<<(g.n_rows + kWarps - 1) / kWarps, kWarps * 32, 0, st>>>(
        g.n_rows, g.row_ptr, g.col_idx, Pl, Pr, a, sh, Hout, hpre, score, mx, sinv);
    ++launches;
    if (g.n_heavy_rows > 0) {
      const size_t smem = (size_t)kHeavyWarps * NV * LPR * (4 + 4 + 16);
      allow_smem(edge_fwd_heavy_kernel<NV, LPR>, smem);
      edge_fwd_heavy_kernel<NV, LPR><<<g.n_heavy_rows, kHeavyWarps * 32, smem, st>>>(
          g.heavy_rows, g.row_ptr, g.col_idx, Pl, Pr, a, sh, Hout, hpre, score, mx, sinv);
      ++launches;
    }
  });
  return launches;
}

int launch_edge_backward_dst(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr,
                             const float* a, const float* Hout, float* gH, const float* score,
                             const float* mx, const float* sinv, float* gPr, uint32_t* rec,
                             float* ga_partials, int* n_partials, float* galpha_dbg, cudaStream_t st) {
  Shape sh;
  int nv, lpr, launches = 0;
  if (!make_shape(H, D, &sh, &nv, &lpr)) return -1;
  sh.sl = g.slopes;
  sh.bias = g.bias;
  sh.ascale = g.ascale;
  *n_partials = 0;
  if (g.n_rows <= 0) return 0;
  GATX_DISPATCH(nv, lpr, {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_bwd_dst_kernel<NV, LPR>, kWarps * 32, 0);
    if (per_sm < 1) per_sm = 1;
    int blocks = kNumSMs * per_sm;
    if (blocks > kEdgeBwdBlocks) blocks = kEdgeBwdBlocks;
    const int need = (g.n_rows + kWarps - 1) / kWarps;
    if (blocks > need) blocks = need;
    edge_bwd_dst_kernel<NV, LPR><<<blocks, kWarps * 32, 0, st>>>(g.n_rows, g.row_ptr, g.col_idx, Pl, Pr, a, sh,
                                                                 Hout, gH, score, mx, sinv, gPr, rec,
                                                                 ga_partials, galpha_dbg);
    ++launches;
    *n_partials = blocks;
    if (g.n_heavy_rows > 0) {
      const size_t smem = (size_t)kHeavyWarps * NV * LPR * 16;
      allow_smem(edge_bwd_dst_heavy_kernel<NV, LPR>, smem);
      edge_bwd_dst_heavy_kernel<NV, LPR><<<g.n_heavy_rows, kHeavyWarps * 32, smem, st>>>(
          g.heavy_rows, g.row_ptr, g.col_idx, Pl, Pr, a, sh, Hout, gH, score, mx, sinv, gPr, rec,
          ga_partials + (int64_t)blocks * sh.F, galpha_dbg);
      ++launches;
      *n_partials += g.n_heavy_rows;
    }
  });
  return launches;
}

int launch_edge_backward_src(const EdgeGraph& g, int H, int D, const float* a, const float* gH,
                             const uint32_t* rec, float* gPl, cudaStream_t st) {
  Shape sh;
  int nv, lpr, launches = 0;
  if (!make_shape(H, D, &sh, &nv, &lpr)) return -1;
  sh.sl = g.slopes;
  sh.bias = g.bias;
  sh.ascale = g.ascale;
  if (g.n_src <= 0) return 0;
  GATX_DISPATCH(nv, lpr, {
    edge_bwd_src_kernel<NV, LPR><<<(g.n_src + kWarps - 1) / kWarps, kWarps * 32, 0, st>>>(
        g.n_src, g.csc_ptr, g.csc_dst, g.csc_eid, a, sh, gH, rec, gPl);
    ++launches;
    if (g.n_heavy_srcs > 0) {
      const size_t smem = (size_t)kHeavyWarps * NV * LPR * 16;
      allow_smem(edge_bwd_src_heavy_kernel<NV, LPR>, smem);
      edge_bwd_src_heavy_kernel<NV, LPR><<<g.n_heavy_srcs, kHeavyWarps * 32, smem, st>>>(
          g.heavy_srcs, g.csc_ptr, g.csc_dst, g.csc_eid, a, sh, gH, rec, gPl);
      ++launches;
    }
  });
  return launches;
}

int launch_reduce_partials(const float* partials, int n_partials, int n, float* out, bool accumulate,
                           cudaStream_t st) {
  if (n <= 0) return 0;
  reduce_partials_kernel<<<(n + 127) / 128, 128, 0, st>>>(partials, n_partials, n, out, accumulate ? 1 : 0);
  return 1;
}

int launch_colsum(const float* M, int n_rows, int cols, float* partials, float* out, bool accumulate, cudaStream_t st) {
  if (cols <= 0) return 0;
  int blocks = n_rows < kColsumBlocks ? (n_rows > 0 ? n_rows : 1) : kColsumBlocks;
  colsum_partial_kernel<<<blocks, 256, 0, st>>>(M, n_rows, cols, partials);
  return 1 + launch_reduce_partials(partials, blocks, cols, out, accumulate, st);
}

int launch_unpack_rec(const uint32_t* rec, int64_t E, int H, int D, float* alpha, float* ge, cudaStream_t st) {
  int nv, lpr;
  if (!make_shape(H, D, nullptr, &nv, &lpr)) return -1;
  if (E <= 0) return 0;
  int64_t blocks = (E * H + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  unpack_rec_kernel<<<(int)blocks, 256, 0, st>>>(rec, E, H, nv, rec_words(H, nv), alpha, ge);
  return 1;
}

int launch_alpha_from_score(const float* score, const int* coo_dst, const float* mx, const float* sinv, int64_t E,
                            int H, float* alpha, cudaStream_t st) {
  if (E <= 0) return 0;
  int64_t blocks = (E * H + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  alpha_from_score_kernel<<<(int)blocks, 256, 0, st>>>(score, coo_dst, mx, sinv, E, H, alpha);
  return 1;
}

int launch_head_mean(const float* Hfull, int N, int H, int D, float* Hout, cudaStream_t st) {
  int64_t blocks = ((int64_t)N * D + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  head_mean_kernel<<<(int)blocks, 256, 0, st>>>(Hfull, N, H, D, Hout);
  return 1;
}
int launch_head_bcast_grad(const float* gHout, int N, int H, int D, float* gHfull, cudaStream_t st) {
  int64_t blocks = ((int64_t)N * H * D + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  head_bcast_grad_kernel<<<(int)blocks, 256, 0, st>>>(gHout, N, H, D, gHfull);
  return 1;
}

}  // namespace gatx
