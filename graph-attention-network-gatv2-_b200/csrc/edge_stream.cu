// Edge-balanced streaming versions of the fused edge passes for feature rows of >= 128 floats.
//
// The CSR edge array is cut into chunks of T consecutive edges; one warp owns one chunk at a time
// (grid-stride, static assignment -> deterministic) and streams it:
//   * the gathered rows (P_l[src] forward / pass 1, g_h[dst] + the per-edge record in pass 2) are fetched
//     with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) into a per-warp shared-memory ring of R
//     slots, completion tracked by one mbarrier per slot -- R rows are always in flight per warp without
//     holding them in registers;
//   * destination rows are segments of the chunk; a segment that covers its whole row is finalised in
//     place, a segment of a row that straddles a chunk boundary writes its partial state to a side buffer
//     and a fix-up kernel merges the pieces in chunk order.
// Every warp processes the same number of edges regardless of the degree distribution, so power-law hubs
// need no special casing, and there are no atomics anywhere.
//
// Same math and reference citations as edge_kernels.cu (which keeps the narrow-row shapes).
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace gatx {
namespace {

constexpr int kSW = 4;  // warps per CTA
// -DGATX_RING_FENCE: fence.proxy.async before every per-edge refill of a bulk-copy ring slot (the slot was read by generic
// loads one instruction earlier; see the row buffer of pass 1).  Off by default: never observed to matter (all-edge checks at
// products size), costs an instruction per edge -- to be measured.
#ifdef GATX_RING_FENCE
#define RING_REFILL_FENCE() fence_proxy_async_smem()
#else
#define RING_REFILL_FENCE()
#endif

struct Shape {
  int H, D, F, lph, lg_lph;  // lph = D / 4: lanes per head in the node-wise kernels' layout (lane + 32 j)
  int lc, lg_lc;             // lc = 32 / H: lanes per head in the lane-contiguous layout of the edge passes
  Slopes slopes;             // LeakyReLU slopes (set by the launchers from EdgeGraph::slopes)
  const float* bias;         // [F] or nullptr (EdgeGraph::bias)
};

struct StreamGraph {
  int E, T, n_chunks, n_rows;
  const int* row_ptr;    // [n_rows + 1]
  const int* chunk_row;  // [n_chunks] row that contains edge c*T
  uint32_t hot;          // bit of a gather index that marks an L2-resident ("hot") row; 0 = no hints in the indices
  uint32_t idx_mask;     // gather index = raw & idx_mask
  Slopes slopes;         // LeakyReLU slopes: attention score / layer activation
  const float* bias;     // [F] added to the aggregate before the activation, or nullptr
  const float* ascale;   // [E][H] attention-dropout scale of the view's edges, or nullptr (EdgeGraph::ascale)
};
// L2 policies of the gathers: hot rows evict_last, everything streamed once evict_first (plain when hints are off)
struct GatherPolicy {
  uint64_t hot, cold;
  uint32_t bit, mask;
  __device__ __forceinline__ GatherPolicy(const StreamGraph& g) : bit(g.hot), mask(g.idx_mask) {
    hot = g.hot ? l2_policy_evict_last() : l2_policy_evict_normal();
    cold = g.hot ? l2_policy_evict_first() : l2_policy_evict_normal();
  }
  __device__ __forceinline__ uint64_t of(int raw) const { return ((uint32_t)raw & bit) ? hot : cold; }
  __device__ __forceinline__ int64_t id(int raw) const { return (int64_t)((uint32_t)raw & mask); }
};

__device__ __forceinline__ float head_reduce(float p, int lph) {
  for (int off = lph >> 1; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
  return p;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// LeakyReLU with 0 < slope < 1 is max(x, slope*x): FMUL + FMNMX
// LeakyReLU for 0 <= slope < 1 (gatx_create rejects anything else): two instructions, no select
__device__ __forceinline__ float lrelu_fast(float x, float slope) { return fmaxf(x, slope * x); }
__device__ __forceinline__ float shx(float v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
// Packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE-rn fp32 operations per issued instruction).  The
// streaming loops are bound by instruction issue at the power-capped clock, and half of their instructions are
// elementwise fp32 math on LDS.128 results, whose components already sit in aligned register pairs.
// -DGATX_SCALAR_FP32 builds the same arithmetic (same operations, same rounding, bit-identical results) from scalar
// FADD / FMUL / FFMA: the A/B baseline of tools/ab_packed_fp32.sh.
#ifdef GATX_SCALAR_FP32
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
#else
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
#endif
// Scalar forms for the kernels where packing loses.  Same-box A/B against -DGATX_SCALAR_FP32 (tools/ab_packed_fp32.sh,
// profiles/r1_ab_packed_fp32.txt): packed is 3.5 % faster in edge_fwd_stream_kernel (512-float rows), 6-9 % in
// edge_fwd_pair_kernel, 2-5 % in edge_bwd_dst_pair_kernel, but 2-5 % SLOWER in edge_bwd_{dst,src}_stream_kernel and
// edge_bwd_src_pair_kernel, whose bodies interleave a select per element (FSETP / FSEL, bit tests) with the arithmetic
// and pay for the aligned-register-pair constraint -- even with only the galpha dot product packed.  Those stay scalar.
__device__ __forceinline__ float2 fma2_s(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 add2_s(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 lrelu_fast2(float2 x, float2 slope) {
  const float2 t = mul2(x, slope);
  return make_float2(fmaxf(x.x, t.x), fmaxf(x.y, t.y));
}
// predicated global stores: one STG with a predicate instead of a divergent branch region per store
__device__ __forceinline__ void st_pred_u32(uint32_t* p, uint32_t v, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %2, 0;\n"
      "@q st.global.b32 [%0], %1;\n"
      "}\n" ::"l"(p),
      "r"(v), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void st_pred_v4(uint32_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q st.global.v4.b32 [%0], {%1, %2, %3, %4};\n"
      "}\n" ::"l"(p),
      "r"(a), "r"(b), "r"(c), "r"(d), "r"((uint32_t)pred)
      : "memory");
}

// Lane-contiguous row layout of the streaming kernels: lane l owns the 4*NV consecutive floats
// [l*4NV, (l+1)*4NV) of a row, so it belongs to exactly ONE head (lanes per head = 32 / H) and a per-head sum is a
// single butterfly over 32/H lanes (3 shuffles for 4 heads, against 9 shuffles + 14 selects when every lane held a
// slice of every head); exp / alpha / ge are evaluated once per lane instead of once per head.  The j-th float4 a
// lane holds is piece (j ^ swizzle(l)) of its span: with a 16*NV-byte lane stride, a quarter-warp of plain LDS.128
// would hit the same banks NV times; the xor makes the eight 16-byte pieces of every quarter-warp distinct.
template <int NV>
__device__ __forceinline__ int lc_off(int lane, int j) {
  constexpr int kShift = NV == 4 ? 1 : (NV == 2 ? 2 : 0);
  const int sw = NV == 1 ? 0 : (lane >> kShift) & (NV - 1);
  return lane * (4 * NV) + 4 * (j ^ sw);
}
// sum over the LC lanes of a head (LC compile-time, 0 = run-time lc_rt); every lane ends with its head's total
template <int LC>
__device__ __forceinline__ float head_sum(float p, int lc_rt) {
  if constexpr (LC == 0) {
    for (int off = lc_rt >> 1; off > 0; off >>= 1) p += shx(p, off);
  } else {
#pragma unroll
    for (int off = LC / 2; off > 0; off >>= 1) p += shx(p, off);
  }
  return p;
}
template <int LC>
__device__ __forceinline__ int lc_head(int lane, const Shape& sh) {
  if constexpr (LC == 0) return lane >> sh.lg_lc;
  else return lane / LC;
}
template <int LC>
__device__ __forceinline__ bool lc_head_lane(int lane, const Shape& sh) {
  if constexpr (LC == 0) return (lane & (sh.lc - 1)) == 0;
  else return (lane & (LC - 1)) == 0;
}

template <int NV>
__device__ __forceinline__ void load_row(float4 (&x)[NV], const float* __restrict__ base, int row, int lane) {
#pragma unroll
  for (int j = 0; j < NV; ++j) x[j] = ldg4(base + (int64_t)row * (NV * 128) + lc_off<NV>(lane, j));
}

// Partial-state slots: [chunk][2][PF] floats.  Slot 0: the segment that starts at the first edge of the chunk,
// slot 1: a segment that starts later and runs past the end of the chunk.
template <int NV>
struct FwdState {
  float m, s;  // one head per lane
  float4 acc[NV];
  __device__ __forceinline__ void init() {
    m = -1e9f;  // EB:336
    s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void merge(float m2, float s2, const float4 (&a2)[NV]) {
    const float mn = fmaxf(m, m2);
    const float c1 = __expf(m - mn), c2 = __expf(m2 - mn);
    s = s * c1 + s2 * c2;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      acc[j].x = acc[j].x * c1 + a2[j].x * c2;
      acc[j].y = acc[j].y * c1 + a2[j].y * c2;
      acc[j].z = acc[j].z * c1 + a2[j].z * c2;
      acc[j].w = acc[j].w * c1 + a2[j].w * c2;
    }
    m = mn;
  }
};
// partial state of a row segment: [F floats acc, natural element order][32 m][32 s] (per lane)
template <int NV>
constexpr int fwd_part_floats() { return NV * 128 + 2 * NV * 32; }

template <int NV>
__device__ __forceinline__ void fwd_store_partial(const FwdState<NV>& st, float* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < NV; ++j) st4(p + lc_off<NV>(lane, j), st.acc[j]);
  p[NV * 128 + lane] = st.m;
  p[NV * 128 + 32 + lane] = st.s;
}
template <int NV>
__device__ __forceinline__ void fwd_finalize(const FwdState<NV>& st, int row, const Shape sh, float* __restrict__ Hout,
                                             float* __restrict__ hpre, float* __restrict__ mx,
                                             float* __restrict__ sinv, int lane) {
  const float inv = 1.0f / (st.s + 1e-8f);  // EB:379
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 h = make_float4(st.acc[j].x * inv, st.acc[j].y * inv, st.acc[j].z * inv, st.acc[j].w * inv);
    if (sh.bias && (NV > 1 || lc_off<NV>(lane, j) < sh.F)) {
      const float4 b = ldg4(sh.bias + lc_off<NV>(lane, j));
      h = make_float4(h.x + b.x, h.y + b.y, h.z + b.z, h.w + b.w);
    }
    if (NV == 1 && lc_off<NV>(lane, j) >= sh.F) continue;  // 64-float rows (pair kernels): lanes 16-31 hold nothing
    const int64_t off = (int64_t)row * sh.F + lc_off<NV>(lane, j);
    if (hpre) st4(hpre + off, h);
    const float sc = sh.slopes.act;
    st4(Hout + off, make_float4(lrelu(h.x, sc), lrelu(h.y, sc), lrelu(h.z, sc), lrelu(h.w, sc)));  // EB:440-457
  }
  if ((lane & (sh.lc - 1)) == 0) {
    const int hd = lane >> sh.lg_lc;
    mx[(int64_t)row * sh.H + hd] = st.m;
    sinv[(int64_t)row * sh.H + hd] = inv;
  }
}

// ------------------------------------------------------------------------------------ forward
// ADROP: attention-coefficient dropout compiled in (a template parameter, not a run-time test of g.ascale: the predicated-off
// address arithmetic and load of the scale cost ~7 issue slots per edge in kernels that are bound by instruction issue)
template <int NV, int R, int LPH, bool ADROP>
__global__ void __launch_bounds__(kSW * 32)
edge_fwd_stream_kernel(StreamGraph g, const int* __restrict__ col_idx, const float* __restrict__ Pl,
                       const float* __restrict__ Pr, const float* __restrict__ a, Shape sh, float* __restrict__ Hout,
                       float* __restrict__ hpre, float* __restrict__ score, float* __restrict__ mx,
                       float* __restrict__ sinv, float* __restrict__ part) {
  constexpr int F = NV * 128;
  constexpr uint32_t kRowBytes = F * 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * R * F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kSW * R * F * 4) + warp * R;
  if (lane == 0) {
    for (int s = 0; s < R; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const GatherPolicy gp(g);
  const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bar);
  float4 av[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) av[j] = ldg4(a + lc_off<NV>(lane, j));
  const bool head_lane = lc_head_lane<LPH>(lane, sh);
  const int hd = lc_head<LPH>(lane, sh);
  int voff[NV];  // byte-free float offsets of this lane's pieces inside a ring slot
#pragma unroll
  for (int j = 0; j < NV; ++j) voff[j] = lc_off<NV>(lane, j);
  const float2 slope2 = splat2(g.slopes.attn);
  const int total_warps = gridDim.x * kSW;
  uint32_t it = 0;  // ring position: edges consumed by this warp so far
  for (int chunk = blockIdx.x * kSW + warp; chunk < g.n_chunks; chunk += total_warps) {
    const int e0 = chunk * g.T;
    const int e1 = min(g.E, e0 + g.T);
    const int n = e1 - e0;
    int r = __ldg(g.chunk_row + chunk);
    const bool started_before = __ldg(g.row_ptr + r) < e0;
    int row_end = __ldg(g.row_ptr + r + 1);
    int idx_cur = lane < n ? __ldg(col_idx + e0 + lane) : 0;
    int idx_nxt = 32 + lane < n ? __ldg(col_idx + e0 + 32 + lane) : 0;
    int idx_nn = 64 + lane < n ? __ldg(col_idx + e0 + 64 + lane) : 0;  // two windows ahead: its latency is never exposed
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int src = __shfl_sync(0xffffffffu, idx_cur, s);
      if (s < n) {
        const uint32_t slot = (it + s) % R;
        bulk_g2s_hint_elect(ring_s + (uint32_t)(slot * F) * 4u, Pl + gp.id(src) * F, kRowBytes, bar_s + slot * 8u, gp.of(src));
      }
    }
    float4 pr[NV], pr_n[NV];
    load_row<NV>(pr, Pr, r, lane);
    int next_end = 0x7fffffff;
    if (r + 1 < g.n_rows) {
      load_row<NV>(pr_n, Pr, r + 1, lane);
      next_end = __ldg(g.row_ptr + r + 2);
    }
    FwdState<NV> st;
    st.init();
    bool first_row = true;
    for (int i = 0; i < n; ++i) {
      const int e = e0 + i;
      while (e >= row_end) {  // warp-uniform: the current row ended inside this chunk
        if (first_row && started_before)
          fwd_store_partial<NV>(st, part + ((int64_t)chunk * 2 + 0) * fwd_part_floats<NV>(), lane);
        else if (e > e0 || !first_row)
          fwd_finalize<NV>(st, r, sh, Hout, hpre, mx, sinv, lane);
        first_row = false;
        ++r;
#pragma unroll
        for (int j = 0; j < NV; ++j) pr[j] = pr_n[j];
        row_end = next_end;
        next_end = 0x7fffffff;
        if (r + 1 < g.n_rows) {
          load_row<NV>(pr_n, Pr, r + 1, lane);
          next_end = __ldg(g.row_ptr + r + 2);
        }
        st.init();
      }
      if ((i & 31) == 0 && i > 0) {
        idx_cur = idx_nxt;
        idx_nxt = idx_nn;
        const int p = i + 64 + lane;
        idx_nn = p < n ? __ldg(col_idx + e0 + p) : 0;
      }
      const uint32_t pos = it + i, slot = pos % R, ph = (pos / R) & 1;
      mbar_wait_s(bar_s + slot * 8u, ph);
      float4 v[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = lds4_s(ring_s + (uint32_t)(slot * F + voff[j]) * 4u);
      __syncwarp();  // every lane has read the slot before it is refilled
      {
        const int ni = i + R;
        const int srcn = __shfl_sync(0xffffffffu, ((ni >> 5) == (i >> 5)) ? idx_cur : idx_nxt, ni & 31);
        if (ni < n) {
          RING_REFILL_FENCE();
          bulk_g2s_hint_elect(ring_s + (uint32_t)(slot * F) * 4u, Pl + gp.id(srcn) * F, kRowBytes, bar_s + slot * 8u, gp.of(srcn));
        }
      }
      float2 pp = make_float2(0.f, 0.f);  // even / odd elements: two interleaved FFMA2 chains
#pragma unroll
      for (int j = 0; j < NV; ++j) {  // EB:303-320
        pp = fma2(lo2(av[j]), lrelu_fast2(add2(lo2(v[j]), lo2(pr[j])), slope2), pp);
        pp = fma2(hi2(av[j]), lrelu_fast2(add2(hi2(v[j]), hi2(pr[j])), slope2), pp);
      }
      float p = head_sum<LPH>(pp.x + pp.y, sh.lc);
      st_pred_u32(reinterpret_cast<uint32_t*>(score + (int64_t)e * sh.H + hd), __float_as_uint(p), head_lane);
      // online form of EB:336-349.  Of exp(m - max) and exp(p - max) one is exp(0) = 1: a single exponential
      const float dlt = p - st.m;
      const bool up = dlt > 0.f;
      const float ex = __expf(-fabsf(dlt));
      const float corr = up ? ex : 1.f, w = up ? 1.f : ex;
      const float mn = up ? p : st.m;
      st.s = st.s * corr + w;
      // attention dropout scales the aggregated term only; the softmax denominator keeps every edge
      float wd = w;
      if constexpr (ADROP) wd = w * __ldg(g.ascale + (int64_t)e * sh.H + hd);
      const float2 corr2 = splat2(corr), w2 = splat2(wd);
#pragma unroll
      for (int j = 0; j < NV; ++j) {  // acc = acc * corr + (w * v), EB:415-422 without atomics (same rounding as the scalar form)
        const float2 a0 = fma2(lo2(st.acc[j]), corr2, mul2(w2, lo2(v[j])));
        const float2 a1 = fma2(hi2(st.acc[j]), corr2, mul2(w2, hi2(v[j])));
        st.acc[j] = make_float4(a0.x, a0.y, a1.x, a1.y);
      }
      st.m = mn;
    }
    // the segment that reaches the end of the chunk
    {
      const bool ended = row_end == e1;
      const bool from_start = first_row;  // segment began at e0
      const bool complete = ended && !(first_row && started_before);
      if (complete)
        fwd_finalize<NV>(st, r, sh, Hout, hpre, mx, sinv, lane);
      else
        fwd_store_partial<NV>(st, part + ((int64_t)chunk * 2 + (from_start ? 0 : 1)) * fwd_part_floats<NV>(), lane);
    }
    it += n;
  }
}

// One warp per chunk boundary: if a row straddles the boundary and this is the first boundary it crosses,
// merge all its pieces (chunk order) and finalise it.
template <int NV>
__global__ void __launch_bounds__(kSW * 32)
edge_fwd_fixup_kernel(StreamGraph g, Shape sh, const float* __restrict__ part, float* __restrict__ Hout,
                      float* __restrict__ hpre, float* __restrict__ mx, float* __restrict__ sinv) {
  const int c = blockIdx.x * kSW + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= g.n_chunks - 1) return;
  const int b = (c + 1) * g.T;
  const int rr = __ldg(g.chunk_row + c + 1);
  const int rs = __ldg(g.row_ptr + rr);
  if (rs >= b || rs < c * g.T) return;
  const int c_last = (__ldg(g.row_ptr + rr + 1) - 1) / g.T;
  constexpr int PF = fwd_part_floats<NV>();
  FwdState<NV> st;
  {
    const float* p = part + ((int64_t)c * 2 + (rs == c * g.T ? 0 : 1)) * PF;
#pragma unroll
    for (int j = 0; j < NV; ++j) st.acc[j] = lds4(p + lc_off<NV>(lane, j));
    st.m = p[NV * 128 + lane];
    st.s = p[NV * 128 + 32 + lane];
  }
  for (int cc = c + 1; cc <= c_last; ++cc) {
    const float* p = part + ((int64_t)cc * 2 + 0) * PF;
    float4 a2[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) a2[j] = lds4(p + lc_off<NV>(lane, j));
    st.merge(p[NV * 128 + lane], p[NV * 128 + 32 + lane], a2);
  }
  fwd_finalize<NV>(st, rr, sh, Hout, hpre, mx, sinv, lane);
}

// rows without any edge: h = 0, m = -1e9, 1/(s+eps) = 1e8
__global__ void fill_empty_fwd_kernel(const int* __restrict__ row_ptr, int n_rows, int F, int H,
                                      float* __restrict__ Hout, float* __restrict__ hpre, float* __restrict__ mx,
                                      float* __restrict__ sinv, const float* __restrict__ bias, float act_slope) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || row_ptr[r + 1] != row_ptr[r]) return;
  for (int k = 0; k < F; ++k) {
    const float h = bias ? bias[k] : 0.f;
    Hout[(int64_t)r * F + k] = lrelu(h, act_slope);
    if (hpre) hpre[(int64_t)r * F + k] = h;
  }
  for (int h = 0; h < H; ++h) {
    mx[(int64_t)r * H + h] = -1e9f;
    sinv[(int64_t)r * H + h] = 1e8f;
  }
}
// warp-cooperative: each lane tests one row, the warp then zeroes every empty row with coalesced 128-bit stores
// (on a multi-GPU shard many sources have no local out-edge, so this must run at memset speed)
__global__ void fill_empty_rows_kernel(const int* __restrict__ row_ptr, int n_rows, int F, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (base >= n_rows) return;
  const int r = base + lane;
  const bool empty = r < n_rows && __ldg(row_ptr + r + 1) == __ldg(row_ptr + r);
  uint32_t mask = __ballot_sync(0xffffffffu, empty);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  while (mask) {
    const int b = __ffs(mask) - 1;
    mask &= mask - 1;
    float* row = out + (int64_t)(base + b) * F;
    for (int k = 4 * lane; k < F; k += 128) st4(row + k, z);
  }
}

// ------------------------------------------------------------------------- backward, preparation
// Node-wise: g_h = gH * LReLU'(h) in place (EB:879-893 / EB:599; LReLU'(h) has the sign of LReLU(h)) and
// cdot[row][h] = gH . Hout = sum over the row's edges of alpha * galpha (the softmax-backward segment sum).
template <int NV>
__global__ void __launch_bounds__(256)
edge_bwd_prep_kernel(int n_rows, Shape sh, const float* __restrict__ Hout, float* __restrict__ gH,
                     float* __restrict__ cdot) {
  const int lane = threadIdx.x & 31;
  const bool head_lane = (lane & (sh.lph - 1)) == 0;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 8) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const bool in_row = NV > 1 || 4 * lane < sh.F;  // 64-float rows: lanes 16-31 only take part in the shuffles
      const int64_t off = (int64_t)row * sh.F + (in_row ? 4 * (lane + 32 * j) : 0);
      const float4 g = in_row ? *reinterpret_cast<const float4*>(gH + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 ho = ldg4(Hout + off);
      const float sc = sh.slopes.act;
      const float4 gp = make_float4(g.x * lrelu_grad(ho.x, sc), g.y * lrelu_grad(ho.y, sc), g.z * lrelu_grad(ho.z, sc),
                                    g.w * lrelu_grad(ho.w, sc));
      // sum_seg alpha*galpha = g_pre . (h - bias) = gH . Hout - g_pre . bias  (LReLU'(h) h = LReLU(h))
      float cd = dot4(g, ho);
      if (sh.bias && in_row) cd -= dot4(gp, ldg4(sh.bias + 4 * (lane + 32 * j)));
      const float c = head_reduce(cd, sh.lph);
      if (in_row) st4(gH + off, gp);
      if (head_lane && in_row) cdot[(int64_t)row * sh.H + ((lane + 32 * j) >> sh.lg_lph)] = c;
    }
  }
}

// ------------------------------------------------------------------------- backward, pass 1
__host__ __device__ inline int rec_words(int H, int NV) { return (4 * NV + 2 * H + 3) / 4 * 4; }

struct RowScalars {
  float c, m, inv;  // of this lane's head
};
__device__ __forceinline__ void load_scalars(RowScalars& q, int row, int H, int hd, const float* __restrict__ cdot,
                                             const float* __restrict__ mx, const float* __restrict__ sinv) {
  const int64_t o = (int64_t)row * H + hd;
  q.c = __ldg(cdot + o);
  q.m = __ldg(mx + o);
  q.inv = __ldg(sinv + o);
}

// smem per warp: ring [R][F] | rowbuf [2F] (g_h row, P_r row) | score window [2][32*H] | barriers [R + 1]
template <int NV, int R, int LPH, bool ADROP>
__global__ void __launch_bounds__(kSW * 32)
edge_bwd_dst_stream_kernel(StreamGraph g, const int* __restrict__ col_idx, const float* __restrict__ Pl,
                           const float* __restrict__ Pr, const float* __restrict__ a, Shape sh,
                           const float* __restrict__ gh, const float* __restrict__ cdot,
                           const float* __restrict__ score, const float* __restrict__ mx,
                           const float* __restrict__ sinv, float* __restrict__ gPr, uint32_t* __restrict__ rec,
                           float* __restrict__ part, float* __restrict__ ga_partials, float* __restrict__ galpha_dbg,
                           int max_h) {
  constexpr int F = NV * 128;
  constexpr uint32_t kRowBytes = F * 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp_floats = R * F + 2 * F + 2 * 32 * max_h;
  float* wbase = reinterpret_cast<float*>(smem_raw) + (size_t)warp * per_warp_floats;
  float* ring = wbase;
  float* rowbuf = wbase + R * F;        // [2F]
  float* scwin = rowbuf + 2 * F;        // [2][32*H]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kSW * per_warp_floats * 4) + warp * (R + 1);
  uint64_t* rbar = bar + R;
  if (lane == 0) {
    for (int s = 0; s < R + 1; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const GatherPolicy gp(g);
  const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bar), rowbuf_s = smem_u32(rowbuf), rbar_s = smem_u32(rbar);
  float4 ga[NV];  // the attention vector a is only needed when a row segment is written: read it there (L1 hit)
#pragma unroll
  for (int j = 0; j < NV; ++j) ga[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int H = LPH > 0 ? 32 / (LPH > 0 ? LPH : 1) : sh.H;  // compile-time when the head width is
  const int RW = rec_words(H, NV);
  const bool head_lane = lc_head_lane<LPH>(lane, sh);
  const int hd = lc_head<LPH>(lane, sh);
  int voff[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) voff[j] = lc_off<NV>(lane, j);
  const int total_warps = gridDim.x * kSW;
  uint32_t it = 0;   // ring position
  uint32_t rk = 0;   // row-buffer position: row loads issued so far
  for (int chunk = blockIdx.x * kSW + warp; chunk < g.n_chunks; chunk += total_warps) {
    const int e0 = chunk * g.T;
    const int e1 = min(g.E, e0 + g.T);
    const int n = e1 - e0;
    int r = __ldg(g.chunk_row + chunk);
    const bool started_before = __ldg(g.row_ptr + r) < e0;
    int row_end = __ldg(g.row_ptr + r + 1);
    int idx_cur = lane < n ? __ldg(col_idx + e0 + lane) : 0;
    int idx_nxt = 32 + lane < n ? __ldg(col_idx + e0 + 32 + lane) : 0;
    int idx_nn = 64 + lane < n ? __ldg(col_idx + e0 + 64 + lane) : 0;  // two windows ahead: its latency is never exposed
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int src = __shfl_sync(0xffffffffu, idx_cur, s);
      if (s < n) {
        const uint32_t slot = (it + s) % R;
        bulk_g2s_hint_elect(ring_s + (uint32_t)(slot * F) * 4u, Pl + gp.id(src) * F, kRowBytes, bar_s + slot * 8u, gp.of(src));
      }
    }
    // row data (g_h row, P_r row) of r goes through the row buffer; once it sits in registers the buffer is free,
    // so row r + 1 is prefetched into the same buffer while row r is being processed
    // generic loads have read the row buffer (previous chunk / row): order them before the async-proxy overwrite
    fence_proxy_async_smem();
    bulk2_g2s_hint_elect(rowbuf_s, gh + (int64_t)r * F, kRowBytes, gp.cold, rowbuf_s + kRowBytes, Pr + (int64_t)r * F,
                         kRowBytes, gp.cold, rbar_s);
    // score window: the scores of 32 consecutive edges are contiguous ([E][H]); the next window is
    // prefetched into registers while the current one is consumed from shared memory
    float scp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < H) {
        const int t = lane + 32 * k;
        scwin[t] = (t < n * H) ? __ldg(score + (int64_t)e0 * H + t) : 0.f;
      }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = 32 * H + lane + 32 * k;
      scp[k] = (k < H && t < n * H) ? __ldg(score + (int64_t)e0 * H + t) : 0.f;
    }
    RowScalars q, qn;
    load_scalars(q, r, H, hd, cdot, mx, sinv);
    qn = q;
    int next_end = 0x7fffffff;
    if (r + 1 < g.n_rows) {
      load_scalars(qn, r + 1, H, hd, cdot, mx, sinv);
      next_end = __ldg(g.row_ptr + r + 2);
    }
    float4 ghr[NV], pr[NV], gpr[NV];
    mbar_wait(rbar, rk & 1);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      ghr[j] = lds4(rowbuf + voff[j]);
      pr[j] = lds4(rowbuf + F + voff[j]);
      gpr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    ++rk;
    if (r + 1 < g.n_rows) {
      fence_proxy_async_smem();
      bulk2_g2s_hint_elect(rowbuf_s, gh + (int64_t)(r + 1) * F, kRowBytes, gp.cold, rowbuf_s + kRowBytes,
                           Pr + (int64_t)(r + 1) * F, kRowBytes, gp.cold, rbar_s);
    }
    bool first_row = true;
    for (int i = 0; i < n; ++i) {
      const int e = e0 + i;
      while (e >= row_end) {
        // finish row r: gP_r[r] (EB:781 summed over the row) complete or partial
        float* dst = (first_row && started_before) ? part + ((int64_t)chunk * 2 + 0) * F : gPr + (int64_t)r * F;
        if (e > e0 || !first_row) {
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const float4 avj = ldg4(a + voff[j]);
            st4(dst + voff[j], make_float4(gpr[j].x * avj.x, gpr[j].y * avj.y, gpr[j].z * avj.z,
                                                       gpr[j].w * avj.w));
          }
        }
        first_row = false;
        ++r;
        q = qn;
        row_end = next_end;
        next_end = 0x7fffffff;
        // the prefetched row r has landed (or is landing) in the row buffer; take it and prefetch r + 1
        mbar_wait(rbar, rk & 1);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          ghr[j] = lds4(rowbuf + voff[j]);
          pr[j] = lds4(rowbuf + F + voff[j]);
          gpr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        ++rk;
        if (r + 1 < g.n_rows) {
          fence_proxy_async_smem();
          bulk2_g2s_hint_elect(rowbuf_s, gh + (int64_t)(r + 1) * F, kRowBytes, gp.cold, rowbuf_s + kRowBytes,
                               Pr + (int64_t)(r + 1) * F, kRowBytes, gp.cold, rbar_s);
          load_scalars(qn, r + 1, H, hd, cdot, mx, sinv);
          next_end = __ldg(g.row_ptr + r + 2);
        }
      }
      if ((i & 31) == 0 && i > 0) {
        idx_cur = idx_nxt;
        idx_nxt = idx_nn;
        const int p = i + 64 + lane;
        idx_nn = p < n ? __ldg(col_idx + e0 + p) : 0;
        float* w = scwin + ((i >> 5) & 1) * 32 * H;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < H) w[lane + 32 * k] = scp[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int t = (i + 32) * H + lane + 32 * k;
          scp[k] = (k < H && t < n * H) ? __ldg(score + (int64_t)e0 * H + t) : 0.f;
        }
      }
      const uint32_t pos = it + i, slot = pos % R, ph = (pos / R) & 1;
      mbar_wait_s(bar_s + slot * 8u, ph);
      float4 v[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = lds4_s(ring_s + (uint32_t)(slot * F + voff[j]) * 4u);
      __syncwarp();
      {
        const int ni = i + R;
        const int srcn = __shfl_sync(0xffffffffu, ((ni >> 5) == (i >> 5)) ? idx_cur : idx_nxt, ni & 31);
        if (ni < n) {
          RING_REFILL_FENCE();
          bulk_g2s_hint_elect(ring_s + (uint32_t)(slot * F) * 4u, Pl + gp.id(srcn) * F, kRowBytes, bar_s + slot * 8u, gp.of(srcn));
        }
      }
      const float* sc = scwin + ((i >> 5) & 1) * 32 * H + (i & 31) * H;
      uint32_t* re = rec + (int64_t)e * RW;
      float2 gal2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NV; ++j) {  // EB:636-646, two interleaved FFMA2 chains (even / odd elements)
        gal2 = fma2_s(lo2(ghr[j]), lo2(v[j]), gal2);
        gal2 = fma2_s(hi2(ghr[j]), hi2(v[j]), gal2);
      }
      float galpha = head_sum<LPH>(gal2.x + gal2.y, sh.lc);
      float alpha = __expf(sc[hd] - q.m) * q.inv;        // EB:378-379
      float dsc = 1.f;  // attention dropout (1 when off)
      if constexpr (ADROP) dsc = __ldg(g.ascale + (int64_t)e * H + hd);
      galpha *= dsc;                                     // h = sum alpha * dsc * P_l
      const float ge = alpha * (galpha - q.c);           // EB:689-690 in closed form
      alpha *= dsc;                                      // the record keeps alpha * dsc for pass 2
      const float ges = ge * g.slopes.attn;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float2 s01 = add2_s(lo2(v[j]), lo2(pr[j])), s23 = add2_s(hi2(v[j]), hi2(pr[j]));
        const float sx = s01.x, sy = s01.y, sz = s23.x, sw = s23.y;
        const bool px = sx > 0.f, py = sy > 0.f, pz = sz > 0.f, pw = sw > 0.f;
        // u = ge * LReLU'(s);  ga += u * s = ge * LReLU(s) (EB:769);  gP_r += a * u (EB:774-781, a applied per row)
        const float2 u01 = make_float2(px ? ge : ges, py ? ge : ges), u23 = make_float2(pz ? ge : ges, pw ? ge : ges);
        const float2 g01 = fma2_s(u01, s01, lo2(ga[j])), g23 = fma2_s(u23, s23, hi2(ga[j]));
        ga[j] = make_float4(g01.x, g01.y, g23.x, g23.y);
        const float2 r01 = add2_s(lo2(gpr[j]), u01), r23 = add2_s(hi2(gpr[j]), u23);
        gpr[j] = make_float4(r01.x, r01.y, r23.x, r23.y);
        const uint32_t bx = __ballot_sync(0xffffffffu, px), by = __ballot_sync(0xffffffffu, py),
                       bz = __ballot_sync(0xffffffffu, pz), bw = __ballot_sync(0xffffffffu, pw);
        // predicated stores (no divergence regions): lane 0 writes the sign words (bit = lane, word = 4 j + component)
        st_pred_v4(re + 4 * j, bx, by, bz, bw, lane == 0);
      }
      st_pred_u32(re + 4 * NV + hd, __float_as_uint(alpha), head_lane);
      st_pred_u32(re + 4 * NV + H + hd, __float_as_uint(ge), head_lane);
      if (galpha_dbg && head_lane) galpha_dbg[(int64_t)e * H + hd] = galpha;
    }
    {
      const bool ended = row_end == e1;
      const bool complete = ended && !(first_row && started_before);
      float* dst = complete ? gPr + (int64_t)r * F : part + ((int64_t)chunk * 2 + (first_row ? 0 : 1)) * F;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 avj = ldg4(a + voff[j]);
        st4(dst + voff[j], make_float4(gpr[j].x * avj.x, gpr[j].y * avj.y, gpr[j].z * avj.z,
                                                   gpr[j].w * avj.w));
      }
    }
    it += n;
    // the row-buffer pipeline restarts at the next chunk: drain the outstanding prefetch of row r + 1
    if (r + 1 < g.n_rows) {
      mbar_wait(rbar, rk & 1);
      ++rk;
    }
    __syncwarp();
  }
  // deterministic block sum of the per-warp ga accumulators
  __syncthreads();
  float4* sm = reinterpret_cast<float4*>(smem_raw);
#pragma unroll
  for (int j = 0; j < NV; ++j) sm[warp * (NV * 32) + lane + 32 * j] = ga[j];  // staging order is private
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 t = sm[lane + 32 * j];
      for (int w = 1; w < kSW; ++w) {
        const float4 o = sm[w * (NV * 32) + lane + 32 * j];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      st4(ga_partials + (int64_t)blockIdx.x * F + voff[j], t);
    }
  }
}

// sums the pieces of a row that straddles chunk boundaries (pass 1: gP_r, pass 2: gP_l)
template <int NV, int F = NV * 128>
__global__ void __launch_bounds__(kSW * 32)
edge_sum_fixup_kernel(StreamGraph g, const float* __restrict__ part, float* __restrict__ out) {
  static_assert(F == NV * 128 || (NV == 1 && F == 64), "row width");
  const int c = blockIdx.x * kSW + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= g.n_chunks - 1) return;
  const int b = (c + 1) * g.T;
  const int rr = __ldg(g.chunk_row + c + 1);
  const int rs = __ldg(g.row_ptr + rr);
  if (rs >= b || rs < c * g.T) return;
  const int c_last = (__ldg(g.row_ptr + rr + 1) - 1) / g.T;
  if (4 * lane >= F) return;
  float4 acc[NV];
  {
    const float* p = part + ((int64_t)c * 2 + (rs == c * g.T ? 0 : 1)) * F;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = lds4(p + 4 * (lane + 32 * j));
  }
  for (int cc = c + 1; cc <= c_last; ++cc) {
    const float* p = part + ((int64_t)cc * 2 + 0) * F;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 o = lds4(p + 4 * (lane + 32 * j));
      acc[j].x += o.x; acc[j].y += o.y; acc[j].z += o.z; acc[j].w += o.w;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) st4(out + (int64_t)rr * F + 4 * (lane + 32 * j), acc[j]);
}

// ------------------------------------------------------------------------- backward, pass 2
// ring slot: [F floats of g_h[dst]] [RW words of the edge record], padded to 32 floats
template <int NV, int R>
__global__ void __launch_bounds__(kSW * 32)
edge_bwd_src_stream_kernel(StreamGraph g, const int* __restrict__ csc_dst, const int* __restrict__ csc_eid,
                           const float* __restrict__ a, Shape sh, const float* __restrict__ gh,
                           const uint32_t* __restrict__ rec, float* __restrict__ gPl, float* __restrict__ part,
                           int slot_floats) {
  constexpr int F = NV * 128;
  constexpr uint32_t kRowBytes = F * 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * R * slot_floats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kSW * R * slot_floats * 4) + warp * R;
  if (lane == 0) {
    for (int s = 0; s < R; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const GatherPolicy gp(g);
  const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bar);
  float4 av[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) av[j] = ldg4(a + lc_off<NV>(lane, j));
  int voff[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) voff[j] = lc_off<NV>(lane, j);
  const int hd = lane >> sh.lg_lc;
  const int RW = rec_words(sh.H, NV);
  const uint32_t kRecBytes = RW * 4;
  const int total_warps = gridDim.x * kSW;
  uint32_t it = 0;
  for (int chunk = blockIdx.x * kSW + warp; chunk < g.n_chunks; chunk += total_warps) {
    const int e0 = chunk * g.T;
    const int e1 = min(g.E, e0 + g.T);
    const int n = e1 - e0;
    int r = __ldg(g.chunk_row + chunk);
    const bool started_before = __ldg(g.row_ptr + r) < e0;
    int row_end = __ldg(g.row_ptr + r + 1);
    int next_end = r + 1 < g.n_rows ? __ldg(g.row_ptr + r + 2) : 0x7fffffff;
    int dst_cur = lane < n ? __ldg(csc_dst + e0 + lane) : 0, eid_cur = lane < n ? __ldg(csc_eid + e0 + lane) : 0;
    int dst_nxt = 32 + lane < n ? __ldg(csc_dst + e0 + 32 + lane) : 0;
    int eid_nxt = 32 + lane < n ? __ldg(csc_eid + e0 + 32 + lane) : 0;
    int dst_nn = 64 + lane < n ? __ldg(csc_dst + e0 + 64 + lane) : 0;  // two windows ahead
    int eid_nn = 64 + lane < n ? __ldg(csc_eid + e0 + 64 + lane) : 0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int d = __shfl_sync(0xffffffffu, dst_cur, s), ee = __shfl_sync(0xffffffffu, eid_cur, s);
      if (s < n) {
        const uint32_t slot = (it + s) % R;
        bulk2_g2s_hint_elect(ring_s + (uint32_t)(slot * slot_floats) * 4u, gh + gp.id(d) * F, kRowBytes, gp.of(d),
                             ring_s + (uint32_t)(slot * slot_floats + F) * 4u, rec + (int64_t)ee * RW, kRecBytes, gp.cold,
                             bar_s + slot * 8u);
      }
    }
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool first_row = true;
    for (int i = 0; i < n; ++i) {
      const int e = e0 + i;
      while (e >= row_end) {
        float* dstp = (first_row && started_before) ? part + ((int64_t)chunk * 2 + 0) * F : gPl + (int64_t)r * F;
        if (e > e0 || !first_row) {
#pragma unroll
          for (int j = 0; j < NV; ++j) st4(dstp + voff[j], acc[j]);
        }
        first_row = false;
        ++r;
        row_end = next_end;
        next_end = r + 1 < g.n_rows ? __ldg(g.row_ptr + r + 2) : 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if ((i & 31) == 0 && i > 0) {
        dst_cur = dst_nxt;
        eid_cur = eid_nxt;
        dst_nxt = dst_nn;
        eid_nxt = eid_nn;
        const int p = i + 64 + lane;
        dst_nn = p < n ? __ldg(csc_dst + e0 + p) : 0;
        eid_nn = p < n ? __ldg(csc_eid + e0 + p) : 0;
      }
      const uint32_t pos = it + i, slot = pos % R, ph = (pos / R) & 1;
      mbar_wait_s(bar_s + slot * 8u, ph);
      const uint32_t sl = ring_s + (uint32_t)(slot * slot_floats) * 4u, rw = sl + F * 4u;
      float4 gv[NV];
      uint4 kk[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        gv[j] = lds4_s(sl + (uint32_t)voff[j] * 4u);
        kk[j] = lds4u_s(rw + 16u * j);
      }
      const float al = __uint_as_float(lds1u_s(rw + (uint32_t)(4 * NV + hd) * 4u));  // this lane's head
      const float ge = __uint_as_float(lds1u_s(rw + (uint32_t)(4 * NV + sh.H + hd) * 4u));
      const float ges = ge * g.slopes.attn;
      const float2 al2 = splat2(al);
      __syncwarp();
      {
        const int ni = i + R;
        const bool same = (ni >> 5) == (i >> 5);
        const int d = __shfl_sync(0xffffffffu, same ? dst_cur : dst_nxt, ni & 31);
        const int ee = __shfl_sync(0xffffffffu, same ? eid_cur : eid_nxt, ni & 31);
        if (ni < n) {
          RING_REFILL_FENCE();
          bulk2_g2s_hint_elect(ring_s + (uint32_t)(slot * slot_floats) * 4u, gh + gp.id(d) * F, kRowBytes, gp.of(d),
                               ring_s + (uint32_t)(slot * slot_floats + F) * 4u, rec + (int64_t)ee * RW, kRecBytes, gp.cold,
                               bar_s + slot * 8u);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        // EB:865-866: g_h[dst] * alpha + ge * a * LReLU'(s), LReLU'(s) from the recorded sign bit
        const float2 u01 = make_float2(((kk[j].x >> lane) & 1u) ? ge : ges, ((kk[j].y >> lane) & 1u) ? ge : ges);
        const float2 u23 = make_float2(((kk[j].z >> lane) & 1u) ? ge : ges, ((kk[j].w >> lane) & 1u) ? ge : ges);
        const float2 a01 = fma2_s(u01, lo2(av[j]), fma2_s(al2, lo2(gv[j]), lo2(acc[j])));
        const float2 a23 = fma2_s(u23, hi2(av[j]), fma2_s(al2, hi2(gv[j]), hi2(acc[j])));
        acc[j] = make_float4(a01.x, a01.y, a23.x, a23.y);
      }
    }
    {
      const bool ended = row_end == e1;
      const bool complete = ended && !(first_row && started_before);
      float* dstp = complete ? gPl + (int64_t)r * F : part + ((int64_t)chunk * 2 + (first_row ? 0 : 1)) * F;
#pragma unroll
      for (int j = 0; j < NV; ++j) st4(dstp + voff[j], acc[j]);
    }
    it += n;
  }
}

#include "edge_stream_pair.inc"

bool make_stream_shape(int H, int D, Shape* sh, int* nv) {
  if (H < 1 || D < 4 || D % 4 || D > 128 || H > 8) return false;
  const int lph = D / 4;
  if (lph & (lph - 1)) return false;
  const int F = H * D;
  const bool narrow = H == 1 && D == 64;  // one head of 64 floats: the PF = 64 pair kernels, half a warp per row
  if (F % 128 && !narrow) return false;
  const int NV = narrow ? 1 : F / 128;
  if (NV != 1 && NV != 2 && NV != 4) return false;
  int lg = 0;
  while ((1 << lg) < lph) ++lg;
  if (32 % H) return false;  // lane-contiguous layout: 32 / H lanes per head
  const int lc = 32 / H;
  int lgc = 0;
  while ((1 << lgc) < lc) ++lgc;
  *sh = Shape{H, D, F, lph, lg, lc, lgc, Slopes{kSlope, kSlope}, nullptr};
  *nv = NV;
  return true;
}

// `reserve`: CTA slots left free for the exchange kernels that run underneath this launch (multi-GPU pipeline).  The
// chunks are assigned statically (grid-stride), so every CTA of the grid must be resident at once: a CTA that had to
// wait for a slot taken by another stream's kernel would run its whole share after the others finished.
int stream_grid(const void* kernel, size_t smem, int n_chunks, int reserve) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSW * 32, smem);
  if (per_sm < 1) per_sm = 1;
  int blocks = kNumSMs * per_sm - reserve;
  if (blocks < kNumSMs) blocks = kNumSMs;
  const int need = (n_chunks + kSW - 1) / kSW;
  return blocks < need ? blocks : need;
}

template <int NV>
constexpr int ring_depth() { return NV == 4 ? 8 : (NV == 2 ? 8 : 16); }

// which hot-set bit of the hinted index arrays a layer with F-float rows uses (wide rows: the smaller hot set)
struct HotSel {
  bool on;
  uint32_t bit, mask;
};
HotSel hot_select(const EdgeGraph& eg, int F) {
  if (!eg.col_idx_hot || !eg.csc_dst_hot) return HotSel{false, 0u, 0xffffffffu};
  return HotSel{true, 2 * F > eg.hot_wide_F ? 0x80000000u : 0x40000000u, 0x3fffffffu};
}

bool use_pair(int nv, const Shape& sh) {
  static const bool off = getenv("GATX_NO_PAIR") != nullptr;
  if (sh.F == 64) return true;  // no other kernel family in this file handles 64-float rows
  return !off && nv == 1 && sh.H == 1 && sh.lph == 32;
}
#define PAIR_DISPATCH_PF(F_, ...)                        \
  do {                                                   \
    if ((F_) == 64) { constexpr int PF = 64; __VA_ARGS__; } \
    else { constexpr int PF = 128; __VA_ARGS__; }        \
  } while (0)

#define STREAM_DISPATCH_NV(nv, ...)                          \
  do {                                                       \
    if (nv == 4) { constexpr int NV = 4; __VA_ARGS__; }      \
    else if (nv == 2) { constexpr int NV = 2; __VA_ARGS__; } \
    else { constexpr int NV = 1; __VA_ARGS__; }              \
  } while (0)
// lanes per head (32 / H): 8 (4 heads) and 16 (2 heads) are compile-time fast paths, anything else is run-time
#define STREAM_DISPATCH(nv, lc, ...)                                           \
  do {                                                                          \
    if (lc == 8) { constexpr int LPH = 8; STREAM_DISPATCH_NV(nv, __VA_ARGS__); }        \
    else if (lc == 16) { constexpr int LPH = 16; STREAM_DISPATCH_NV(nv, __VA_ARGS__); } \
    else { constexpr int LPH = 0; STREAM_DISPATCH_NV(nv, __VA_ARGS__); }                \
  } while (0)

}  // namespace

bool edge_stream_supported(int H, int D) {
  Shape sh;
  int nv;
  return make_stream_shape(H, D, &sh, &nv);
}
int64_t edge_stream_part_floats(int H, int D, int n_chunks) {
  Shape sh;
  int nv;
  if (!make_stream_shape(H, D, &sh, &nv)) return 0;
  return (int64_t)n_chunks * 2 * (nv * 128 + 2 * nv * 32);  // fwd_part_floats<NV>(), the largest user
}

int launch_edge_forward_stream(const EdgeGraph& eg, int H, int D, const float* Pl, const float* Pr, const float* a,
                               float* Hout, float* hpre, float* score, float* mx, float* sinv, float* part,
                               cudaStream_t st) {
  Shape sh;
  int nv, launches = 0;
  if (!make_stream_shape(H, D, &sh, &nv) || eg.E >= 0x7fffffffLL) return -1;
  sh.slopes = eg.slopes;
  sh.bias = eg.bias;
  if (eg.n_rows <= 0) return 0;
  fill_empty_fwd_kernel<<<(eg.n_rows + 255) / 256, 256, 0, st>>>(eg.row_ptr, eg.n_rows, sh.F, H, Hout, hpre, mx, sinv, eg.bias,
                                                                       eg.slopes.act);
  ++launches;
  if (eg.E == 0) return launches;
  const HotSel hs = hot_select(eg, sh.F);
  const int* colx = hs.on ? eg.col_idx_hot : eg.col_idx;
  StreamGraph g{(int)eg.E, eg.chunk_T, eg.n_chunks, eg.n_rows, eg.row_ptr, eg.chunk_row, hs.bit, hs.mask, eg.slopes, eg.bias, eg.ascale};
  if (use_pair(nv, sh)) {  // one head of 128 / 64 floats: two edges per loop iteration
    PAIR_DISPATCH_PF(sh.F, {
      constexpr int R = 16;
      const size_t smem = (size_t)kSW * R * PF * 4 + (size_t)kSW * R * 8;
      auto kern = eg.ascale ? edge_fwd_pair_kernel<R, true, PF> : edge_fwd_pair_kernel<R, false, PF>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const int blocks = stream_grid((const void*)kern, smem, g.n_chunks, eg.reserve_ctas);
      if (eg.kernel_events) cudaEventRecord(eg.kernel_events[0], st);
      // 512-byte rows: the cache-hint form of the bulk copy is slower than the plain one at this size (measured), no hints
      kern<<<blocks, kSW * 32, smem, st>>>(g, eg.col_idx, Pl, Pr, a, Hout, hpre, score, mx, sinv, part);
      if (eg.kernel_events) cudaEventRecord(eg.kernel_events[1], st);
    });
    ++launches;
    if (g.n_chunks > 1) {
      edge_fwd_fixup_kernel<1><<<(g.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(g, sh, part, Hout, hpre, mx, sinv);
      ++launches;
    }
    return launches;
  }
  STREAM_DISPATCH(nv, sh.lc, {
    constexpr int R = ring_depth<NV>();
    const size_t smem = (size_t)kSW * R * NV * 128 * 4 + (size_t)kSW * R * 8;
    auto kern = eg.ascale ? edge_fwd_stream_kernel<NV, R, LPH, true> : edge_fwd_stream_kernel<NV, R, LPH, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = stream_grid((const void*)kern, smem, g.n_chunks, eg.reserve_ctas);
    if (eg.kernel_events) cudaEventRecord(eg.kernel_events[0], st);
    kern<<<blocks, kSW * 32, smem, st>>>(g, colx, Pl, Pr, a, sh, Hout, hpre, score, mx, sinv, part);
    if (eg.kernel_events) cudaEventRecord(eg.kernel_events[1], st);
    ++launches;
    if (g.n_chunks > 1) {
      edge_fwd_fixup_kernel<NV><<<(g.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(g, sh, part, Hout, hpre, mx, sinv);
      ++launches;
    }
  });
  return launches;
}

int launch_edge_backward_stream(const EdgeGraph& eg, int H, int D, const float* Pl, const float* Pr, const float* a,
                                const float* Hout, float* gH, float* cdot, const float* score, const float* mx,
                                const float* sinv, float* gPr, float* gPl, uint32_t* rec, float* part,
                                float* ga_partials, int* n_partials, float* galpha_dbg, cudaStream_t st, int phases) {
  Shape sh;
  int nv, launches = 0;
  if (!make_stream_shape(H, D, &sh, &nv) || eg.E >= 0x7fffffffLL) return -1;
  sh.slopes = eg.slopes;
  sh.bias = eg.bias;
  const bool do_p1 = (phases & 1) != 0, do_p2 = (phases & 2) != 0;
  const bool prepared = (phases & 4) != 0;  // the input-gradient GEMM's epilogue already wrote g_pre and cdot
  if (do_p1) *n_partials = 0;
  if (eg.n_rows <= 0 && !do_p2) return 0;
  if (eg.n_rows <= 0 && eg.n_src <= 0) return 0;
  const HotSel hs = hot_select(eg, sh.F);
  const int* colx = hs.on ? eg.col_idx_hot : eg.col_idx;
  const int* cdstx = hs.on ? eg.csc_dst_hot : eg.csc_dst;
  StreamGraph gd{(int)eg.E, eg.chunk_T, eg.n_chunks, eg.n_rows, eg.row_ptr, eg.chunk_row, hs.bit, hs.mask, eg.slopes, eg.bias, eg.ascale};
  StreamGraph gs{(int)eg.E, eg.chunk_T, eg.n_chunks, eg.n_src, eg.csc_ptr, eg.chunk_src, hs.bit, hs.mask, eg.slopes, eg.bias, eg.ascale};
  if (use_pair(nv, sh)) {
    constexpr int R = 16;
    const int F = sh.F;
    if (do_p1 && eg.n_rows > 0) {
      int blocks = (eg.n_rows + 7) / 8;
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      if (!prepared) {
        edge_bwd_prep_kernel<1><<<blocks, 256, 0, st>>>(eg.n_rows, sh, Hout, gH, cdot);
        ++launches;
      }
      fill_empty_rows_kernel<<<(eg.n_rows + 255) / 256, 256, 0, st>>>(eg.row_ptr, eg.n_rows, F, gPr);
      ++launches;
    }
    if (do_p2) {
      fill_empty_rows_kernel<<<(eg.n_src + 255) / 256, 256, 0, st>>>(eg.csc_ptr, eg.n_src, F, gPl);
      ++launches;
    }
    if (eg.E > 0) {
      if (do_p1) {
        PAIR_DISPATCH_PF(F, {
          const size_t per_warp = (size_t)(R * PF + 2 * PF + 64) * 4;
          const size_t smem = kSW * per_warp + (size_t)kSW * (R + 1) * 8;
          auto kern = eg.ascale ? edge_bwd_dst_pair_kernel<R, true, PF> : edge_bwd_dst_pair_kernel<R, false, PF>;
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          const int blocks = stream_grid((const void*)kern, smem, gd.n_chunks, eg.reserve_ctas);
          if (eg.kernel_events) cudaEventRecord(eg.kernel_events[2], st);
          kern<<<blocks, kSW * 32, smem, st>>>(gd, eg.col_idx, Pl, Pr, a, gH, cdot, score, mx, sinv, gPr, rec, part,
                                               ga_partials, galpha_dbg);
          if (eg.kernel_events) cudaEventRecord(eg.kernel_events[3], st);
          *n_partials = blocks;
          ++launches;
          if (gd.n_chunks > 1) {
            edge_sum_fixup_kernel<1, PF><<<(gd.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(gd, part, gPr);
            ++launches;
          }
        });
      }
      if (do_p2) {
        PAIR_DISPATCH_PF(F, {
          const size_t smem = (size_t)kSW * R * (PF + 32) * 4 + (size_t)kSW * R * 8;
          auto kern = edge_bwd_src_pair_kernel<R, PF>;
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          const int blocks = stream_grid((const void*)kern, smem, gs.n_chunks, eg.reserve_ctas);
          if (eg.kernel_events) cudaEventRecord(eg.kernel_events[4], st);
          kern<<<blocks, kSW * 32, smem, st>>>(gs, eg.csc_dst, eg.csc_eid, a, gH, rec, gPl, part);
          if (eg.kernel_events) cudaEventRecord(eg.kernel_events[5], st);
          ++launches;
          if (gs.n_chunks > 1) {
            edge_sum_fixup_kernel<1, PF><<<(gs.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(gs, part, gPl);
            ++launches;
          }
        });
      }
    }
    return launches;
  }
  STREAM_DISPATCH(nv, sh.lc, {
    constexpr int F = NV * 128;
    if (do_p1 && eg.n_rows > 0) {
      int blocks = (eg.n_rows + 7) / 8;
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      if (!prepared) {
        edge_bwd_prep_kernel<NV><<<blocks, 256, 0, st>>>(eg.n_rows, sh, Hout, gH, cdot);
        ++launches;
      }
      fill_empty_rows_kernel<<<(eg.n_rows + 255) / 256, 256, 0, st>>>(eg.row_ptr, eg.n_rows, F, gPr);  // 8 warps x 32 rows
      ++launches;
    }
    if (do_p2) {
      fill_empty_rows_kernel<<<(eg.n_src + 255) / 256, 256, 0, st>>>(eg.csc_ptr, eg.n_src, F, gPl);
      ++launches;
    }
    if (eg.E > 0) {
      if (do_p1) {
        constexpr int R = NV == 4 ? 4 : 8;
        const size_t per_warp = (size_t)(R * F + 2 * F + 2 * 32 * H) * 4;
        const size_t smem = kSW * per_warp + (size_t)kSW * (R + 1) * 8;
        auto kern = eg.ascale ? edge_bwd_dst_stream_kernel<NV, R, LPH, true> : edge_bwd_dst_stream_kernel<NV, R, LPH, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int blocks = stream_grid((const void*)kern, smem, gd.n_chunks, eg.reserve_ctas);
        if (eg.kernel_events) cudaEventRecord(eg.kernel_events[2], st);
        kern<<<blocks, kSW * 32, smem, st>>>(gd, colx, Pl, Pr, a, sh, gH, cdot, score, mx, sinv, gPr, rec, part,
                                             ga_partials, galpha_dbg, H);
        if (eg.kernel_events) cudaEventRecord(eg.kernel_events[3], st);
        *n_partials = blocks;
        ++launches;
        if (gd.n_chunks > 1) {
          edge_sum_fixup_kernel<NV><<<(gd.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(gd, part, gPr);
          ++launches;
        }
      }
      if (do_p2) {
        constexpr int R = ring_depth<NV>();
        const int slot_floats = F + (rec_words(H, NV) + 31) / 32 * 32;
        const size_t smem = (size_t)kSW * R * slot_floats * 4 + (size_t)kSW * R * 8;
        auto kern = edge_bwd_src_stream_kernel<NV, R>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int blocks = stream_grid((const void*)kern, smem, gs.n_chunks, eg.reserve_ctas);
        if (eg.kernel_events) cudaEventRecord(eg.kernel_events[4], st);
        kern<<<blocks, kSW * 32, smem, st>>>(gs, cdstx, eg.csc_eid, a, sh, gH, rec, gPl, part, slot_floats);
        if (eg.kernel_events) cudaEventRecord(eg.kernel_events[5], st);
        ++launches;
        if (gs.n_chunks > 1) {
          edge_sum_fixup_kernel<NV><<<(gs.n_chunks - 1 + kSW - 1) / kSW, kSW * 32, 0, st>>>(gs, part, gPl);
          ++launches;
        }
      }
    }
  });
  return launches;
}

// chunk_row[c] = row that contains edge c*T (binary search on row_ptr)
__global__ void chunk_rows_kernel(const int* __restrict__ row_ptr, int n_rows, int E, int T, int n_chunks,
                                  int* __restrict__ chunk_row) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const int e = c * T;
  int lo = 0, hi = n_rows;  // smallest r with row_ptr[r + 1] > e
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (row_ptr[mid + 1] > e) hi = mid; else lo = mid + 1;
  }
  chunk_row[c] = lo;
}
int launch_chunk_rows(const int* row_ptr, int n_rows, int64_t E, int T, int n_chunks, int* chunk_row, cudaStream_t st) {
  if (n_chunks <= 0) return 0;
  chunk_rows_kernel<<<(n_chunks + 255) / 256, 256, 0, st>>>(row_ptr, n_rows, (int)E, T, n_chunks, chunk_row);
  return 1;
}

}  // namespace gatx
