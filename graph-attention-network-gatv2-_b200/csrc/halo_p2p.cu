// Halo exchange over NVLink peer memory: the two per-layer exchange steps of the
// destination-row partition, written as kernels that load / store the peers' buffers directly instead of NCCL
// collectives over the full [N][F] matrices.
//
//   forward : rank r owns rows [r0, r1) of P_l = W_l x.  halo_push_kernel reads each own row ONCE and stores it into
//             the P_l buffer of every peer whose edge slice references that source (bit p of ref_mask[row]); peers
//             that never gather the row do not receive it.  On the products-shaped R-MAT graph a rank references
//             93 / 82 / 67 % of all sources at 2 / 4 / 8 ranks, so the halo moves 0.86 / 0.76 / 0.63 of the bytes an
//             all-gather moves.
//   backward: every rank holds partial sums gP_l[src] over ITS edges for all sources it references.  Default
//             (halo_pull_kernel): the owner of a row reads the partial rows of exactly the peers whose mask bit is set
//             and adds them in ascending rank order -- a fixed order, so the result is reproducible run to run (NCCL's
//             reduction tree gives no such promise across topologies).  Alternative (GATX_HALO_MODE=bulk): the senders
//             scatter their partial rows into the owners' staging buffers with bulk copies (posted stores only) and
//             halo_sum_kernel adds them on the owner, from local memory, in the same rank order.
// Both kernels are one warp per row with 128-bit accesses; rows are 0.5-2 KB, NVLink sees full-line transfers.
// Ordering between ranks is by halo_barrier_kernel: flags in peer memory (st.release.sys / ld.acquire.sys), one
// 32-thread CTA per barrier on the exchange stream -- no NCCL call, no host involvement.  The exchange runs on a stream
// of its own, block of destination rows by block, so that it overlaps the edge passes (gatx_api.cu).
#include "common.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace gatx {

// The streaming edge kernels run with the SM's shared-memory carve-out at its maximum (64-70 KB per CTA, three CTAs
// per SM).  A kernel that prefers the default carve-out cannot share an SM with them: the SM has to drain before its
// L1 / shared split is changed, so an exchange CTA would keep the whole SM away from the edge pass for as long as it
// runs (measured: the edge forward made no progress underneath the push).  Ask for the same carve-out.
template <typename K>
static void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// Two rows per warp and trip: the loads of both rows (4 KB) are in flight before the first store is issued, which
// halves the number of warps needed to keep the link busy (the kernel runs in a fixed, small number of CTA slots).
__global__ void __launch_bounds__(256, 4)
halo_push_kernel(const float* __restrict__ own_rows, int r0, int n_rows, int F, const uint16_t* __restrict__ ref_mask,
                 PeerPtrs peers, int me) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t not_me = ~(1u << me);
  for (int rowA = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rowA < n_rows; rowA += 2 * warps) {
    const int rowB = rowA + warps;
    const uint32_t mA = ref_mask[rowA] & not_me, mB = rowB < n_rows ? (ref_mask[rowB] & not_me) : 0u;
    if (!(mA | mB)) continue;
    const float* srcA = own_rows + (int64_t)rowA * F;
    const float* srcB = own_rows + (int64_t)rowB * F;
    const int64_t offA = (int64_t)(r0 + rowA) * F, offB = (int64_t)(r0 + rowB) * F;
    for (int k0 = 0; k0 < F; k0 += 512) {  // up to 8 float4 per lane in flight
      float4 va[4], vb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 4 * (lane + 32 * j);
        if (k < F && mA) va[j] = ldg4(srcA + k);
        if (k < F && mB) vb[j] = ldg4(srcB + k);
      }
      uint32_t mm = mA;
      while (mm) {
        const int p = __ffs(mm) - 1;
        mm &= mm - 1;
        float* dst = peers.p[p] + offA;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * (lane + 32 * j);
          if (k < F) st4(dst + k, va[j]);
        }
      }
      mm = mB;
      while (mm) {
        const int p = __ffs(mm) - 1;
        mm &= mm - 1;
        float* dst = peers.p[p] + offB;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * (lane + 32 * j);
          if (k < F) st4(dst + k, vb[j]);
        }
      }
    }
  }
  __threadfence_system();  // the peer-memory stores are performed before the kernel (and the barrier behind it) completes
}

// ---- bulk-copy (TMA) transport ------------------------------------------------------------------------------------------
// Driving NVLink with ld / st from the SMs needs hundreds of warps to keep enough bytes in flight, and those warps take
// CTA slots and issue bandwidth from the edge pass that runs at the same time.  Here one elected lane per warp moves whole
// rows with bulk async copies: global -> shared ring (cp.async.bulk, mbarrier completion), then shared -> the peers'
// global memory (cp.async.bulk.global.shared::cta, bulk groups).  A warp keeps LA rows of loads and LA rows of stores in
// flight from a ring of 2 LA slots without holding a byte in registers, so 32 CTAs of 4 warps saturate the link.
// job.mode 0: forward push of own rows to the peers of ref_mask; 1: backward scatter of partial rows to their owner.
struct BulkJob {
  int mode, F, me, n_items;
  const float* src;             // mode 0: first own row of the block; mode 1: partial gP_l, global rows
  int64_t push_off;             // mode 0: (global row of the block's first row) * F
  const uint16_t* ref_mask;     // mode 0: [n_items]
  const unsigned char* my_ref;  // mode 1: [N]
  PeerPtrs base;                // mode 0: peers' P_l; mode 1: owners' staging buffers
  ScatterPlan plan;             // mode 1 (dst[] = owner's slot of this rank, relative to owner_row0)
};
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_all() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int kBulkWarps = 4;
// shared memory: ring [warps][2 LA][row] | mbarriers [warps][2 LA] u64 | offsets [warps][2 LA] i64 | masks [warps][2 LA] u32
template <int LA>
__global__ void __launch_bounds__(kBulkWarps * 32)
halo_bulk_kernel(BulkJob job) {
  constexpr int NS = 2 * LA;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)job.F * 4u;
  const uint32_t ring_s = smem_u32(smem_raw) + (uint32_t)warp * NS * row_bytes;
  uint8_t* tail = smem_raw + (size_t)kBulkWarps * NS * row_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail) + warp * NS;
  int64_t* meta_off = reinterpret_cast<int64_t*>(tail + (size_t)kBulkWarps * NS * 8) + warp * NS;
  uint32_t* meta_mask = reinterpret_cast<uint32_t*>(tail + (size_t)kBulkWarps * NS * 16) + warp * NS;
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const uint32_t bar_s = smem_u32(bar);
  const int total_warps = gridDim.x * kBulkWarps;
  uint32_t issued = 0, consumed = 0;  // warp-uniform; lane 0 alone issues the asynchronous operations (bulk groups are
                                      // per-thread state)

  auto consume = [&]() {  // stores of the oldest loaded row
    if (lane == 0) {
      const uint32_t slot = consumed % NS, ph = (consumed / NS) & 1;
      mbar_wait_s(bar_s + slot * 8u, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t m = meta_mask[slot];
      const int64_t off = meta_off[slot];
      while (m) {
        const int p = __ffs(m) - 1;
        m &= m - 1;
        bulk_s2g(job.base.p[p] + off, ring_s + slot * row_bytes, row_bytes);
      }
      bulk_commit_all();
    }
    ++consumed;
  };
  auto step = [&](const float* src, uint32_t mask, int64_t off) {  // warp-uniform arguments
    if (issued - consumed >= (uint32_t)LA) consume();
    if (lane == 0) {
      bulk_wait_read_all<LA>();  // the stores that last read this slot (issued 2 LA rows ago) have drained it
      const uint32_t slot = issued % NS;
      meta_mask[slot] = mask;
      meta_off[slot] = off;
      mbar_expect_tx_s(bar_s + slot * 8u, row_bytes);
      bulk_g2s_s(ring_s + slot * row_bytes, src, row_bytes, bar_s + slot * 8u);
    }
    ++issued;
  };

  for (int base = (blockIdx.x * kBulkWarps + warp) * 32; base < job.n_items; base += total_warps * 32) {
    const int t = base + lane;
    uint32_t mask = 0;
    int64_t off = 0;
    const float* src = nullptr;
    if (t < job.n_items) {
      if (job.mode == 0) {
        mask = job.ref_mask[t] & ~(1u << job.me);
        off = job.push_off + (int64_t)t * job.F;
        src = job.src + (int64_t)t * job.F;
      } else {
        int sgm = 0;
        while (t >= job.plan.cum[sgm + 1]) ++sgm;
        const int row = job.plan.row0[sgm] + (t - job.plan.cum[sgm]);
        if (job.my_ref[row]) {
          // plan.dst[sgm] - base.p[owner] is constant per segment; keep the offset relative to the owner's base
          const int owner = job.plan.owner[sgm];
          mask = 1u << owner;
          off = (job.plan.dst[sgm] - job.base.p[owner]) + (int64_t)(row - job.plan.owner_row0[sgm]) * job.F;
          src = job.src + (int64_t)row * job.F;
        }
      }
    }
    uint32_t bal = __ballot_sync(0xffffffffu, mask != 0);
    while (bal) {
      const int l = __ffs(bal) - 1;
      bal &= bal - 1;
      const uint32_t m_u = __shfl_sync(0xffffffffu, mask, l);
      const int64_t off_u = __shfl_sync(0xffffffffu, off, l);
      const uint64_t src_u = __shfl_sync(0xffffffffu, (uint64_t)(uintptr_t)src, l);
      step(reinterpret_cast<const float*>((uintptr_t)src_u), m_u, off_u);
    }
  }
  while (consumed < issued) consume();
  if (lane == 0) bulk_wait_all();  // every store has been performed, not only read out of shared memory
  __syncwarp();
  __threadfence_system();
}

// One warp per own row.  The peers that hold a partial row are visited in ascending rank order (a fixed order: the sum
// is reproducible), software-pipelined: the loads of the next contributing peer are in flight while the current one is
// added, so a warp keeps two remote 2 KB reads outstanding instead of one (NVLink round trips are ~3 us).
__global__ void __launch_bounds__(256, 4)
halo_pull_kernel(float* __restrict__ own_rows, int r0, int n_rows, int F, const uint16_t* __restrict__ ref_mask,
                 PeerPtrs peers, int me, int world) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
    const uint32_t m = ref_mask[row];
    if ((m & ~(1u << me)) == 0) continue;  // only this rank (or nobody) touches the row: already complete
    float* mine = own_rows + (int64_t)row * F;
    const int64_t off = (int64_t)(r0 + row) * F;
    for (int k0 = 0; k0 < F; k0 += 512) {
      float4 acc[4], nxt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t mm = m;
      int p = __ffs(mm) - 1;
      mm &= mm - 1;
      {
        const float* src = p == me ? mine : peers.p[p] + off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * (lane + 32 * j);
          nxt[j] = k < F ? ld_sys4(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);  // peer memory: system-scope load
        }
      }
      while (true) {
        float4 cur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        const bool more = mm != 0;
        if (more) {
          p = __ffs(mm) - 1;
          mm &= mm - 1;
          const float* src = p == me ? mine : peers.p[p] + off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + 4 * (lane + 32 * j);
            nxt[j] = k < F ? ld_sys4(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j].x += cur[j].x; acc[j].y += cur[j].y; acc[j].z += cur[j].z; acc[j].w += cur[j].w;
        }
        if (!more) break;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 4 * (lane + 32 * j);
        if (k < F) st4(mine + k, acc[j]);
      }
    }
  }
}

// Backward exchange, owner side (local memory only): own row = own partial + the staged partials of the ranks whose
// mask bit is set, added in ascending rank order -- a fixed order, so the sum is reproducible run to run.
__global__ void __launch_bounds__(256)
halo_sum_kernel(float* __restrict__ own_rows, int row_off, int n_rows, int n_rows_total, int F,
                const uint16_t* __restrict__ ref_mask, const float* __restrict__ stage, int me, int world) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
    const uint32_t m = ref_mask[row];
    if ((m & ~(1u << me)) == 0) continue;  // only this rank (or nobody) touches the row: already complete
    float* mine = own_rows + (int64_t)row * F;
    for (int k0 = 0; k0 < F; k0 += 512) {
      float4 acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < world; ++p) {
        if (!((m >> p) & 1u)) continue;
        const float* src = p == me ? mine : stage + ((int64_t)p * n_rows_total + row_off + row) * F;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * (lane + 32 * j);
          if (k < F) {
            const float4 v = ld_sys4(src + k);  // written by a peer over NVLink: never a stale cached line
            acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 4 * (lane + 32 * j);
        if (k < F) st4(mine + k, acc[j]);
      }
    }
  }
}

__global__ void halo_barrier_kernel(PeerFlags flags, int me, int world, uint32_t seq) {
  const int p = threadIdx.x;
  if (p >= world || p == me) return;
  __threadfence_system();
  // arrive: slot `me` of peer p's array
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[p] + me), "r"(seq) : "memory");
  // wait for peer p's arrival in slot p of my array (flags only grow: a peer may already be one barrier ahead)
  const uint32_t* mine = flags.p[me] + p;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - seq) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) {  // 20 s: a peer died or the ranks issued different exchange sequences
      printf("gatx: halo barrier %u timed out on rank %d waiting for rank %d (flag %u)\n", seq, me, p, v);
      __trap();
    }
    __nanosleep(64);
  }
  __threadfence_system();
}

int launch_halo_barrier(const PeerFlags& flags, int me, int world, uint32_t seq, cudaStream_t st) {
  prefer_max_shared(halo_barrier_kernel);
  halo_barrier_kernel<<<1, 32, 0, st>>>(flags, me, world, seq);
  return 1;
}

// Both kernels are CTAs of 256 threads with at most 64 registers per thread (launch bounds): 16 K registers and no
// shared memory, i.e. one exchange CTA fits the slot of one streaming edge CTA (~20 K registers).  In the pipelined
// epoch they run in halo_cta_slots() CTAs underneath an edge pass whose grid leaves exactly that many slots free.
// GATX_HALO_CTAS=S runs the exchange kernels in S CTA slots and makes the streaming edge kernels leave exactly S slots free
// (their chunks are assigned statically, so every CTA of their grid must be resident).  Measured on the products shape
// (profiles/r2_multi_gpu_transports.md): few slots starve the exchange (S = 24: pull at 98 GB/s), many starve the edge pass
// (S = 148 on 2 GPUs: 114 ms instead of 102), and no S beats the default -- no cap and no reservation: the exchange kernels
// take whatever the edge pass leaves and the edge pass absorbs the imbalance.
int halo_cta_slots(int /*world*/) {
  static const int forced = [] {
    const char* e = getenv("GATX_HALO_CTAS");
    const int v = e ? atoi(e) : 0;
    return v >= 4 && v <= kNumSMs * 2 ? v : 0;
  }();
  return forced;
}
static int halo_blocks(int n_rows, int max_ctas) {
  int blocks = (n_rows + 7) / 8;
  const int cap = max_ctas > 0 ? max_ctas : kNumSMs * 8;
  return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

int launch_halo_push(const float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers,
                     int me, cudaStream_t st, int max_ctas) {
  if (n_rows <= 0) return 0;
  prefer_max_shared(halo_push_kernel);
  halo_push_kernel<<<halo_blocks(n_rows, max_ctas), 256, 0, st>>>(own_rows, r0, n_rows, F, ref_mask, peers, me);
  return 1;
}

static int launch_bulk(const BulkJob& job, cudaStream_t st, int max_ctas) {
  if (job.n_items <= 0) return 0;
  const size_t ring = (size_t)16384;  // bytes of ring per warp
  int la = (int)(ring / ((size_t)job.F * 4) / 2);
  la = la >= 16 ? 16 : (la >= 8 ? 8 : (la >= 4 ? 4 : 2));
  const size_t smem = (size_t)kBulkWarps * 2 * la * job.F * 4 + (size_t)kBulkWarps * 2 * la * (8 + 4 + 8) + 64;
  int blocks = (job.n_items + kBulkWarps * 32 - 1) / (kBulkWarps * 32);
  const int cap = max_ctas > 0 ? max_ctas : kNumSMs;
  if (blocks > cap) blocks = cap;
#define BULK_LAUNCH(LA)                                                                                          \
  do {                                                                                                           \
    cudaFuncSetAttribute(halo_bulk_kernel<LA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    prefer_max_shared(halo_bulk_kernel<LA>);                                                                     \
    halo_bulk_kernel<LA><<<blocks, kBulkWarps * 32, smem, st>>>(job);                                            \
  } while (0)
  if (la == 16) BULK_LAUNCH(16);
  else if (la == 8) BULK_LAUNCH(8);
  else if (la == 4) BULK_LAUNCH(4);
  else BULK_LAUNCH(2);
#undef BULK_LAUNCH
  return 1;
}
bool halo_bulk_supported(int F) { return F % 4 == 0 && (size_t)F * 4 * 2 * 2 * kBulkWarps <= 200 * 1024; }

int launch_halo_push_bulk(const float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers,
                          int me, cudaStream_t st, int max_ctas) {
  BulkJob job{};
  job.mode = 0; job.F = F; job.me = me; job.n_items = n_rows; job.src = own_rows; job.push_off = (int64_t)r0 * F;
  job.ref_mask = ref_mask; job.base = peers;
  return launch_bulk(job, st, max_ctas);
}
int launch_halo_scatter_bulk(const float* partial, int F, const unsigned char* my_ref, const ScatterPlan& plan,
                             const PeerPtrs& stage_base, int me, cudaStream_t st, int max_ctas) {
  BulkJob job{};
  job.mode = 1; job.F = F; job.me = me; job.n_items = plan.cum[plan.n_seg]; job.src = partial; job.my_ref = my_ref;
  job.base = stage_base; job.plan = plan;
  return launch_bulk(job, st, max_ctas);
}

int launch_halo_pull(float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers, int me,
                     int world, cudaStream_t st, int max_ctas) {
  if (n_rows <= 0) return 0;
  prefer_max_shared(halo_pull_kernel);
  halo_pull_kernel<<<halo_blocks(n_rows, max_ctas), 256, 0, st>>>(own_rows, r0, n_rows, F, ref_mask, peers, me, world);
  return 1;
}

int launch_halo_sum(float* own_rows, int row_off, int n_rows, int n_rows_total, int F, const uint16_t* ref_mask,
                    const float* stage, int me, int world, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  halo_sum_kernel<<<halo_blocks(n_rows, 0), 256, 0, st>>>(own_rows, row_off, n_rows, n_rows_total, F, ref_mask, stage, me,
                                                          world);
  return 1;
}

}  // namespace gatx
