// tcgen05 tensor-core GEMMs for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
// These are the dense contractions the reference recomputes per edge with scalar FMAs:
//   projection   P_l|P_r = X [W_l;W_r]^T                (EB:303-316, EB:415-420)      -> gemm_tn
//   input grad   gX = gP_l W_l + gP_r W_r               (EB:859-869)                  -> gemm_tn (dual operand)
//   weight grad  gW_l|gW_r = gP^T X                     (EB:771-782)                  -> gemm_atb (MN-major operands)
// Inputs stay fp32 in HBM; kind::tf32 reads the fp32 words and uses their top 19 bits, accumulation is
// fp32.  Stated tolerance (tests/test_gpu_parity.py): 5e-3 relative on forward tensors.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (each owns the TMEM lane quadrant warp_id % 4).  One CTA computes one 128 x BN
// output tile; two CTAs fit per SM so one tile's epilogue overlaps the other's main loop.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace gatx {
namespace {

constexpr int BM = 128;      // UMMA_M (cta_group::1)
constexpr int BK = 32;       // fp32 elements per k-block = 128 bytes = one swizzle span
constexpr int UMMA_K = 8;    // tf32
constexpr int kThreads = 192;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store of a box from (swizzled) shared memory; completion is tracked by bulk async-groups of the issuing thread
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], single-thread issue
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version = 1 in [46,48),
// layout type in [61,64) (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format TF32 = 2 at
// [7,10) / [10,13), a_major at 15, b_major at 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct TnSmem {
  static constexpr int kStages = BN >= 256 ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr int kABytes = BM * BK * 4, kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue buffers alias the drained pipeline stages: 4 warps x 2 x (32 rows x 128 B) for the TMA-store path,
  // 4 x 32 x 33 floats for the scalar path
  static constexpr int kStagingFloats = 4 * 2 * 32 * 32;
  static_assert(kStages * kStageBytes >= kStagingFloats * 4, "staging must fit in the pipeline buffers");
  static constexpr int kBytes = 1024 /*align slack*/ + kStages * kStageBytes + 256;
};

// C[m][n] (+)= sum_k A0[m][k] B0[n][k] + sum_k A1[m][k] B1[n][k]   (second pair optional, K1 = 0)
// columns n >= n_split go to C1 (column n - n_split); used to write P_l and P_r from one pass over X.
template <int BN>
__global__ void __launch_bounds__(kThreads)
gemm_tf32_tn_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1, int K0,
                    int K1, float* __restrict__ C0, float* __restrict__ C1, int n_split, int64_t ldc, int M, int N,
                    int accumulate, int tma_store) {
  using S = TnSmem<BN>;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem);  // reused only after tmem_full_bar (all TMA + MMA done)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb0 = (K0 + BK - 1) / BK, kb1 = (K1 + BK - 1) / BK, num_kb = kb0 + kb1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (K1 > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    if (tma_store) {
      tma_prefetch_desc(&tmC0);
      tma_prefetch_desc(&tmC1);
    }
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::kStages;
        const uint32_t ph = (kb / S::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = tiles + s * S::kStageBytes;
        uint8_t* b_dst = a_dst + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        if (kb < kb0) {
          tma_load_2d(&tmA0, &full_bar[s], a_dst, kb * BK, m0);
          tma_load_2d(&tmB0, &full_bar[s], b_dst, kb * BK, n0);
        } else {
          tma_load_2d(&tmA1, &full_bar[s], a_dst, (kb - kb0) * BK, m0);
          tma_load_2d(&tmB1, &full_bar[s], b_dst, (kb - kb0) * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN < 16 ? 16 : BN, 0, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::kStages;
        const uint32_t ph = (kb / S::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major SWIZZLE_128B: rows are 128 B apart, 8-row groups 1024 B apart (SBO); a UMMA_K step of
          // 8 tf32 = 32 B advances the start address inside the swizzle span
          const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 4, 16, 1024);
          const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 4, 16, 1024);
          umma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (tma_store) {
      // TMEM -> registers -> 128B-swizzled shared-memory box (32 rows x 32 floats) -> TMA store.  Lane = tile row, so
      // a lane writes its 128-byte row as eight 16-byte chunks at chunk ^ (row & 7): the quarter-warps of every
      // st.shared.v4 hit distinct banks, and the tensor map (SWIZZLE_128B) undoes the permutation.  Two boxes per
      // warp alternate, so the store engine drains one while the next 32 columns are read from TMEM.  Rows / columns
      // past M / N are clipped by the tensor map.
      const uint32_t box0 = smem_u32(staging) + (uint32_t)(warp - 2) * 8192u;
      const bool to_c1 = n0 >= n_split;
      const CUtensorMap* tmc = to_c1 ? &tmC1 : &tmC0;
      const int ccol0 = to_c1 ? n0 - n_split : n0;
      const int row0 = m0 + q * 32;
      const uint32_t rsw = (uint32_t)(lane & 7);
      int buf = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (lane == 0) bulk_wait_read<1>();  // the store that last read this box (two chunks ago) is done with it
        __syncwarp();
        const uint32_t box = box0 + (uint32_t)buf * 4096u;
        const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts4(rowaddr + ((((uint32_t)c) ^ rsw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < M) {
          tma_store_2d(tmc, reinterpret_cast<const void*>(staging + ((warp - 2) * 8192 + buf * 4096) / 4), ccol0 + c0, row0);
          bulk_commit();
        }
        buf ^= 1;
      }
      if (lane == 0) bulk_wait_read<0>();  // shared memory must stay valid until the last store has read it
      __syncwarp();
    } else {
      // scalar path (accumulate mode, or an output whose pitch / base no tensor map can describe):
      // per-warp smem transpose -> 128-byte row stores
      float* st = staging + (warp - 2) * 32 * 33;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int col = n0 + c0 + lane;
        float* Cb = C0;
        int ccol = col;
        if (col >= n_split) {
          Cb = C1;
          ccol = col - n_split;
        }
        if (col < N && c0 + lane < BN) {
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const int row = m0 + q * 32 + r;
            if (row < M) {
              float* p = Cb + (int64_t)row * ldc + ccol;
              const float val = st[r * 33 + lane];
              *p = accumulate ? *p + val : val;
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---- persistent variant for 256-column tiles ---------------------------------------------------------------
// One CTA per SM walks the tile list (n fastest, so the CTAs that run at the same time share the A tile in L2):
// a 4-stage TMA ring (4 x 48 KB) feeds the MMA thread across tile boundaries, the accumulator is double-buffered
// in TMEM (2 x 256 columns), and the four epilogue warps drain tile i (TMEM -> swizzled smem boxes -> TMA store)
// while the tensor core already works on tile i + 1.  Compared with one tile per CTA (two co-resident CTAs with a
// 2-stage ring each) this removes the per-tile TMEM allocation / barrier setup and keeps four k-blocks in flight
// for a single main loop.
struct TnPersist {
  static constexpr int BN = 256, kStages = 4;
  static constexpr int kABytes = BM * BK * 4, kBBytes = BN * BK * 4, kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = 8;  // two per TMEM lane quadrant, each takes one half of the 256 columns
  static constexpr int kThreadsP = 64 + 32 * kEpiWarps;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kBytes = 1024 + kStages * kStageBytes + kStagingBytes + 256;
};

__global__ void __launch_bounds__(TnPersist::kThreadsP)
gemm_tf32_tn_persistent_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                               const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
                               int K0, int K1, int n_split, int M, int N, int tiles_n, int num_tiles, GemmFuse fuse) {
  using S = TnPersist;
  constexpr int BN = S::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint8_t* staging = smem + S::kStages * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + S::kStagingBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = (K0 + BK - 1) / BK, kb1 = (K1 + BK - 1) / BK, num_kb = kb0 + kb1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (K1 > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    tma_prefetch_desc(&tmC0);
    tma_prefetch_desc(&tmC1);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], S::kEpiWarps);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t kbg = 0;  // k-blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % S::kStages;
          const uint32_t ph = (kbg / S::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = tiles + s * S::kStageBytes;
          uint8_t* b_dst = a_dst + S::kABytes;
          mbar_expect_tx(&full_bar[s], S::kStageBytes);
          if (kb < kb0) {
            tma_load_2d(&tmA0, &full_bar[s], a_dst, kb * BK, m0);
            tma_load_2d(&tmB0, &full_bar[s], b_dst, kb * BK, n0);
          } else {
            tma_load_2d(&tmA1, &full_bar[s], a_dst, (kb - kb0) * BK, m0);
            tma_load_2d(&tmB1, &full_bar[s], b_dst, (kb - kb0) * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0, 0);
      uint32_t kbg = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % S::kStages;
          const uint32_t ph = (kbg / S::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + s * S::kStageBytes);
          const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 4, 16, 1024);
            const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 4, 16, 1024);
            umma_tf32(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===== epilogue (see gemm_tf32_tn_kernel): TMEM -> registers -> swizzled smem box -> TMA store =====
    // Eight warps: warp w may only touch TMEM lanes 32 (w % 4) .. + 31, so warps w and w + 4 share a lane quadrant
    // and split the tile's columns.  Draining a 128 x 256 fp32 tile with four warps took longer than its main loop.
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* box_ptr = staging + (warp - 2) * 4096;
    const uint32_t box = smem_u32(box_ptr);
    const uint32_t rsw = (uint32_t)(lane & 7);
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      const bool to_c1 = n0 >= n_split;
      const CUtensorMap* tmc = to_c1 ? &tmC1 : &tmC0;
      const int ccol0 = to_c1 ? n0 - n_split : n0;
      const int row0 = m0 + q * 32;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      float hdot = 0.f;  // fused epilogue: running g . Hout (- g_pre . bias) of the head the chunks belong to
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t v[32];
        float4 ho[8];
        const int frow = row0 + lane, fcol = n0 + c0;
        const bool fz = fuse.Hout != nullptr && frow < M;
        if (fz) {  // this lane's row of the layer output: 128 contiguous bytes, in flight while the accumulator is read
          const float4* hp = reinterpret_cast<const float4*>(fuse.Hout + (int64_t)frow * fuse.ld + fcol);
#pragma unroll
          for (int c = 0; c < 8; ++c) ho[c] = __ldg(hp + c);
        }
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, v);
        if (fuse.Hout != nullptr) {
          if (fz) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float h4[4] = {ho[c].x, ho[c].y, ho[c].z, ho[c].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float g = __uint_as_float(v[4 * c + t]);
                const float gp = g * (h4[t] > 0.f ? 1.f : fuse.act_slope);  // EB:879-893: LReLU'(h) has the sign of LReLU(h)
                hdot = fmaf(g, h4[t], hdot);
                if (fuse.bias) hdot = fmaf(-gp, __ldg(fuse.bias + fcol + 4 * c + t), hdot);
                v[4 * c + t] = __float_as_uint(gp);
              }
            }
          }
          if ((fcol + 32) % fuse.head_dim == 0) {  // the head ends with this chunk (head_dim % 32 == 0)
            if (fz) fuse.cdot[(int64_t)frow * fuse.heads + fcol / fuse.head_dim] = hdot;
            hdot = 0.f;
          }
        }
        if (lane == 0) bulk_wait_read<0>();  // the previous store of this warp has read the box
        __syncwarp();
        const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts4(rowaddr + ((((uint32_t)c) ^ rsw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < M) {
          tma_store_2d(tmc, box_ptr, ccol0 + c0, row0);
          bulk_commit();
        }
      }
      // every tcgen05.ld of this accumulator has completed (tmem_ld32 waits): hand it back to the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---- host side: tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] with row pitch ld (elements); box = box_cols x box_rows, 128B swizzle
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 4) % 16 || rows <= 0 || cols <= 0) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
int launch_tn(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1,
              const CUtensorMap& c0, const CUtensorMap& c1, bool tma_store, int K0,
              int K1, float* C0, float* C1, int n_split, int64_t ldc, int M, int N, bool accumulate, cudaStream_t st) {
  // per launch: the attribute is per device, and one process may drive several devices (train_gatx --gpus N)
  if (cudaFuncSetAttribute(gemm_tf32_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TnSmem<BN>::kBytes) !=
      cudaSuccess)
    return -1;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32_tn_kernel<BN><<<grid, kThreads, TnSmem<BN>::kBytes, st>>>(a0, b0, a1, b1, c0, c1, K0, K1, C0, C1, n_split,
                                                                      ldc, M, N, accumulate ? 1 : 0, tma_store ? 1 : 0);
  return 1;
}

int pick_bn(int N, int n_split) {
  // widest tile that does not straddle the C0/C1 split.  256 halves the re-reads of the A tile through L2
  // (N/BN column tiles share it) and still lets two CTAs share an SM (2 x 96 KB smem, 2 x 256 TMEM columns).
  static const bool wide = getenv("GATX_GEMM_BN128") == nullptr;
  for (int bn : {256, 128, 64, 32, 16}) {
    if (bn == 256 && (!wide || N < 256)) continue;
    if (n_split < N && n_split % bn) continue;
    if (bn == 16 || N >= bn || N > bn / 2) return bn;
  }
  return -1;
}


// ---------------------------------------------------------------------------------------------------
// C[m][n] += sum_k A[k][m] B[k][n]: both operands are MN-major (the contraction index k -- the node id --
// is the slow index in memory).  TMA boxes of 32 floats (one 128-byte swizzle span along m / n) x 32 rows
// (k) are laid down slab by slab.  For 32-bit MN-major operands tcgen05 accepts only the SWIZZLE_128B_BASE32B
// layout (32-byte chunks XOR-ed with the row index mod 4; TMA mode SWIZZLE_128B_ATOM_32B): k-groups of 4 rows are
// 512 B apart (SBO), 32-float slabs are 32 rows * 128 B apart (LBO), and one UMMA (K = 8) spans two k-groups.  The node dimension is split across
// CTAs (split-K); partial tiles go to a workspace and are added in a fixed order (deterministic).
constexpr int BKN = 32;  // nodes per k-block

template <int BN>
struct AtbSmem {
  static constexpr int kStages = BN >= 256 ? 4 : 6;
  static constexpr int kSlabBytes = BKN * 128;
  static constexpr int kABytes = (BM / 32) * kSlabBytes, kBBytes = (BN / 32) * kSlabBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBytes = 1024 + kStages * kStageBytes + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads)
gemm_tf32_atb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int64_t K,
                     int kb_per_split, float* __restrict__ out, int64_t ld_out, int64_t split_stride, int M, int N,
                     int n_tiles_n, int accumulate) {
  using S = AtbSmem<BN>;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x % n_tiles_n) * BN, m0 = (blockIdx.x / n_tiles_n) * BM;
  const int total_kb = (int)((K + BKN - 1) / BKN);
  const int kb_begin = blockIdx.y * kb_per_split;
  int num_kb = total_kb - kb_begin;
  if (num_kb > kb_per_split) num_kb = kb_per_split;
  if (num_kb < 0) num_kb = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::kStages;
        const uint32_t ph = (kb / S::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = tiles + s * S::kStageBytes;
        uint8_t* b_dst = a_dst + S::kABytes;
        const int krow = (kb_begin + kb) * BKN;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
#pragma unroll
        for (int j = 0; j < BM / 32; ++j) tma_load_2d(&tmA, &full_bar[s], a_dst + j * S::kSlabBytes, m0 + 32 * j, krow);
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmB, &full_bar[s], b_dst + j * S::kSlabBytes, n0 + 32 * j, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 1, 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::kStages;
        const uint32_t ph = (kb / S::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
        for (int kg = 0; kg < BKN / UMMA_K; ++kg) {
          const uint64_t da = make_smem_desc(a_addr + kg * 1024, S::kSlabBytes, 512, 1);
          const uint64_t db = make_smem_desc(b_addr + kg * 1024, S::kSlabBytes, 512, 1);
          umma_tf32(tmem_base, da, db, idesc, (kb | kg) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    float* st = staging + (warp - 2) * 32 * 33;
    float* dst = out + (int64_t)blockIdx.y * split_stride;
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int col = n0 + c0 + lane;
      if (col < N && c0 + lane < BN) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int row = m0 + q * 32 + r;
          if (row < M) {
            float* p = dst + (int64_t)row * ld_out + col;
            const float val = st[r * 33 + lane];
            *p = accumulate ? *p + val : val;
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

__global__ void atb_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float* __restrict__ C,
                                  int64_t ldc) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * total + i];  // fixed order
    C[(i / N) * ldc + (i % N)] += s;
  }
}

template <int BN>
int launch_atb(const CUtensorMap& a, const CUtensorMap& b, int64_t K, float* C, int64_t ldc, int M, int N, float* ws,
               size_t ws_bytes, cudaStream_t st) {
  if (cudaFuncSetAttribute(gemm_tf32_atb_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           AtbSmem<BN>::kBytes) != cudaSuccess)
    return -1;
  const int tn = (N + BN - 1) / BN, tm = (M + BM - 1) / BM, tiles = tn * tm;
  const int total_kb = (int)((K + BKN - 1) / BKN);
  int splits = (kNumSMs + tiles - 1) / tiles;
  if (splits > total_kb) splits = total_kb;
  const size_t per = sizeof(float) * (size_t)M * N;
  if (splits > 1 && (!ws || per * splits > ws_bytes)) splits = ws ? (int)(ws_bytes / per) : 1;
  if (splits < 1) splits = 1;
  int kbps = (total_kb + splits - 1) / splits;
  splits = (total_kb + kbps - 1) / kbps;
  dim3 grid(tiles, splits);
  if (splits == 1) {
    gemm_tf32_atb_kernel<BN><<<grid, kThreads, AtbSmem<BN>::kBytes, st>>>(a, b, K, kbps, C, ldc, 0, M, N, tn, 1);
    return 1;
  }
  gemm_tf32_atb_kernel<BN><<<grid, kThreads, AtbSmem<BN>::kBytes, st>>>(a, b, K, kbps, ws, N, (int64_t)M * N, M, N, tn,
                                                                        0);
  const int64_t total = (int64_t)M * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  atb_reduce_kernel<<<blocks, 256, 0, st>>>(ws, splits, M, N, C, ldc);
  return 2;
}

}  // namespace

int launch_gemm_tc_tn2(const float* A0, int64_t lda0, const float* B0, int64_t ldb0, int K0, const float* A1,
                       int64_t lda1, const float* B1, int64_t ldb1, int K1, float* C0, float* C1, int n_split,
                       int64_t ldc, int M, int N, bool accumulate, cudaStream_t st, const GemmFuse* fuse) {
  if (M <= 0 || N <= 0 || K0 <= 0) return -1;
  if (n_split <= 0 || n_split > N) n_split = N;
  const int bn = pick_bn(N, n_split);
  if (bn < 0) return -1;
  const bool want_fuse = fuse && fuse->Hout;
  // the fused epilogue lives in the persistent 256-column kernel: whole heads per 128-column half, one output matrix
  if (want_fuse && (bn != 256 || accumulate || n_split != N || fuse->head_dim < 32 || fuse->head_dim % 32 || 128 % fuse->head_dim ||
                    N % fuse->head_dim || fuse->ld % 4 || (reinterpret_cast<uintptr_t>(fuse->Hout) & 15)))
    return -1;
  CUtensorMap a0, b0, a1, b1;
  if (!make_map(&a0, A0, M, K0, lda0, BK, BM) || !make_map(&b0, B0, N, K0, ldb0, BK, bn)) return -1;
  if (K1 > 0) {
    if (!make_map(&a1, A1, M, K1, lda1, BK, BM) || !make_map(&b1, B1, N, K1, ldb1, BK, bn)) return -1;
  } else {
    a1 = a0;
    b1 = b0;
  }
  // output tensor maps for the TMA-store epilogue (32 x 32 boxes, 128B swizzle); the scalar epilogue takes over when
  // the kernel must accumulate or an output cannot be described (pitch / base not 16-byte aligned)
  static const bool no_tma_store = getenv("GATX_GEMM_NO_TMA_STORE") != nullptr;
  CUtensorMap c0, c1;
  bool tma_store = !accumulate && !no_tma_store && bn >= 32;
  if (tma_store) {
    tma_store = make_map(&c0, C0, M, n_split, ldc, 32, 32);
    if (tma_store && n_split < N) tma_store = make_map(&c1, C1, M, N - n_split, ldc, 32, 32);
    else if (tma_store) c1 = c0;
  }
  if (!tma_store) {
    c0 = a0;
    c1 = a0;
  }
  static const bool no_persist = getenv("GATX_GEMM_NO_PERSISTENT") != nullptr;
  if (bn == 256 && tma_store && !no_persist) {
    if (cudaFuncSetAttribute(gemm_tf32_tn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TnPersist::kBytes) != cudaSuccess)
      return -1;
    const int tiles_n = (N + 255) / 256, tiles_m = (M + BM - 1) / BM;
    const int num_tiles = tiles_n * tiles_m;
    const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
    gemm_tf32_tn_persistent_kernel<<<grid, TnPersist::kThreadsP, TnPersist::kBytes, st>>>(a0, b0, a1, b1, c0, c1, K0, K1, n_split, M, N,
                                                                              tiles_n, num_tiles, want_fuse ? *fuse : GemmFuse{});
    return 1;
  }
  if (want_fuse) return -1;
  switch (bn) {
    case 256: return launch_tn<256>(a0, b0, a1, b1, c0, c1, tma_store, K0, K1, C0, C1, n_split, ldc, M, N, accumulate, st);
    case 128: return launch_tn<128>(a0, b0, a1, b1, c0, c1, tma_store, K0, K1, C0, C1, n_split, ldc, M, N, accumulate, st);
    case 64: return launch_tn<64>(a0, b0, a1, b1, c0, c1, tma_store, K0, K1, C0, C1, n_split, ldc, M, N, accumulate, st);
    case 32: return launch_tn<32>(a0, b0, a1, b1, c0, c1, tma_store, K0, K1, C0, C1, n_split, ldc, M, N, accumulate, st);
    default: return launch_tn<16>(a0, b0, a1, b1, c0, c1, tma_store, K0, K1, C0, C1, n_split, ldc, M, N, accumulate, st);
  }
}

int launch_gemm_tc_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                      int K, bool accumulate, cudaStream_t st) {
  return launch_gemm_tc_tn2(A, lda, B, ldb, K, nullptr, 0, nullptr, 0, 0, C, C, N, ldc, M, N, accumulate, st);
}

int launch_gemm_tc_atb(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                       int64_t K, float* ws, size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  CUtensorMap a, b;
  if (!make_map(&a, A, K, M, lda, 32, BKN, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map(&b, B, K, N, ldb, 32, BKN, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return -1;
  switch (bn) {
    case 256: return launch_atb<256>(a, b, K, C, ldc, M, N, ws, ws_bytes, st);
    case 128: return launch_atb<128>(a, b, K, C, ldc, M, N, ws, ws_bytes, st);
    default: return launch_atb<64>(a, b, K, C, ldc, M, N, ws, ws_bytes, st);
  }
}

}  // namespace gatx
