// Inline-PTX wrappers shared by the sm_100a kernels: mbarrier, bulk async copy (cp.async.bulk, SASS UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gatx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy ones (bulk copies into a
// buffer that generic loads have just read)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// 1-D bulk async copy global -> shared, completion (bytes) signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Variants on precomputed 32-bit shared-memory addresses: the streaming loops compute `smem_u32(base)` once and then
// only add slot offsets (the generic-pointer forms re-derive the shared window address -- an S2R plus a few integer
// instructions -- on every trip).
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint_s(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// Same copy with an L2 eviction-priority hint (createpolicy): rows of frequently gathered nodes are kept
// (evict_last), rows touched once are dropped first (evict_first) so they cannot push the hot set out.
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// ---- single-lane issue through elect.sync --------------------------------------------------------------------------
// `if (lane == 0) { expect_tx; cp.async.bulk }` makes ptxas emit a loop over the active lanes around UBLKCP (its operands
// live in uniform registers, and the compiler cannot prove that one lane is active): ELECT + 4 R2UR + branch, ~20 issue
// slots per copy.  With elect.sync the predicate is known to select exactly one lane, the shuffled gather index is moved
// to the uniform datapath once and the address arithmetic runs there (UIMAD.WIDE): ~8 issue slots.  Call with all 32
// lanes converged and warp-uniform operands; the condition around the call must be warp-uniform too.
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(dst),
      "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint_elect(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar,
                                                    uint64_t policy) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
      "}\n" ::"r"(dst),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// two copies completing on one barrier (a gathered row + its edge record, or the two row-buffer halves)
__device__ __forceinline__ void bulk2_g2s_hint_elect(uint32_t dst0, const void* src0, uint32_t bytes0, uint64_t policy0,
                                                     uint32_t dst1, const void* src1, uint32_t bytes1, uint64_t policy1,
                                                     uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "add.u32 t, %2, %6;\n"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%8], t;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%8], %3;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%4], [%5], %6, [%8], %7;\n"
      "}\n" ::"r"(dst0),
      "l"(src0), "r"(bytes0), "l"(policy0), "r"(dst1), "l"(src1), "r"(bytes1), "l"(policy1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk2_g2s_elect(uint32_t dst0, const void* src0, uint32_t bytes0, uint32_t dst1,
                                                const void* src1, uint32_t bytes1, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "add.u32 t, %2, %5;\n"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%6], t;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%6];\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%4], %5, [%6];\n"
      "}\n" ::"r"(dst0),
      "l"(src0), "r"(bytes0), "r"(dst1), "l"(src1), "r"(bytes1), "r"(bar)
      : "memory");
}
// 128-bit shared-memory load on a 32-bit shared address (no generic-to-shared conversion in the loop); volatile: the ring
// slots are rewritten by the async proxy, the load must stay behind its mbarrier wait
__device__ __forceinline__ float4 lds4_s(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds4u_s(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds1u_s(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// ---- cp.async (LDGSTS): 16 bytes per lane global -> shared, completion by per-thread groups ---------------------------
// A warp-wide cp.async moves one 512-byte row with ONE instruction and no uniform-register traffic (a bulk copy of the
// same row costs ~20 issue slots: elect, four R2UR, expect_tx, UBLKCP); the consumer waits with cp.async.wait_group
// instead of polling an mbarrier.  .cg: cached in L2 only (the gathered rows are used once per warp).
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

}  // namespace gatx
