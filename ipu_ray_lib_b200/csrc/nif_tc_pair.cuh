// PTX wrappers of the CTA-pair NIF kernel (nif_tc_pair2.cuh): tcgen05.mma.cta_group::2 (M = 256 over the two SMs of a
// cluster, each SM holding half of the B operand), its multicast commit, cluster addressing, and the TMA tensor copy
// that lets BOTH CTAs signal the leader's mbarrier directly (cp.async.bulk.tensor ... .cta_group::2).
// EXPERIMENT BUILD ONLY (-DB200RT_NIF_PAIR_KERNEL, scripts/build_variant.sh): see scripts/experiments/README.md.
#pragma once
#include <cuda.h>
#include "nif_tc.cuh"

namespace rt {
namespace tc {


__device__ __forceinline__ uint32_t mapa_u32(uint32_t smemAddr, uint32_t ctaRank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smemAddr), "r"(ctaRank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t clusterAddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dstSmem, uint32_t cols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dstSmem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma2_f16_lo(uint32_t dTmem, uint32_t aLo, uint32_t bLo, uint32_t descHi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(dTmem),
      "r"(aLo), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_commit_elect(uint64_t* bar) {  // arrives on this barrier in both CTAs
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// 2-D TMA tensor copy global -> this CTA's shared memory whose completion (transaction bytes) is counted on the mbarrier
// of the PAIR'S LEADER: with .cta_group::2 the barrier operand may name the peer CTA (CUTLASS SM100_TMA_2SM_LOAD masks
// the CTA-rank bit of the shared::cluster address the same way), so no relay through the peer is needed.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* dstSmem, const CUtensorMap* map, int x, int y, uint64_t* leaderBarLocalAlias) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dstSmem)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(leaderBarLocalAlias) & kPeerBitMask)
      : "memory");
}

}  // namespace tc
}  // namespace rt
