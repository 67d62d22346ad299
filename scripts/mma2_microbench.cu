// Semantics check + timing of tcgen05.mma.cta_group::2 (a CTA pair on one TPC) with the operand layouts the NIF kernel
// uses: fp16 K-major no-swizzle core-matrix tiles, A = 128 rows per CTA, B = N/2 columns per CTA, D = fp32 in each
// CTA's TMEM (its own 128 rows, all N columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma2 scripts/mma2_microbench.cu && timeout 60 /tmp/mma2
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <vector>

constexpr int kRows = 128, kN = 160, kK = 64, kNh = kN / 2;
constexpr int kPlaneA = kRows * 16, kPlaneB = kNh * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lboBytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sboBytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma2_kernel(const __half* A, const __half* B, float* D, int iters, int commitEvery, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                              // [K/8][128][8]
  unsigned char* sB = sA + (kK / 8) * kPlaneA;            // [K/8][N/2][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (kK / 8) * kPlaneB);
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // A rows [128 rank, 128 rank + 128), B columns [80 rank, 80 rank + 80)
  for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    reinterpret_cast<__half*>(sA)[((k / 8) * kRows + r) * 8 + (k % 8)] = A[(size_t)(rank * kRows + r) * kK + k];
  }
  for (int i = threadIdx.x; i < kNh * kK; i += blockDim.x) {
    const int n = i / kK, k = i % kK;
    reinterpret_cast<__half*>(sB)[((k / 8) * kNh + n) * 8 + (k % 8)] = B[(size_t)(rank * kNh + n) * kK + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1u << 20);  // never completes: only the cost of committing onto it is measured
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemBase = *tmemPtr;

  long long t0 = 0, t1 = 0;
  if (rank == 0 && warp == 1) {
    // the leader issues for the pair: M = 256 (128 rows per CTA), N = 160 (80 columns of B per CTA)
    const uint32_t idesc = instr_desc(2 * kRows, kN);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (commitEvery && it % commitEvery == 0 && it) {
        // a commit in the MMA stream, as the ring-stage release of the NIF pair kernels issues it (nobody waits on bar2)
        asm volatile(
            "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
            "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
                smem_u32(bar + 1)),
            "h"((uint16_t)3)
            : "memory");
      }
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * kPlaneA, kPlaneA, 128);
        const uint64_t db = smem_desc(smem_u32(sB) + ks * 2 * kPlaneB, kPlaneB, 128);
        const uint32_t acc = ks > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemBase),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    // completion of everything issued so far, signalled on the barrier at this offset in BOTH CTAs
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  }
  mbar_wait(bar, 0);
  if (rank == 0 && warp == 1) t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // every CTA reads its own 128 rows x 160 columns
  {
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmemBase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < kN; c += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + (uint32_t)c)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)(rank * kRows + row) * kN + c + j] = __uint_as_float(r[j]);
    }
  }
  if (rank == 0 && warp == 1 && lane == 0) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(256u) : "memory");
}

int main() {
  std::vector<__half> hA(2 * kRows * kK), hB(kN * kK);
  for (int r = 0; r < 2 * kRows; ++r)
    for (int k = 0; k < kK; ++k) hA[(size_t)r * kK + k] = __float2half((float)((r * 7 + k * 3) % 11 - 5) * 0.25f);
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kK; ++k) hB[(size_t)n * kK + k] = __float2half((float)((n * 5 + k) % 7 - 3) * 0.5f);
  __half *dA, *dB;
  float* dD;
  long long* dCyc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 2 * kRows * kN * 4); cudaMalloc(&dCyc, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = (kK / 8) * (kPlaneA + kPlaneB) + 128;
  cudaFuncSetAttribute(mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 4; ++mode) {
    const int iters = mode == 0 ? 1 : 1000;
    const int commitEvery = mode == 2 ? 1 : (mode == 3 ? 4 : 0);  // a multicast commit every 4 / every 16 MMAs
    cudaMemset(dD, 0xff, 2 * kRows * kN * 4);
    mma2_kernel<<<2, 128, smem>>>(dA, dB, dD, iters, commitEvery, dCyc);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(2 * kRows * kN);
    long long cyc = 0;
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
    int bad = 0, firstBad = -1;
    for (int r = 0; r < 2 * kRows; ++r)
      for (int n = 0; n < kN; ++n) {
        float ref = 0.f;
        for (int k = 0; k < kK; ++k) ref += __half2float(hA[(size_t)r * kK + k]) * __half2float(hB[(size_t)n * kK + k]);
        if (std::fabs(ref - hD[(size_t)r * kN + n]) > 1e-3f) { if (firstBad < 0) firstBad = r * kN + n; ++bad; }
      }
    std::printf("iters %d, commit every %d x 4 MMAs: %d of %d outputs wrong", iters, commitEvery, bad, 2 * kRows * kN);
    if (bad) std::printf(" (first at row %d col %d: got %g)", firstBad / kN, firstBad % kN, hD[firstBad]);
    std::printf("; %lld cycles = %.1f per MMA (M=256 across the pair, N=%d, K=16)\n", cyc, (double)cyc / (iters * (kK / 16)), kN);
  }
  return 0;
}
