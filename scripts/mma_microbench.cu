// Micro-benchmark: cycles per tcgen05.mma (M=128, fp16, K=16, no-swizzle K-major operands in smem) for a few shapes
// and commit frequencies. All loop arithmetic is mask-based (no integer division) so the numbers are the MMA pipe's.
// Run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mmab scripts/mma_microbench.cu -I ipu_ray_lib_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "nif_tc.cuh"
using namespace rt::tc;

template <int N, int ACC_MASK, int COMMIT_MASK /* -1 = never */>
__global__ void __launch_bounds__(384, 1) bench(int iters, long long* out, int spinners, int sleepNs, int randomData = 0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmemPtr;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // two fp16 values in roughly [-1, 1): exponent 0x3800-0x3bff range, random sign/mantissa
    reinterpret_cast<uint32_t*>(smem)[i] = randomData ? ((h & 0x83ff83ffu) | 0x38003800u) : 0u;
  }
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmemPtr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmemPtr;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = instr_desc(128, N);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t acc = (uint32_t)i & ACC_MASK;
      const uint32_t aAddr = a0 + (((uint32_t)i & 7u) << 12);
      const uint32_t bAddr = b0 + (((uint32_t)i & 3u) * 10240u) + acc * (uint32_t)N * 16u;
      mma_f16(tmem + acc * N, smem_desc(aAddr, 2048u, 128u), smem_desc(bAddr, 5120u, 128u), idesc, i > ACC_MASK ? 1u : 0u);
      if (COMMIT_MASK >= 0 && (i & COMMIT_MASK) == COMMIT_MASK) mma_commit(&bar[1]);
    }
    const long long t1 = clock64();
    mma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  else if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + spinners) {
    // what the epilogue warps do while the MMAs run: wait for the accumulator barrier
    if (sleepNs == 0) {
      mbar_wait(&bar[0], 0);
    } else if (sleepNs > 0) {
      if ((threadIdx.x & 31) == 0) {
        const uint32_t addr = smem_u32(&bar[0]);
        uint32_t done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(addr), "r"(0u) : "memory");
          if (!done) __nanosleep(sleepNs);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int ACC_MASK, int COMMIT_MASK>
void run(const char* name, long long* d, int grid = 1, int spinners = 0, int sleepNs = 0, int randomData = 0) {
  const int iters = 4096;
  auto k = bench<N, ACC_MASK, COMMIT_MASK>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<grid, 384, 200 * 1024>>>(iters, d, spinners, sleepNs, randomData);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double macs = 128.0 * N * 16;
  printf("%-40s grid %3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma -> %5.0f MAC/cyc/SM\n", name, grid, (double)h[0] / iters,
         (double)h[1] / iters, macs / ((double)h[1] / iters));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<160, 1, -1>("N160 2 acc, no commit", d);
  run<160, 1, 3>("N160 2 acc, commit every 4", d);
  run<160, 1, 7>("N160 2 acc, commit every 8", d);
  run<160, 1, 1>("N160 2 acc, commit every 2", d);
  run<160, 0, -1>("N160 1 acc (dependent), no commit", d);
  run<256, 0, -1>("N256 1 acc, no commit", d);
  run<256, 1, 3>("N256 2 acc, commit every 4", d);
  run<128, 1, -1>("N128 2 acc, no commit", d);
  run<64, 3, -1>("N64 4 acc, no commit", d);
  run<16, 3, -1>("N16 4 acc, no commit", d);
  run<160, 1, 3>("N160 c/4 + 256 thr try_wait spin", d, 1, 256, 0);
  run<160, 1, 3>("N160 c/4 + 32 thr try_wait spin", d, 1, 32, 0);
  run<160, 1, 3>("N160 c/4 + 8 lanes test_wait+sleep200", d, 1, 256, 200);
  run<160, 1, 3>("N160 c/4 RANDOM data", d, 1, 0, 0, 1);
  run<256, 1, 3>("N256 c/4 RANDOM data", d, 1, 0, 0, 1);
  run<160, 1, 3>("N160 c/4 RANDOM data", d, 148, 0, 0, 1);
  run<256, 1, 3>("N256 c/4 RANDOM data", d, 148, 0, 0, 1);
  run<160, 1, 3>("N160 2 acc, commit every 4", d, 148);
  run<256, 1, 3>("N256 2 acc, commit every 4", d, 148);
  return 0;
}
