// Tensor-core (tcgen05 / TMEM) implementation of the NIF MLP wavefront for sm_100a.
//
// One persistent CTA per SM evaluates 128 escaped rays (= 128 TMEM lanes) at a time through ALL layers:
//   * activations live in shared memory as the fp16 A operand, K-major, no-swizzle "core matrix" layout
//     X[k/8][row][k%8] (a plane of 128 rows x 16 B per 8 columns of K), so a row's 8 consecutive
//     features are one 16-byte store and 32 lanes store 512 contiguous bytes (no bank conflicts);
//   * weights are pre-arranged on the host into the matching B-operand image W[k/8][n][k%8] and streamed
//     from L2 through a 3-stage ring of cp.async.bulk (TMA engine, SASS UBLKCP) copies signalled on mbarriers;
//   * each layer is a chain of tcgen05.mma (M=128, N<=256 per instruction, K=16, fp16 x fp16 -> fp32 in
//     TMEM) issued by ONE thread; tcgen05.commit releases ring stages and publishes "accumulator ready";
//   * the epilogue (4 warps, one TMEM lane quarter each) reads the accumulators with tcgen05.ld, adds the
//     bias, applies ReLU, rounds to fp16 and writes the next layer's A operand in place; the encoded
//     input stays parked behind the activations so the skip-concat layer is a plain longer K.
// Roles: warps 0-3 = encode + epilogue (thread t <-> row t), warp 4 lane 0 = weight producer,
// warp 5 lane 0 = MMA issuer.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace rt {
namespace tc {

constexpr int kRows = 128;          // rows (escaped rays) per tile == TMEM lanes
constexpr int kStages = 3;          // weight ring depth
constexpr int kStageK = 64;         // K elements per ring stage (4 MMAs of K=16)
constexpr int kMaxLayers = 16;
constexpr int kThreads = 192;
constexpr int kPlaneBytes = kRows * 16;  // one K-chunk (8 columns) of the A operand

struct Layer {
  const __half* wimg;   // [K/8][Npad][8] fp16
  const float* bias;    // [Npad] fp32
  int K, N, Npad;       // Npad = N rounded up to 16
  int relu;
  int aPlane0;          // first A plane this layer reads
  int copyFeatTo;       // >= 0: before this layer, copy the parked features to this column (concat at odd width)
};

struct Params {
  Layer layers[kMaxLayers];
  int numLayers;
  int embed;            // E, features F = 4E
  int featCol;          // column where the encoded input is parked (multiple of 8)
  int xPlanes;          // planes in the X buffer
  int stageBytes;       // bytes of one ring stage
  int biasFloats;       // total bias floats (all layers, padded)
  float maxv, mean0, mean1, mean2;
  int logToneMap;
  int swapLboSbo;       // debug switch: exchange the descriptor's leading/stride offsets
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dstSmem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dstSmem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(dTmem),
      "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: the mbarrier is arrived on once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4, [16,30) leading-dimension byte offset >> 4 (stride between the two 8-column
//   K chunks of one K=16 MMA), [32,46) stride-dimension byte offset >> 4 (stride between 8-row groups),
//   [46,48) version = 1 on sm_100, [61,64) layout type 0 = no swizzle.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lboBytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sboBytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32, A = B = fp16, both K-major, dense.
__device__ __forceinline__ uint32_t instr_desc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Encode one (u,v) into 4E fp16 features, packed as 8-column chunks (src/neural_networks/NifModel.cpp:186-219).
__device__ __forceinline__ void encode_row(float u, float v, int E, unsigned char* xRow /* X + row*16 */, int featPlane0) {
  const float un = (u - 1.f) * 2.f, vn = (v - 1.f) * 2.f;
  // feature index f = j (sin u), E + j (sin v), 2E + j (cos u), 3E + j (cos v)
  float c = 1.f;
  for (int j = 0; j < E; ++j, c *= 2.f) {
    const float au = __half2float(__float2half_rn(un * c));
    const float av = __half2float(__float2half_rn(vn * c));
    float su, cu, sv, cv;
    sincosf(au, &su, &cu);
    sincosf(av, &sv, &cv);
    const int f[4] = {j, E + j, 2 * E + j, 3 * E + j};
    const float val[4] = {su, sv, cu, cv};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __half* dst = reinterpret_cast<__half*>(xRow + (size_t)(featPlane0 + (f[q] >> 3)) * kPlaneBytes) + (f[q] & 7);
      *dst = __float2half_rn(val[q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
nif_mlp_tc_kernel(const Params p, const float* __restrict__ uvDirect, const float* __restrict__ slotEscape,
                  const uint32_t* __restrict__ queue, const uint32_t* __restrict__ dCount, uint32_t directCount,
                  float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [X planes][ring stages][bias floats][barriers][tmem ptr]
  unsigned char* X = smem;
  unsigned char* ring = X + (size_t)p.xPlanes * kPlaneBytes;
  float* biasS = reinterpret_cast<float*>(ring + (size_t)kStages * p.stageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(biasS + ((p.biasFloats + 1) & ~1));
  uint64_t* fullBar = bars;                 // [kStages] weights landed
  uint64_t* emptyBar = bars + kStages;      // [kStages] MMAs that read the stage have completed
  uint64_t* actBar = bars + 2 * kStages;    // A operand of the next layer is ready (128 arrivals)
  uint64_t* accBar = bars + 2 * kStages + 1;  // accumulator of the current layer is complete (1 arrival via commit)
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t count = uvDirect ? directCount : min(*dCount, directCount);
  const uint32_t numTiles = (count + kRows - 1) / kRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(fullBar + s, 1); mbar_init(emptyBar + s, 1); }
    mbar_init(actBar, kRows);
    mbar_init(accBar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.biasFloats; i += kThreads) {
    // biases of all layers, concatenated in layer order
    int l = 0, off = i;
    while (off >= p.layers[l].Npad) { off -= p.layers[l].Npad; ++l; }
    biasS[i] = p.layers[l].bias[off];
  }
  if (warp == 0) tmem_alloc(tmemPtr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemPtr;

  if (warp == 4) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int l = 0; l < p.numLayers; ++l) {
          const Layer& L = p.layers[l];
          const uint32_t planeBytes = (uint32_t)L.Npad * 16u;
          const uint32_t totalPlanes = (uint32_t)L.K / 8u;
          for (uint32_t pl = 0; pl < totalPlanes; pl += kStageK / 8) {
            const uint32_t planes = min((uint32_t)(kStageK / 8), totalPlanes - pl);
            const uint32_t bytes = planes * planeBytes;
            mbar_wait(emptyBar + stage, phase ^ 1u);
            mbar_expect_tx(fullBar + stage, bytes);
            bulk_load(ring + (size_t)stage * p.stageBytes, reinterpret_cast<const unsigned char*>(L.wimg) + (size_t)pl * planeBytes,
                      bytes, fullBar + stage);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, actPhase = 0;
      const uint32_t xAddr = smem_u32(X), ringAddr = smem_u32(ring);
      for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int l = 0; l < p.numLayers; ++l) {
          const Layer& L = p.layers[l];
          const uint32_t planeBytesB = (uint32_t)L.Npad * 16u;
          mbar_wait(actBar, actPhase);
          actPhase ^= 1u;
          tc_fence_after();
          const uint32_t slices = (uint32_t)L.K / 16u;
          for (uint32_t ks = 0; ks < slices; ++ks) {
            const uint32_t inStage = ks % (kStageK / 16);
            if (inStage == 0) {
              mbar_wait(fullBar + stage, phase);
              tc_fence_after();
            }
            const uint32_t aAddr = xAddr + (uint32_t)(L.aPlane0 + 2 * (int)ks) * kPlaneBytes;
            const uint32_t bAddr = ringAddr + stage * (uint32_t)p.stageBytes + inStage * 2u * planeBytesB;
            const uint64_t aDesc = p.swapLboSbo ? smem_desc(aAddr, 128u, kPlaneBytes) : smem_desc(aAddr, kPlaneBytes, 128u);
            for (int n0 = 0; n0 < L.Npad; n0 += 160) {
              int nc = L.Npad - n0;
              if (nc > 160) nc = 160;  // N per instruction: multiple of 16, at most 256; 320 = 160 + 160
              const uint32_t bA = bAddr + (uint32_t)n0 * 16u;
              const uint64_t bDesc = p.swapLboSbo ? smem_desc(bA, 128u, planeBytesB) : smem_desc(bA, planeBytesB, 128u);
              mma_f16(tmemBase + (uint32_t)n0, aDesc, bDesc, instr_desc(kRows, nc), ks > 0 ? 1u : 0u);
            }
            if (inStage == kStageK / 16 - 1 || ks == slices - 1) {
              mma_commit(emptyBar + stage);  // stage reusable once these MMAs have read it
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
          mma_commit(accBar);  // accumulator of layer l complete
        }
      }
    }
  } else {
    // ===== encode + epilogue: thread t owns row t (TMEM lane t) =====
    const int row = threadIdx.x;
    unsigned char* xRow = X + (size_t)row * 16;
    const uint32_t laneTaddr = tmemBase + ((uint32_t)(warp * 32) << 16);
    uint32_t accPhase = 0;
    const int F = 4 * p.embed;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      const uint32_t r = tile * kRows + (uint32_t)row;
      float u = 0.f, v = 0.f;
      uint32_t slot = 0xFFFFFFFFu;
      if (r < count) {
        if (uvDirect) { slot = r; u = uvDirect[2 * (size_t)r]; v = uvDirect[2 * (size_t)r + 1]; }
        else { slot = queue[r]; u = slotEscape[5 * (size_t)slot + 3]; v = slotEscape[5 * (size_t)slot + 4]; }
      }
      encode_row(u, v, p.embed, xRow, p.featCol / 8);
      fence_proxy_async();
      mbar_arrive(actBar);

      int bOff = 0;
      for (int l = 0; l < p.numLayers; ++l) {
        const Layer& L = p.layers[l];
        mbar_wait(accBar, accPhase);
        accPhase ^= 1u;
        tc_fence_after();
        const bool last = l == p.numLayers - 1;
        if (last) {
          uint32_t acc[32];
          tmem_ld16(laneTaddr, acc);
          tmem_ld_wait();
          if (slot != 0xFFFFFFFFu) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float y = __uint_as_float(acc[c]) + biasS[bOff + c];
              if (L.relu) y = y > 0.f ? y : 0.f;
              y = __half2float(__float2half_rn(y));  // layer outputs are fp16 (NifModel.cpp:313-315)
              const float mean = c == 0 ? p.mean0 : (c == 1 ? p.mean1 : p.mean2);
              y = y * p.maxv + mean;                  // decode (NifModel.cpp:222-246)
              if (p.logToneMap) y = expf(y);
              out[3 * (size_t)slot + c] = y;
            }
          }
          tc_fence_before();
        } else {
          for (int c0 = 0; c0 < L.Npad; c0 += 32) {
            uint32_t acc[32];
            const bool full = c0 + 32 <= L.Npad;
            if (full) tmem_ld32(laneTaddr + (uint32_t)c0, acc); else tmem_ld16(laneTaddr + (uint32_t)c0, acc);
            tmem_ld_wait();
            const int cols = full ? 32 : 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 8 < cols) {
                uint32_t packed[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float y0 = __uint_as_float(acc[q * 8 + 2 * e]) + biasS[bOff + c0 + q * 8 + 2 * e];
                  float y1 = __uint_as_float(acc[q * 8 + 2 * e + 1]) + biasS[bOff + c0 + q * 8 + 2 * e + 1];
                  if (L.relu) { y0 = y0 > 0.f ? y0 : 0.f; y1 = y1 > 0.f ? y1 : 0.f; }
                  const __half2 h = __floats2half2_rn(y0, y1);
                  packed[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(xRow + (size_t)((c0 >> 3) + q) * kPlaneBytes) =
                    make_uint4(packed[0], packed[1], packed[2], packed[3]);
              }
            }
          }
          const Layer& Nx = p.layers[l + 1];
          if (Nx.copyFeatTo >= 0) {
            for (int f = 0; f < F; f += 8)
              *reinterpret_cast<uint4*>(xRow + (size_t)((Nx.copyFeatTo + f) >> 3) * kPlaneBytes) =
                  *reinterpret_cast<const uint4*>(xRow + (size_t)((p.featCol + f) >> 3) * kPlaneBytes);
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(actBar);
        }
        bOff += L.Npad;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmemBase, 512);
}

}  // namespace tc
}  // namespace rt
