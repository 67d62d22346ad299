// Tensor-core (tcgen05 / TMEM) implementation of the NIF MLP wavefront for sm_100a.
//
// One persistent CTA per SM evaluates 128 escaped rays (= 128 TMEM lanes) at a time through ALL layers:
//   * activations live in shared memory as the fp16 A operand, K-major, no-swizzle "core matrix" layout
//     X[k/8][row][k%8] (a plane of 128 rows x 16 B per 8 columns of K), so a row's 8 consecutive
//     features are one 16-byte store and 32 lanes store 512 contiguous bytes (no bank conflicts);
//   * behind the activations sits a static region S = [ones | encoded input]: the "ones" slice turns every
//     layer's bias into one more K-slice of the GEMM (no bias add in the epilogue), and the encoded input
//     makes the skip-concat layer a plain longer K;
//   * weights (+ bias row) are pre-arranged on the host into the matching B-operand image W[k/8][n][k%8] and
//     streamed from L2 through a ring of cp.async.bulk (TMA engine, SASS UBLKCP) copies signalled on mbarriers;
//   * each layer is a chain of tcgen05.mma (M=128, N<=256 per instruction, K=16, fp16 x fp16 -> fp32 in
//     TMEM) issued by ONE thread; tcgen05.commit releases ring stages and publishes "accumulator ready";
//   * the epilogue (8 warps: TMEM lane quarter = warp%4, column half = warp/4) reads the accumulators with
//     tcgen05.ld, rounds to fp16, applies ReLU on packed halves and writes the next layer's A operand in place.
// Roles: warps 0-7 = encode + epilogue, warp 8 lane 0 = weight producer, warps 9-10 lane 0 = MMA issuers (one per
// N-half: a single thread's scalar issue stream costs ~200 cycles per MMA, the pipe needs one every 80).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace rt {
namespace tc {

constexpr int kRows = 128;          // rows (escaped rays) per tile == TMEM lanes
constexpr int kStages = 3;          // weight ring depth
constexpr int kStageK = 64;         // K elements per ring stage (4 MMA K-slices)
constexpr int kMaxLayers = 16;
constexpr int kEpiWarps = 8;
constexpr int kIssuers = 2;          // MMA issuer threads: one per N-half, so two scalar issue streams feed the pipe
constexpr int kThreads = (kEpiWarps + 1 + kIssuers) * 32;
constexpr int kPlaneBytes = kRows * 16;  // one K-chunk (8 columns) of the A operand
constexpr int kStaticPlanesMax = 2 + 8;  // ones slice (2 planes) + up to 64 encoded features

struct Layer {
  const __half* wimg;   // [Keff/8][Npad][8] fp16, K order = [activation rows | bias row + 15 zero rows | feature rows]
  int actSlices;        // K=16 slices read from the activation planes (0 for the first layer)
  int staticSlices;     // slices read from S: 1 (ones) or 1 + F/16 (ones + encoded input)
  int N, Npad;          // Npad = N rounded up to 16
  int relu;
};

struct Params {
  Layer layers[kMaxLayers];
  int numLayers;
  int embed;            // E, features F = 4E
  int actPlanes;        // planes of the activation buffer (max hidden width / 8)
  int stageBytes;       // bytes of one ring stage
  float maxv, mean0, mean1, mean2;
  int logToneMap;
  unsigned long long* prof;  // optional [gridDim.x][16] cycle counters (B200RT_NIF_PROFILE=1), else nullptr
};

// slots of the per-CTA profile record
enum : int { PF_TOTAL = 0, PF_PROD_WAIT_EMPTY, PF_MMA_WAIT_ACT, PF_MMA_WAIT_FULL, PF_MMA_ISSUE, PF_EPI_WAIT_ACC,
             PF_EPI_ENCODE, PF_EPI_DRAIN, PF_TILES, PF_COUNT };
#define NIF_PROF_T0() const long long t0__ = p.prof ? clock64() : 0
#define NIF_PROF_ADD(var) do { if (p.prof) var += (unsigned long long)(clock64() - t0__); } while (0)

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
// Same instruction with both descriptors given as (low word, shared high word); executed by a whole converged
// warp, one elected lane issues (the CUTLASS idiom), so operands stay in uniform registers.
__device__ __forceinline__ void mma_f16_lo(uint32_t dTmem, uint32_t aLo, uint32_t bLo, uint32_t descHi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(dTmem),
      "r"(aLo), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: the mbarrier is arrived on once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {  // whole converged warp calls, one lane commits
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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

// relu + round-to-fp16 of 8 fp32 accumulators -> one 16-byte chunk of the next A operand
__device__ __forceinline__ uint4 pack8(const uint32_t* acc, bool relu) {
  uint32_t w[4];
  const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __half2 h = __floats2half2_rn(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
    if (relu) h = __hmax2(h, zero);
    w[e] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
nif_mlp_tc_kernel(const Params p, const float* __restrict__ uvDirect, const float* __restrict__ slotEscape,
                  const uint32_t* __restrict__ queue, const uint32_t* __restrict__ dCount, uint32_t directCount,
                  float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [activation planes][static planes: ones, encoded input][ring stages][barriers][tmem ptr]
  unsigned char* X = smem;
  unsigned char* S = X + (size_t)p.actPlanes * kPlaneBytes;
  unsigned char* ring = S + (size_t)kStaticPlanesMax * kPlaneBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kStages * p.stageBytes);
  uint64_t* fullBar = bars;                   // [kStages] weights landed
  uint64_t* emptyBar = bars + kStages;        // [kStages] MMAs that read the stage have completed
  uint64_t* actBar = bars + 2 * kStages;      // A operand of the next layer is ready (256 arrivals)
  uint64_t* accBar = bars + 2 * kStages + 1;  // accumulator of the current layer is complete (1 arrival via commit)
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  // warp index made provably warp-uniform (shuffle from lane 0), so the role branches below are uniform branches and
  // the issuer's descriptors live in uniform registers instead of being re-broadcast per MMA
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t count = uvDirect ? directCount : min(*dCount, directCount);
  const uint32_t numTiles = (count + kRows - 1) / kRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(fullBar + s, 1); mbar_init(emptyBar + s, kIssuers); }
    mbar_init(actBar, kEpiWarps * 32);
    mbar_init(accBar, kIssuers);
    fence_barrier_init();
  }
  // the ones slice: column 0 = 1.0, columns 1..15 = 0 (the matching weight rows hold the bias and zeros)
  for (int i = threadIdx.x; i < 2 * kRows * 8; i += kThreads) {
    const int plane = i / (kRows * 8), e = i % 8;
    reinterpret_cast<__half*>(S)[i] = __float2half((plane == 0 && e == 0) ? 1.f : 0.f);
  }
  if (warp == 0) tmem_alloc(tmemPtr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemPtr;

  if (warp == kEpiWarps) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long waitEmpty = 0;
      for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int l = 0; l < p.numLayers; ++l) {
          const Layer& L = p.layers[l];
          const uint32_t planeBytes = (uint32_t)L.Npad * 16u;
          const uint32_t totalPlanes = 2u * (uint32_t)(L.actSlices + L.staticSlices);
          for (uint32_t pl = 0; pl < totalPlanes; pl += kStageK / 8) {
            const uint32_t planes = min((uint32_t)(kStageK / 8), totalPlanes - pl);
            const uint32_t bytes = planes * planeBytes;
            { NIF_PROF_T0(); mbar_wait(emptyBar + stage, phase ^ 1u); NIF_PROF_ADD(waitEmpty); }
            mbar_expect_tx(fullBar + stage, bytes);
            bulk_load(ring + (size_t)stage * p.stageBytes,
                      reinterpret_cast<const unsigned char*>(L.wimg) + (size_t)pl * planeBytes, bytes, fullBar + stage);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (p.prof) p.prof[(size_t)blockIdx.x * 16 + PF_PROD_WAIT_EMPTY] = waitEmpty;
    }
  } else if (warp > kEpiWarps) {
    // ===== MMA issuers: issuer h owns output columns [160 h, 160 h + n_h) =====
    // The issuing thread is a scalar instruction stream on the critical path of the tensor pipe (one MMA must be
    // issued every <= 80 cycles per pipe), so the K loop is kept to a handful of 32-bit adds per MMA: only the low
    // word of a descriptor (start address) changes, everything else is hoisted per layer.
    const uint32_t issuer = (uint32_t)(warp - kEpiWarps - 1);
    {
      uint32_t stage = 0, phase = 0, actPhase = 0;
      unsigned long long waitAct = 0, mmaPhase = 0, tiles = 0;
      const long long tStart = p.prof ? clock64() : 0;
      const uint32_t xAddr = smem_u32(X), sAddr = smem_u32(S), ringAddr = smem_u32(ring);
      const uint32_t descHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, version 1 (bit 46)
      const uint32_t aLoX = ((xAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
      const uint32_t aLoS = ((sAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
      const uint32_t stageStep = (uint32_t)p.stageBytes >> 4;
      constexpr uint32_t kSlicesPerStage = kStageK / 16;
      for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int l = 0; l < p.numLayers; ++l) {
          const uint32_t actSlices = (uint32_t)p.layers[l].actSlices;
          const uint32_t slices = actSlices + (uint32_t)p.layers[l].staticSlices;
          const uint32_t npad = (uint32_t)p.layers[l].Npad;
          // N per instruction: multiple of 16, at most 256; 320 = 160 + 160 (issuer 0 / issuer 1)
          const uint32_t n0 = issuer * 160u;
          const uint32_t nMine = npad > n0 ? (npad - n0 > 160u ? 160u : npad - n0) : 0u;
          const uint32_t idesc = instr_desc(kRows, (int)(nMine ? nMine : 16u));
          const uint32_t dTmem = tmemBase + n0;
          const uint32_t bLoBase = (((ringAddr + n0 * 16u) >> 4) & 0x3FFFu) | (npad << 16);  // LBO = Npad * 16 B
          const uint32_t bSliceStep = (2u * npad * 16u) >> 4;
          constexpr uint32_t aSliceStep = (2u * kPlaneBytes) >> 4;
          { NIF_PROF_T0(); mbar_wait(actBar, actPhase); NIF_PROF_ADD(waitAct); }
          actPhase ^= 1u;
          tc_fence_after();
          NIF_PROF_T0();
          uint32_t aLo = actSlices ? aLoX : aLoS;
          uint32_t ks = 0;
          while (ks < slices) {
            mbar_wait(fullBar + stage, phase);
            tc_fence_after();
            uint32_t bLo = bLoBase + stage * stageStep;
#pragma unroll
            for (uint32_t j = 0; j < kSlicesPerStage; ++j) {
              if (ks < slices) {
                if (ks == actSlices) aLo = aLoS;  // activations exhausted: continue with [ones | encoded input]
                if (nMine) mma_f16_lo(dTmem, aLo, bLo, descHi, idesc, ks > 0 ? 1u : 0u);
                aLo += aSliceStep;
                bLo += bSliceStep;
                ++ks;
              }
            }
            mma_commit_elect(emptyBar + stage);  // stage reusable once these MMAs have read it
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          mma_commit_elect(accBar);  // this issuer's half of the layer-l accumulator is complete
          NIF_PROF_ADD(mmaPhase);
        }
        tiles += 1;
      }
      if (p.prof && issuer == 0 && lane == 0) {
        unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
        r[PF_TOTAL] = (unsigned long long)(clock64() - tStart);
        r[PF_MMA_WAIT_ACT] = waitAct; r[PF_MMA_ISSUE] = mmaPhase; r[PF_TILES] = tiles;
      }
    }
  } else {
    // ===== encode + epilogue: row = (warp % 4) * 32 + lane (TMEM lane), column half = warp / 4 =====
    const int row = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    unsigned char* xRow = X + (size_t)row * 16;
    unsigned char* sRow = S + (size_t)row * 16;
    const uint32_t laneTaddr = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t accPhase = 0;
    unsigned long long waitAcc = 0, encodeCyc = 0, drainCyc = 0;
    const int E = p.embed;
    // the last layer whose MMAs read the encoded input: once ITS accumulator is complete the feature planes may be
    // overwritten with the next tile's features, which hides the encode behind the following layer's MMAs
    int lastFeatLayer = 0;
    for (int l = 0; l < p.numLayers; ++l)
      if (p.layers[l].staticSlices > 1) lastFeatLayer = l;

    // Encode (src/neural_networks/NifModel.cpp:186-219): half 0 does the u features, half 1 the v features.
    // Feature order: [sin u]_E [sin v]_E [cos u]_E [cos v]_E, parked in S after the ones slice. Returns the slot.
    auto encode_tile = [&](uint32_t tile) -> uint32_t {
      NIF_PROF_T0();
      const uint32_t r = tile * kRows + (uint32_t)row;
      float u = 0.f, v = 0.f;
      uint32_t slot = 0xFFFFFFFFu;
      if (r < count) {
        if (uvDirect) { slot = r; u = uvDirect[2 * (size_t)r]; v = uvDirect[2 * (size_t)r + 1]; }
        else { slot = queue[r]; u = slotEscape[5 * (size_t)slot + 3]; v = slotEscape[5 * (size_t)slot + 4]; }
      }
      const float w = ((half == 0 ? u : v) - 1.f) * 2.f;
      float c = 1.f;
      for (int j = 0; j < E; ++j, c *= 2.f) {
        const float a = __half2float(__float2half_rn(w * c));
        float sn, cs;
        sincosf(a, &sn, &cs);
        const int fs = half * E + j, fc = 2 * E + half * E + j;
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fs >> 3)) * kPlaneBytes)[fs & 7] = __float2half_rn(sn);
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fc >> 3)) * kPlaneBytes)[fc & 7] = __float2half_rn(cs);
      }
      NIF_PROF_ADD(encodeCyc);
      return slot;
    };

    uint32_t slot = 0xFFFFFFFFu, slotNext = 0xFFFFFFFFu;
    if (blockIdx.x < numTiles) {
      slot = encode_tile(blockIdx.x);
      fence_proxy_async();
      mbar_arrive(actBar);
    }
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      const bool haveNext = tile + gridDim.x < numTiles;
      for (int l = 0; l < p.numLayers; ++l) {
        const Layer& L = p.layers[l];
        { NIF_PROF_T0(); mbar_wait(accBar, accPhase); NIF_PROF_ADD(waitAcc); }
        accPhase ^= 1u;
        tc_fence_after();
        NIF_PROF_T0();
        const bool last = l == p.numLayers - 1;
        if (last) {
          if (half == 0) {
            uint32_t acc[8];
            tmem_ld8(laneTaddr, acc);
            tmem_ld_wait();
            if (slot != 0xFFFFFFFFu) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float y = __uint_as_float(acc[c]);
                if (L.relu) y = y > 0.f ? y : 0.f;
                y = __half2float(__float2half_rn(y));  // layer outputs are fp16 (NifModel.cpp:313-315)
                const float mean = c == 0 ? p.mean0 : (c == 1 ? p.mean1 : p.mean2);
                y = y * p.maxv + mean;                  // decode (NifModel.cpp:222-246)
                if (p.logToneMap) y = expf(y);
                out[3 * (size_t)slot + c] = y;
              }
            }
          }
        } else {
          // this thread's columns: [c0, c1); the two halves split Npad when it is a multiple of 64
          const bool split = (L.Npad & 63) == 0;
          const int c0 = split ? half * (L.Npad >> 1) : 0;
          const int c1 = split ? c0 + (L.Npad >> 1) : (half == 0 ? L.Npad : 0);
          const bool relu = L.relu != 0;
          // software-pipelined drain: the TMEM load of the next 32 columns is in flight while the current 32 are
          // rounded to fp16 and stored
          uint32_t bufA[32], bufB[32];
          int c = c0;
          if (c + 32 <= c1) tmem_ld32(laneTaddr + (uint32_t)c, bufA);
          while (c + 32 <= c1) {
            tmem_ld_wait();
            const bool moreB = c + 64 <= c1;
            if (moreB) tmem_ld32(laneTaddr + (uint32_t)(c + 32), bufB);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(xRow + (size_t)((c >> 3) + q) * kPlaneBytes) = pack8(bufA + 8 * q, relu);
            c += 32;
            if (moreB) {
              tmem_ld_wait();
              if (c + 64 <= c1) tmem_ld32(laneTaddr + (uint32_t)(c + 32), bufA);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(xRow + (size_t)((c >> 3) + q) * kPlaneBytes) = pack8(bufB + 8 * q, relu);
              c += 32;
            }
          }
          for (; c < c1; c += 8) {
            uint32_t acc[8];
            tmem_ld8(laneTaddr + (uint32_t)c, acc);
            tmem_ld_wait();
            *reinterpret_cast<uint4*>(xRow + (size_t)(c >> 3) * kPlaneBytes) = pack8(acc, relu);
          }
        }
        NIF_PROF_ADD(drainCyc);
        // release the next layer (or the next tile's first layer); before the LAST release of a tile the next
        // tile's features must already be in place
        if (last && haveNext && lastFeatLayer >= l) slotNext = encode_tile(tile + gridDim.x);
        if (!last || haveNext) {
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(actBar);
        }
        if (!last && haveNext && l == lastFeatLayer) slotNext = encode_tile(tile + gridDim.x);
      }
      slot = slotNext;
    }
    if (p.prof && threadIdx.x == 0) {
      unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
      r[PF_EPI_WAIT_ACC] = waitAcc; r[PF_EPI_ENCODE] = encodeCyc; r[PF_EPI_DRAIN] = drainCyc;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmemBase, 512);
}

}  // namespace tc
}  // namespace rt
