// Tensor-core (tcgen05 / TMEM) implementation of the NIF MLP wavefront for sm_100a.
//
// One persistent CTA per SM evaluates 128 escaped rays (= 128 TMEM lanes) at a time through ALL layers:
//   * activations live in shared memory as the fp16 A operand, K-major, no-swizzle "core matrix" layout
//     X[k/8][row][k%8] (a plane of 128 rows x 16 B per 8 columns of K), so a row's 8 consecutive
//     features are one 16-byte store and 32 lanes store 512 contiguous bytes (no bank conflicts);
//   * behind the activations sits a static region S = [ones | encoded input]: the "ones" slice turns every
//     layer's bias into one more K-slice of the GEMM (no bias add in the epilogue), and the encoded input
//     makes the skip-concat layer a plain longer K;
//   * weights (+ bias row) are pre-arranged on the host into B-operand images W[k/8][n][k%8] and streamed from L2
//     through a ring of cp.async.bulk (TMA engine, SASS UBLKCP) copies signalled on mbarriers;
//   * each layer is a chain of tcgen05.mma (M=128, N=160 per instruction, K=16, fp16 x fp16 -> fp32 in TMEM);
//     tcgen05.commit releases ring stages and publishes "accumulator ready";
//   * the epilogue (8 warps: TMEM lane quarter = warp%4, column sub-half = warp/4) reads the accumulators with
//     tcgen05.ld, rounds to fp16, applies ReLU on packed halves and writes the next layer's A operand in place.
//
// Layer pipeline. The output columns of a layer are split into two halves N0 = [0,160) and N1 = [160,320); the K range
// into "lo" = the activations produced by the previous layer's N0 plus the static slices, and "hi" = those produced by
// its N1. The MMAs of a layer are issued as four blocks in the order
//       B0 = N0 x K_lo     B1 = N1 x K_lo     B2 = N0 x K_hi  (N0 complete)     B3 = N1 x K_hi  (N1 complete)
// and the accumulators rotate through three 160-column TMEM slots (slot = (2 layer + half) mod 3). That lets the
// epilogue hide entirely behind the tensor pipe:
//       while B3 runs   the epilogue drains N0 into REGISTERS (the A operand is still being read by B3),
//       B3 completes    the registers are stored over the lo columns (10 x 16 B per thread) -> next layer's B0/B1 start,
//       while B0/B1 run the epilogue drains N1 straight into the hi columns               -> next layer's B2/B3 may start.
// In-place update stays safe (nothing is stored into X before all MMAs of the layer have completed), and no MMA
// ever targets a TMEM slot that is still being drained (the third slot).
// Measured limits (B200): every CTA re-streams the 1.16 MB of weights per 128-row tile, 28.4 K cycles per tile =
// 41 B/cycle/SM (0.96 of 6300 B/cycle chip-wide over 148 SMs). The cycles per tile do not move with 74 ... 148 CTAs
// (B200RT_NIF_GRID, profiles/r02_nif_grid.txt) nor with 2.7 % fewer weight bytes: they are ~17 K of MMA issue + 7.4 K at
// the layer hand-offs (TMEM drains at 64 B/cycle/SM: 2560 cycles per layer against 3360 cycles of MMAs, activations
// updated in place) + 2.6-3.8 K of ring refill latency (a bulk copy has ~200-250 cycles of fixed cost: stages below
// 20 KB starve the ring). Variants that hide the epilogue completely (128/192 column split, double-buffered lo columns,
// feature encode on dedicated warps) pay with a shallower ring and end at the same 28 K; the cta_group::2 CTA pair
// (nif_tc_pair2.cuh, experiment build) halves the stream but lengthens every hand-off. scripts/experiments/README.md.
// Roles: warps 0-7 = encode + epilogue, warp 8 lane 0 = weight producer, warp 9 = MMA issuer (whole warp converged,
// one elected lane issues, so descriptors stay in uniform registers).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace rt {
namespace tc {

constexpr int kRows = 128;          // rows (escaped rays) per tile == TMEM lanes
constexpr int kHalfN = 160;         // output columns per MMA instruction / per accumulator slot
// A/B knob: __launch_bounds__ minimum blocks per SM of the kernel = its register cap (1: 96 registers as compiled;
// 3: 64). Fewer registers would leave more warps of the bounce kernels beside the CTA under the chunk overlap.
#ifndef B200RT_NIF_MINBLOCKS
#define B200RT_NIF_MINBLOCKS 1
#endif
#ifndef B200RT_NIF_STAGES
#define B200RT_NIF_STAGES 6
#endif
constexpr int kStages = B200RT_NIF_STAGES;  // weight ring depth
constexpr int kStageK = 64;         // K elements per ring stage (4 MMA K-slices of one block)
constexpr int kMaxLayers = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kPlaneBytes = kRows * 16;  // one K-chunk (8 columns) of the A operand
constexpr int kStaticPlanesMax = 2 + 8;  // ones slice (2 planes) + up to 64 encoded features
constexpr int kStageBytes = (kStageK / 8) * kHalfN * 16;

struct Layer {
  const __half* wimg;   // blocks B0..B3 back to back; block = [K_b/8][n_b][8] fp16 with
                        // K order of B0/B1 = [lo activation rows | bias row + 15 zero rows | feature rows], B2/B3 = [hi rows]
  int actLoSlices;      // K=16 slices read from the lo activation planes (0 for the first layer)
  int actHiSlices;      // ... from the hi activation planes
  int staticSlices;     // slices read from S: 1 (ones) or 1 + F/16 (ones + encoded input)
  int N, Npad;          // Npad = N rounded up to 16, <= 320
  int n0, n1;           // columns of the two halves: n0 = min(Npad, 160), n1 = Npad - n0
  int relu;
  uint32_t pairRow0, pairRankRows;  // experiment build (nif_tc_pair2.cuh): the layer's per-rank images in the 128-B-row view
};

struct Params {
  Layer layers[kMaxLayers];
  int numLayers;
  int embed;            // E, features F = 4E
  int actPlanes;        // planes of the activation buffer (max hidden width / 8)
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
// Drain `n` (multiple of 8) accumulator columns starting at TMEM address `taddr` into packed fp16 pairs:
// dst[i] = columns (2i, 2i+1). The next load is in flight while the current one is packed.
template <int kCols>
__device__ __forceinline__ void drain_to_regs(uint32_t taddr, bool relu, uint32_t (&dst)[kCols / 2]) {
  static_assert(kCols % 8 == 0, "columns per thread must be a multiple of 8");
  constexpr int kFull = kCols / 32, kRem = (kCols % 32) / 8;
  const __half2 zero = __float2half2_rn(0.f);
  auto pack = [&](const uint32_t* acc, int count, int at) {
#pragma unroll
    for (int e = 0; e < count / 2; ++e) {
      __half2 h = __floats2half2_rn(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
      if (relu) h = __hmax2(h, zero);
      dst[at / 2 + e] = *reinterpret_cast<const uint32_t*>(&h);
    }
  };
  uint32_t bufA[32], bufB[32], tail[8];
  if (kFull > 0) tmem_ld32(taddr, bufA);
#pragma unroll
  for (int i = 0; i < kFull; ++i) {
    tmem_ld_wait();
    if (i + 1 < kFull) { if (i & 1) tmem_ld32(taddr + 32u * (i + 1), bufA); else tmem_ld32(taddr + 32u * (i + 1), bufB); }
    pack((i & 1) ? bufB : bufA, 32, 32 * i);
  }
#pragma unroll
  for (int i = 0; i < kRem; ++i) {
    tmem_ld8(taddr + 32u * kFull + 8u * i, tail);
    tmem_ld_wait();
    pack(tail, 8, 32 * kFull + 8 * i);
  }
}

__global__ void __launch_bounds__(kThreads, B200RT_NIF_MINBLOCKS)
nif_mlp_tc_kernel(const Params p, const float* __restrict__ uvDirect, const float* __restrict__ slotEscape,
                  const uint32_t* __restrict__ queue, const uint32_t* __restrict__ dCount, uint32_t directCount,
                  uint32_t first, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [activation planes][static planes: ones, encoded input][ring stages][barriers][tmem ptr]
  unsigned char* X = smem;
  unsigned char* S = X + (size_t)p.actPlanes * kPlaneBytes;
  unsigned char* ring = S + (size_t)kStaticPlanesMax * kPlaneBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kStages * kStageBytes);
  uint64_t* fullBar = bars;                    // [kStages] weights landed
  uint64_t* emptyBar = bars + kStages;         // [kStages] MMAs that read the stage have completed
  uint64_t* actLoBar = bars + 2 * kStages;     // lo columns (+ features) of the next A operand are in place (256 arrivals)
  uint64_t* actHiBar = bars + 2 * kStages + 1; // hi columns are in place (256 arrivals)
  uint64_t* accBar0 = bars + 2 * kStages + 2;  // N0 accumulator of the current layer complete (commit)
  uint64_t* accBar1 = bars + 2 * kStages + 3;  // N1 accumulator complete == every MMA of the layer complete (commit)
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  // warp index made provably warp-uniform (shuffle from lane 0), so the role branches below are uniform branches and
  // the issuer's descriptors live in uniform registers instead of being re-broadcast per MMA
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // queue mode: this launch covers entries [first, first + directCount) of the device-side queue (`queue` already
  // points at entry `first`); the queue's length is only known on the device
  const uint32_t count = uvDirect ? directCount : min(*dCount > first ? *dCount - first : 0u, directCount);
  const uint32_t numTiles = (count + kRows - 1) / kRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(fullBar + s, 1); mbar_init(emptyBar + s, 1); }
    mbar_init(actLoBar, kEpiWarps * 32);
    mbar_init(actHiBar, kEpiWarps * 32);
    mbar_init(accBar0, 1);
    mbar_init(accBar1, 1);
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
    // ===== weight producer: the blocks of every layer in issue order, <= kStageK rows of one block per stage =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long waitEmpty = 0;
      for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int l = 0; l < p.numLayers; ++l) {
          const Layer& L = p.layers[l];
          const unsigned char* src = reinterpret_cast<const unsigned char*>(L.wimg);
          const uint32_t loPlanes = 2u * (uint32_t)(L.actLoSlices + L.staticSlices), hiPlanes = 2u * (uint32_t)L.actHiSlices;
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            const uint32_t planes = b < 2 ? loPlanes : hiPlanes;
            const uint32_t planeBytes = (uint32_t)((b & 1) ? L.n1 : L.n0) * 16u;
            if (planeBytes == 0u) continue;
            for (uint32_t pl = 0; pl < planes; pl += kStageK / 8) {
              const uint32_t bytes = min((uint32_t)(kStageK / 8), planes - pl) * planeBytes;
              { NIF_PROF_T0(); mbar_wait(emptyBar + stage, phase ^ 1u); NIF_PROF_ADD(waitEmpty); }
              mbar_expect_tx(fullBar + stage, bytes);
              bulk_load(ring + (size_t)stage * kStageBytes, src + (size_t)pl * planeBytes, bytes, fullBar + stage);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            src += (size_t)planes * planeBytes;
          }
        }
      }
      if (p.prof) p.prof[(size_t)blockIdx.x * 16 + PF_PROD_WAIT_EMPTY] = waitEmpty;
    }
  } else if (warp > kEpiWarps) {
    // ===== MMA issuer =====
    // The issuing thread is a scalar instruction stream on the critical path of the tensor pipe (one MMA must be
    // issued every <= 80 cycles), so the K loop is kept to a handful of 32-bit adds per MMA: only the low word of a
    // descriptor (start address) changes, everything else is hoisted per block.
    uint32_t stage = 0, phase = 0, actLoPhase = 0, actHiPhase = 0;
    unsigned long long waitAct = 0, waitFull = 0, mmaPhase = 0, tiles = 0;
    const long long tStart = p.prof ? clock64() : 0;
    const uint32_t xAddr = smem_u32(X), sAddr = smem_u32(S), ringAddr = smem_u32(ring);
    const uint32_t descHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, version 1 (bit 46)
    const uint32_t aLoX = ((xAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
    const uint32_t aLoS = ((sAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
    constexpr uint32_t kSlicesPerStage = kStageK / 16;
    constexpr uint32_t aSliceStep = (2u * kPlaneBytes) >> 4;
    constexpr uint32_t stageStep = (uint32_t)kStageBytes >> 4;

    // One block: `slices` K-slices of an n-column half accumulated into TMEM columns [dTmem, dTmem + n). The A operand
    // comes from `aStart` for the first `switchAt` slices and from the static region S afterwards.
    auto run_block = [&](uint32_t dTmem, uint32_t n, uint32_t aStart, uint32_t switchAt, uint32_t slices, bool fresh) {
      const uint32_t idesc = instr_desc(kRows, (int)n);
      const uint32_t bLoBase = ((ringAddr >> 4) & 0x3FFFu) | (n << 16);  // LBO = n * 16 B
      const uint32_t bSliceStep = 2u * n;                                // two planes of n * 16 B, >> 4
      uint32_t aLo = switchAt ? aStart : aLoS;
      uint32_t ks = 0;
      while (ks < slices) {
        { const long long w0 = p.prof ? clock64() : 0; mbar_wait(fullBar + stage, phase); if (p.prof) waitFull += (unsigned long long)(clock64() - w0); }
        tc_fence_after();
        uint32_t bLo = bLoBase + stage * stageStep;
#pragma unroll
        for (uint32_t j = 0; j < kSlicesPerStage; ++j) {
          if (ks < slices) {
            if (ks == switchAt && switchAt) aLo = aLoS;  // activations exhausted: continue with [ones | encoded input]
            mma_f16_lo(dTmem, aLo, bLo, descHi, idesc, (fresh && ks == 0) ? 0u : 1u);
            aLo += aSliceStep;
            bLo += bSliceStep;
            ++ks;
          }
        }
        mma_commit_elect(emptyBar + stage);  // stage reusable once these MMAs have read it
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    };

    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      for (int l = 0; l < p.numLayers; ++l) {
        const uint32_t lo = (uint32_t)p.layers[l].actLoSlices, hi = (uint32_t)p.layers[l].actHiSlices;
        const uint32_t loSlices = lo + (uint32_t)p.layers[l].staticSlices;
        const uint32_t n0 = (uint32_t)p.layers[l].n0, n1 = (uint32_t)p.layers[l].n1;
        const uint32_t d0 = tmemBase + (uint32_t)((2 * l) % 3) * kHalfN, d1 = tmemBase + (uint32_t)((2 * l + 1) % 3) * kHalfN;
        { NIF_PROF_T0(); mbar_wait(actLoBar, actLoPhase); NIF_PROF_ADD(waitAct); }
        actLoPhase ^= 1u;
        tc_fence_after();
        NIF_PROF_T0();
        run_block(d0, n0, aLoX, lo, loSlices, true);
        if (n1) run_block(d1, n1, aLoX, lo, loSlices, true);
        { const long long w0 = p.prof ? clock64() : 0; mbar_wait(actHiBar, actHiPhase); if (p.prof) waitAct += (unsigned long long)(clock64() - w0); }
        actHiPhase ^= 1u;
        tc_fence_after();
        if (hi) run_block(d0, n0, aLoX + lo * aSliceStep, hi, hi, false);
        mma_commit_elect(accBar0);  // N0 of layer l complete
        if (hi && n1) run_block(d1, n1, aLoX + lo * aSliceStep, hi, hi, false);
        mma_commit_elect(accBar1);  // every MMA of layer l complete
        NIF_PROF_ADD(mmaPhase);
      }
      tiles += 1;
    }
    if (p.prof && lane == 0) {
      unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
      r[PF_TOTAL] = (unsigned long long)(clock64() - tStart);
      r[PF_MMA_WAIT_ACT] = waitAct; r[PF_MMA_WAIT_FULL] = waitFull; r[PF_MMA_ISSUE] = mmaPhase; r[PF_TILES] = tiles;
    }
  } else {
    // ===== encode + epilogue: row = (warp % 4) * 32 + lane (TMEM lane), column sub-half = warp / 4 =====
    const int row = (warp & 3) * 32 + lane;
    const int sub = warp >> 2;
    unsigned char* xRow = X + (size_t)row * 16;
    unsigned char* sRow = S + (size_t)row * 16;
    const uint32_t laneTaddr = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t accPhase = 0;
    unsigned long long waitAcc = 0, encodeCyc = 0, drainCyc = 0;
    const int E = p.embed;
    // the last layer whose MMAs read the encoded input: once ITS accumulators are complete the feature planes may be
    // overwritten with the next tile's features, which hides the encode behind the following layers' MMAs
    int lastFeatLayer = 0;
    for (int l = 0; l < p.numLayers; ++l)
      if (p.layers[l].staticSlices > 1) lastFeatLayer = l;

    // Encode (src/neural_networks/NifModel.cpp:186-219): sub 0 does the u features, sub 1 the v features.
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
      const float w = ((sub == 0 ? u : v) - 1.f) * 2.f;
      float c = 1.f;
      for (int j = 0; j < E; ++j, c *= 2.f) {
        const float a = __half2float(__float2half_rn(w * c));
        float sn, cs;
        sincosf(a, &sn, &cs);
        const int fs = sub * E + j, fc = 2 * E + sub * E + j;
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fs >> 3)) * kPlaneBytes)[fs & 7] = __float2half_rn(sn);
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fc >> 3)) * kPlaneBytes)[fc & 7] = __float2half_rn(cs);
      }
      NIF_PROF_ADD(encodeCyc);
      return slot;
    };
    auto release = [&](uint64_t* bar) {  // generic-proxy stores (and TMEM reads) before, async-proxy MMAs after
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar);
    };

    uint32_t slot = 0xFFFFFFFFu, slotNext = 0xFFFFFFFFu;
    if (blockIdx.x < numTiles) {
      slot = encode_tile(blockIdx.x);
      release(actLoBar);
      release(actHiBar);
    }
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      const bool haveNext = tile + gridDim.x < numTiles;
      for (int l = 0; l < p.numLayers; ++l) {
        const Layer& L = p.layers[l];
        const bool last = l == p.numLayers - 1;
        const bool relu = L.relu != 0;
        const uint32_t t0 = laneTaddr + (uint32_t)((2 * l) % 3) * kHalfN, t1 = laneTaddr + (uint32_t)((2 * l + 1) % 3) * kHalfN;
        { NIF_PROF_T0(); mbar_wait(accBar0, accPhase); NIF_PROF_ADD(waitAcc); }
        tc_fence_after();
        if (last) {
          // decode (NifModel.cpp:222-246); the last layer is narrower than one half, so N1 is empty
          NIF_PROF_T0();
          if (sub == 0) {
            uint32_t acc[8];
            tmem_ld8(t0, acc);
            tmem_ld_wait();
            if (slot != 0xFFFFFFFFu) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float y = __uint_as_float(acc[c]);
                if (relu) y = y > 0.f ? y : 0.f;
                y = __half2float(__float2half_rn(y));  // layer outputs are fp16 (NifModel.cpp:313-315)
                const float mean = c == 0 ? p.mean0 : (c == 1 ? p.mean1 : p.mean2);
                y = y * p.maxv + mean;
                if (p.logToneMap) y = expf(y);
                out[3 * (size_t)slot + c] = y;
              }
            }
          }
          NIF_PROF_ADD(drainCyc);
          { NIF_PROF_T0(); mbar_wait(accBar1, accPhase); NIF_PROF_ADD(waitAcc); }
          accPhase ^= 1u;
          if (haveNext) {
            if (lastFeatLayer >= l) slotNext = encode_tile(tile + gridDim.x);
            release(actLoBar);
            release(actHiBar);
          }
          continue;
        }
        // hidden layer, full-width halves (n0 == kHalfN; n1 == kHalfN or 0): this thread owns 80 columns of each half
        uint32_t h0[kHalfN / 4];
        {
          NIF_PROF_T0();
          drain_to_regs<kHalfN / 2>(t0 + (uint32_t)sub * (kHalfN / 2), relu, h0);  // overlaps block B3 of this layer
          NIF_PROF_ADD(drainCyc);
        }
        { NIF_PROF_T0(); mbar_wait(accBar1, accPhase); NIF_PROF_ADD(waitAcc); }
        accPhase ^= 1u;
        tc_fence_after();
        {
          NIF_PROF_T0();
          const int c0 = sub * (kHalfN / 2);
#pragma unroll
          for (int q = 0; q < kHalfN / 16; ++q)
            *reinterpret_cast<uint4*>(xRow + (size_t)((c0 >> 3) + q) * kPlaneBytes) =
                make_uint4(h0[4 * q], h0[4 * q + 1], h0[4 * q + 2], h0[4 * q + 3]);
          release(actLoBar);  // next layer's B0 / B1 may start
          if (L.n1) {
            uint32_t h1[kHalfN / 4];
            drain_to_regs<kHalfN / 2>(t1 + (uint32_t)sub * (kHalfN / 2), relu, h1);  // overlaps B0 / B1 of the next layer
            const int c1 = kHalfN + sub * (kHalfN / 2);
#pragma unroll
            for (int q = 0; q < kHalfN / 16; ++q)
              *reinterpret_cast<uint4*>(xRow + (size_t)((c1 >> 3) + q) * kPlaneBytes) =
                  make_uint4(h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
          }
          release(actHiBar);  // next layer's B2 / B3 may start
          NIF_PROF_ADD(drainCyc);
        }
        if (haveNext && l == lastFeatLayer) slotNext = encode_tile(tile + gridDim.x);
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
