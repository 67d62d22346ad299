// CTA-pair NIF kernel with the fully overlapped layer pipeline: tcgen05.mma.cta_group::2 (M = 256 over two SMs, each SM
// streams half of the B operand -> half the L2 weight traffic per SM) combined with the pipeline in which no hand-off
// sits on the tensor pipe's critical path:
//   * output columns split N0 = 128 / N1 = 192 so that both TMEM drains (64 B/cycle) fit behind MMAs,
//   * lo columns of the A operand double-buffered by layer parity: N0 is drained into the buffer the NEXT layer reads
//     and that layer's first blocks are released while block B3 of the current layer still runs,
//   * Fourier features of the next tile encoded by dedicated warps.
// The hand-offs that cross the pair: epilogue warps of both CTAs -> leader's activation barriers (remote arrive),
// leader's commits -> both CTAs (multicast), and the weight arrivals: BOTH producers load their half of a ring stage
// with a TMA tensor copy that counts its bytes on the LEADER's barrier (cp.async.bulk.tensor .cta_group::2), so the
// leader waits on one local barrier per stage and nothing is relayed.
// EXPERIMENT BUILD ONLY (-DB200RT_NIF_PAIR_KERNEL). Shares the PTX wrappers of nif_tc.cuh / nif_tc_pair.cuh; has its
// own tiling constants and weight images (tc2::).
#pragma once
#include "nif_tc_pair.cuh"

namespace rt {
namespace tc2 {
using namespace tc;

constexpr int kRows = 128;          // rows (escaped rays) per tile == TMEM lanes
// Output-column split of a hidden layer: drain(N0) = 8 N0 cycles must fit behind block B3 ((N1/16) MMAs of N1/2 cycles),
// drain(N1) = 8 N1 cycles behind the next layer's B0 + B1 ((N0/16 + 1) slices of 160 cycles): 128 / 192.
constexpr int kN0 = 128, kN1 = 192;
constexpr int kStageK = 48;         // K elements per ring stage (3 MMA K-slices of one block)
constexpr int kStages = 10;         // weight ring depth (half-width stages: each CTA holds n/2 columns of B)
constexpr int kEpiWarps = 8;
constexpr int kEncWarps = 8;        // Fourier-feature encoders for the NEXT tile
constexpr int kThreads = (kEpiWarps + 2 + kEncWarps) * 32;
constexpr int kPlaneBytes = kRows * 16;
constexpr int kStaticPlanesMax = 2 + 8;
constexpr int kLoPlanes = kN0 / 8, kHiPlanes = kN1 / 8;
constexpr int kActPlanes = 2 * kLoPlanes + kHiPlanes;  // lo columns double-buffered by layer parity + hi columns
constexpr int kStageBytes = (kStageK / 8) * (kN1 / 2) * 16;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}

// Drain kCols (64 or 32) accumulator columns starting at TMEM address `taddr` into packed fp16 pairs:
// dst[i] = columns (2i, 2i+1), ReLU applied on the packed halves.
template <int kCols>
__device__ __forceinline__ void drain(uint32_t taddr, bool relu, uint32_t (&dst)[kCols / 2]) {
  static_assert(kCols == 64 || kCols == 32, "one tcgen05.ld per call");
  uint32_t raw[kCols];
  if (kCols == 64) tmem_ld64(taddr, raw); else tmem_ld32(taddr, reinterpret_cast<uint32_t (&)[32]>(raw));
  tmem_ld_wait();
  const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
  for (int e = 0; e < kCols / 2; ++e) {
    __half2 h = __floats2half2_rn(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1]));
    if (relu) h = __hmax2(h, zero);
    dst[e] = *reinterpret_cast<const uint32_t*>(&h);
  }
}
// ... and straight into the A operand: columns [col, col + kCols) of this thread's row
template <int kCols>
__device__ __forceinline__ void drain_to_x(uint32_t taddr, bool relu, unsigned char* xRow, int col) {
  uint32_t h[kCols / 2];
  drain<kCols>(taddr, relu, h);
#pragma unroll
  for (int q = 0; q < kCols / 8; ++q)
    *reinterpret_cast<uint4*>(xRow + (size_t)((col >> 3) + q) * kPlaneBytes) =
        make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
}
static_assert(kN0 == 128 && kN1 == 192, "the epilogue splits N0 into 2 x 64 and N1 into 2 x (64 + 32) columns");

// The pair images of all layers live in ONE buffer viewed as rows of 128 B; a ring stage is `planes` (2, 4 or 6) planes of
// planeBytes (128 B: output layer, 1024 B: N0 halves, 1536 B: N1 halves) = a box of planes x planeBytes / 128 rows. One
// tensor map per box height, indexed [plane size class][planes / 2 - 1].
struct PairMaps {
  CUtensorMap m[9];
};
__host__ __device__ inline int pair_map_index(uint32_t planeBytes, uint32_t planes) {
  return (planeBytes == 128u ? 0 : (planeBytes == 1024u ? 1 : 2)) * 3 + (int)(planes / 2u) - 1;
}

__global__ void __launch_bounds__(kThreads, 1)
nif_mlp_tc_pair2_kernel(const Params p, const __grid_constant__ PairMaps maps, const float* __restrict__ uvDirect,
                  const float* __restrict__ slotEscape, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ dCount,
                  uint32_t directCount, uint32_t first, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [activation planes: lo x 2, hi][static planes: ones, encoded input][ring stages][barriers][tmem ptr]
  unsigned char* X = smem;                                   // [lo buffer 0][lo buffer 1][hi]
  unsigned char* S = X + (size_t)kActPlanes * kPlaneBytes;
  unsigned char* ring = S + (size_t)kStaticPlanesMax * kPlaneBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kStages * kStageBytes);
  uint64_t* fullBar = bars;                    // [kStages] (leader's) BOTH halves of the stage have landed
  uint64_t* emptyBar = bars + kStages;         // [kStages] MMAs that read the stage have completed
  uint64_t* actLoBar = bars + 2 * kStages;     // lo columns (+ features) of the next A operand are in place (256 arrivals)
  uint64_t* actHiBar = bars + 2 * kStages + 1; // hi columns are in place (256 arrivals)
  uint64_t* accBar0 = bars + 2 * kStages + 2;  // N0 accumulator of the current layer complete (commit)
  uint64_t* accBar1 = bars + 2 * kStages + 3;  // N1 accumulator complete == every MMA of the layer complete (commit)
  uint64_t* featFreeBar = bars + 2 * kStages + 4;   // the feature planes may be overwritten (256 epilogue arrivals per tile)
  uint64_t* featReadyBar = bars + 2 * kStages + 5;  // the next tile's features are in place (256 encoder arrivals per tile)
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 6);

  // warp index made provably warp-uniform (shuffle from lane 0), so the role branches below are uniform branches and
  // the issuer's descriptors live in uniform registers instead of being re-broadcast per MMA
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t count = uvDirect ? directCount : min(*dCount > first ? *dCount - first : 0u, directCount);
  const uint32_t numTiles = (count + kRows - 1) / kRows;
  // CTA pair: rank 0 (leader) issues the MMAs for both; pair q works on tiles 2g + rank, g = q, q + numPairs, ...
  const uint32_t crank = cluster_ctarank();
  const uint32_t pairId = blockIdx.x >> 1, numPairs = gridDim.x >> 1;
  const uint32_t numGroups = (numTiles + 1u) >> 1;
  const uint32_t actLoRemote = mapa_u32(smem_u32(actLoBar), 0), actHiRemote = mapa_u32(smem_u32(actHiBar), 0);
  const uint32_t featReadyRemote = mapa_u32(smem_u32(featReadyBar), 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(fullBar + s, 1); mbar_init(emptyBar + s, 1); }
    mbar_init(actLoBar, 2 * kEpiWarps);  // one arrive per epilogue warp of BOTH CTAs, on the leader's barrier
    mbar_init(actHiBar, 2 * kEpiWarps);
    mbar_init(accBar0, 1);
    mbar_init(accBar1, 1);
    mbar_init(featFreeBar, kEpiWarps * 32);
    mbar_init(featReadyBar, 2 * kEncWarps);  // one arrive per encoder warp of both CTAs, on the leader's barrier
    fence_barrier_init();
  }
  // the ones slice: column 0 = 1.0, columns 1..15 = 0 (the matching weight rows hold the bias and zeros)
  for (int i = threadIdx.x; i < 2 * kRows * 8; i += kThreads) {
    const int plane = i / (kRows * 8), e = i % 8;
    reinterpret_cast<__half*>(S)[i] = __float2half((plane == 0 && e == 0) ? 1.f : 0.f);
  }
  if (warp == 0) tmem_alloc2(tmemPtr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers and TMEM exist before either signals the other
  tc_fence_after();
  const uint32_t tmemBase = *tmemPtr;

  if (warp == kEpiWarps) {
    // ===== weight producer: the blocks of every layer in issue order, <= kStageK rows of one block per stage =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long waitEmpty = 0;
      for (uint32_t g = pairId; g < numGroups; g += numPairs) {
        for (int l = 0; l < p.numLayers; ++l) {
          const Layer& L = p.layers[l];
          // this CTA's image of the layer starts at row pairRow0 + rank * pairRankRows of the 128-byte-row view
          uint32_t row = L.pairRow0 + crank * L.pairRankRows;
          const uint32_t loPlanes = 2u * (uint32_t)(L.actLoSlices + L.staticSlices), hiPlanes = 2u * (uint32_t)L.actHiSlices;
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            const uint32_t planes = b < 2 ? loPlanes : hiPlanes;
            const uint32_t planeBytes = (uint32_t)((b & 1) ? L.n1 : L.n0) * 8u;  // n/2 columns x 16 B
            if (planeBytes == 0u) continue;
            for (uint32_t pl = 0; pl < planes; pl += kStageK / 8) {
              const uint32_t np = min((uint32_t)(kStageK / 8), planes - pl);
              const uint32_t bytes = np * planeBytes;
              { NIF_PROF_T0(); mbar_wait(emptyBar + stage, phase ^ 1u); NIF_PROF_ADD(waitEmpty); }
              // the leader's barrier counts the bytes of both halves; the peer's copy signals it directly
              if (crank == 0) mbar_expect_tx(fullBar + stage, 2u * bytes);
              tma_load_2d_pair(ring + (size_t)stage * kStageBytes, &maps.m[pair_map_index(planeBytes, np)], 0,
                               (int)(row + pl * (planeBytes >> 7)), fullBar + stage);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            row += planes * (planeBytes >> 7);
          }
        }
      }
      if (p.prof) p.prof[(size_t)blockIdx.x * 16 + PF_PROD_WAIT_EMPTY] = waitEmpty;
    }
  } else if (warp > kEpiWarps + 1) {
    // ===== encoders (src/neural_networks/NifModel.cpp:186-219): row = (w % 4) * 32 + lane, w / 4 picks u or v =====
    // Feature order: [sin u]_E [sin v]_E [cos u]_E [cos v]_E, parked in S after the ones slice.
    const int w = warp - (kEpiWarps + 2);
    const int row = (w & 3) * 32 + lane;
    const int sub = w >> 2;
    unsigned char* sRow = S + (size_t)row * 16;
    const int E = p.embed;
    uint32_t freePhase = 0;
    unsigned long long encodeCyc = 0;
    for (uint32_t g = pairId; g < numGroups; g += numPairs) {
      if (g != pairId) { mbar_wait(featFreeBar, freePhase); freePhase ^= 1u; }
      NIF_PROF_T0();
      const uint32_t r = (2u * g + crank) * kRows + (uint32_t)row;
      float u = 0.f, v = 0.f;
      if (r < count) {
        if (uvDirect) { u = uvDirect[2 * (size_t)r]; v = uvDirect[2 * (size_t)r + 1]; }
        else { const uint32_t slot = queue[r]; u = slotEscape[5 * (size_t)slot + 3]; v = slotEscape[5 * (size_t)slot + 4]; }
      }
      const float x = ((sub == 0 ? u : v) - 1.f) * 2.f;
      float c = 1.f;
      for (int j = 0; j < E; ++j, c *= 2.f) {
        const float a = __half2float(__float2half_rn(x * c));
        float sn, cs;
        sincosf(a, &sn, &cs);
        const int fs = sub * E + j, fc = 2 * E + sub * E + j;
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fs >> 3)) * kPlaneBytes)[fs & 7] = __float2half_rn(sn);
        reinterpret_cast<__half*>(sRow + (size_t)(2 + (fc >> 3)) * kPlaneBytes)[fc & 7] = __float2half_rn(cs);
      }
      NIF_PROF_ADD(encodeCyc);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { if (crank == 0) mbar_arrive(featReadyBar); else mbar_arrive_remote(featReadyRemote); }
    }
    if (p.prof && w == 0 && lane == 0) p.prof[(size_t)blockIdx.x * 16 + PF_EPI_ENCODE] = encodeCyc;
  } else if (warp > kEpiWarps) {
    // ===== MMA issuer (leader CTA; the peer's warp idles) =====
    if (crank == 0) {
    // The issuing thread is a scalar instruction stream on the critical path of the tensor pipe (one MMA must be
    // issued every <= 80 cycles), so the K loop is kept to a handful of 32-bit adds per MMA: only the low word of a
    // descriptor (start address) changes, everything else is hoisted per block.
    uint32_t stage = 0, phase = 0, actLoPhase = 0, actHiPhase = 0, featPhase = 0;
    unsigned long long waitAct = 0, waitFull = 0, mmaPhase = 0, tiles = 0;
    const long long tStart = p.prof ? clock64() : 0;
    const uint32_t xAddr = smem_u32(X), sAddr = smem_u32(S), ringAddr = smem_u32(ring);
    const uint32_t descHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, version 1 (bit 46)
    const uint32_t aLoX = ((xAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);       // lo buffer 0
    constexpr uint32_t kLoBufStep = (uint32_t)(kLoPlanes * kPlaneBytes) >> 4;                   // lo buffer 1 = + this
    const uint32_t aHiX = aLoX + 2u * kLoBufStep;                                                // hi columns
    const uint32_t aLoS = ((sAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
    constexpr uint32_t kSlicesPerStage = kStageK / 16;
    constexpr uint32_t aSliceStep = (2u * kPlaneBytes) >> 4;
    constexpr uint32_t stageStep = (uint32_t)kStageBytes >> 4;

    // One block: `slices` K-slices of an n-column half accumulated into TMEM columns [dTmem, dTmem + n). The A operand
    // comes from `aStart` for the first `switchAt` slices and from the static region S afterwards.
    auto run_block = [&](uint32_t dTmem, uint32_t n, uint32_t aStart, uint32_t switchAt, uint32_t slices, bool fresh) {
      const uint32_t idesc = instr_desc(2 * kRows, (int)n);  // M = 256: 128 rows in each CTA of the pair
      const uint32_t bLoBase = ((ringAddr >> 4) & 0x3FFFu) | ((n >> 1) << 16);  // each CTA holds n/2 columns: LBO = n/2 * 16 B
      const uint32_t bSliceStep = n;                                            // two planes of n/2 * 16 B, >> 4
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
            mma2_f16_lo(dTmem, aLo, bLo, descHi, idesc, (fresh && ks == 0) ? 0u : 1u);
            aLo += aSliceStep;
            bLo += bSliceStep;
            ++ks;
          }
        }
        mma2_commit_elect(emptyBar + stage);  // stage reusable (in both CTAs) once these MMAs have read it
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    };

    for (uint32_t g = pairId; g < numGroups; g += numPairs) {
      for (int l = 0; l < p.numLayers; ++l) {
        const uint32_t lo = (uint32_t)p.layers[l].actLoSlices, hi = (uint32_t)p.layers[l].actHiSlices;
        const uint32_t loSlices = lo + (uint32_t)p.layers[l].staticSlices;
        const uint32_t n0 = (uint32_t)p.layers[l].n0, n1 = (uint32_t)p.layers[l].n1;
        const uint32_t d0 = tmemBase, d1 = tmemBase + (uint32_t)(kN0 + (l & 1) * kN1);
        if (l == 0) { NIF_PROF_T0(); mbar_wait(featReadyBar, featPhase); NIF_PROF_ADD(waitAct); featPhase ^= 1u; }
        { NIF_PROF_T0(); mbar_wait(actLoBar, actLoPhase); NIF_PROF_ADD(waitAct); }
        actLoPhase ^= 1u;
        tc_fence_after();
        NIF_PROF_T0();
        const uint32_t aLoCur = aLoX + (uint32_t)(l & 1) * kLoBufStep;  // layer l reads the lo buffer of its parity
        run_block(d0, n0, aLoCur, lo, loSlices, true);
        if (n1) run_block(d1, n1, aLoCur, lo, loSlices, true);
        { const long long w0 = p.prof ? clock64() : 0; mbar_wait(actHiBar, actHiPhase); if (p.prof) waitAct += (unsigned long long)(clock64() - w0); }
        actHiPhase ^= 1u;
        tc_fence_after();
        if (hi) run_block(d0, n0, aHiX, hi, hi, false);
        mma2_commit_elect(accBar0);  // N0 of layer l complete (signalled in both CTAs)
        if (hi && n1) run_block(d1, n1, aHiX, hi, hi, false);
        mma2_commit_elect(accBar1);  // every MMA of layer l complete
        NIF_PROF_ADD(mmaPhase);
      }
      tiles += 1;
    }
    if (p.prof && lane == 0) {
      unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
      r[PF_TOTAL] = (unsigned long long)(clock64() - tStart);
      r[PF_MMA_WAIT_ACT] = waitAct; r[PF_MMA_WAIT_FULL] = waitFull; r[PF_MMA_ISSUE] = mmaPhase; r[PF_TILES] = tiles;
    }
    }  // leader
  } else {
    // ===== epilogue: row = (warp % 4) * 32 + lane (TMEM lane), column sub-half = warp / 4 =====
    const int row = (warp & 3) * 32 + lane;
    const int sub = warp >> 2;
    unsigned char* xRow = X + (size_t)row * 16;
    const uint32_t laneTaddr = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t accPhase = 0;
    unsigned long long waitAcc = 0, drainCyc = 0;
    // the last layer whose MMAs read the encoded input: once ITS accumulators are complete the feature planes may be
    // overwritten with the next tile's features, which hides the encode behind the following layers' MMAs
    int lastFeatLayer = 0;
    for (int l = 0; l < p.numLayers; ++l)
      if (p.layers[l].staticSlices > 1) lastFeatLayer = l;

    auto release = [&](uint32_t leaderBar) {  // generic-proxy stores (and TMEM reads) before, async-proxy MMAs after
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {  // one arrive per warp: the leader's barrier collects the 8 epilogue warps of both CTAs
        if (crank == 0) mbar_arrive(leaderBar == actLoRemote ? actLoBar : actHiBar); else mbar_arrive_remote(leaderBar);
      }
    };

    if (pairId < numGroups) {  // nothing to wait for before the first layer of the first tile
      release(actLoRemote);
      release(actHiRemote);
    }
    for (uint32_t g = pairId; g < numGroups; g += numPairs) {
      const bool haveNext = g + numPairs < numGroups;
      const uint32_t tile = 2u * g + crank;
      for (int l = 0; l < p.numLayers; ++l) {
        const Layer& L = p.layers[l];
        const bool last = l == p.numLayers - 1;
        const bool relu = L.relu != 0;
        const uint32_t t0 = laneTaddr, t1 = laneTaddr + (uint32_t)(kN0 + (l & 1) * kN1);
        { NIF_PROF_T0(); mbar_wait(accBar0, accPhase); NIF_PROF_ADD(waitAcc); }
        tc_fence_after();
        if (last) {
          // decode (NifModel.cpp:222-246); the last layer is narrower than one half, so N1 is empty
          NIF_PROF_T0();
          if (sub == 0) {
            uint32_t acc[8];
            tmem_ld8(t0, acc);
            tmem_ld_wait();
            const uint32_t r = tile * kRows + (uint32_t)row;
            const uint32_t slot = r < count ? (uvDirect ? r : queue[r]) : 0xFFFFFFFFu;
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
            if (lastFeatLayer >= l) mbar_arrive(featFreeBar);
            release(actLoRemote);
            release(actHiRemote);
          }
          continue;
        }
        // hidden layer of full width (n0 == kN0, n1 == kN1): this thread owns half of the columns of each part
        {
          NIF_PROF_T0();
          // N0 -> the lo buffer the NEXT layer reads (not the one this layer's MMAs are still reading); overlaps B3
          unsigned char* xLoNext = xRow + (size_t)(((l + 1) & 1) * kLoPlanes) * kPlaneBytes;
          drain_to_x<64>(t0 + (uint32_t)sub * (kN0 / 2), relu, xLoNext, sub * (kN0 / 2));
          release(actLoRemote);  // next layer's B0 / B1 may queue up behind B3
          NIF_PROF_ADD(drainCyc);
        }
        { NIF_PROF_T0(); mbar_wait(accBar1, accPhase); NIF_PROF_ADD(waitAcc); }
        accPhase ^= 1u;
        tc_fence_after();
        {
          NIF_PROF_T0();
          // N1 -> the hi columns in place: every MMA of this layer has completed; overlaps B0 / B1 of the next layer
          unsigned char* xHi = xRow + (size_t)(2 * kLoPlanes) * kPlaneBytes;
          const uint32_t th = t1 + (uint32_t)sub * (kN1 / 2);
          const int c1 = sub * (kN1 / 2);
          drain_to_x<64>(th, relu, xHi, c1);
          drain_to_x<32>(th + 64u, relu, xHi, c1 + 64);
          release(actHiRemote);  // next layer's B2 / B3 may start
          NIF_PROF_ADD(drainCyc);
        }
        if (haveNext && l == lastFeatLayer) mbar_arrive(featFreeBar);  // every MMA that reads the features is complete
      }
    }
    if (p.prof && threadIdx.x == 0) {
      unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
      r[PF_EPI_WAIT_ACC] = waitAcc; r[PF_EPI_DRAIN] = drainCyc;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA leaves (or frees TMEM) while the other may still signal it or the leader's MMAs run
  if (warp == 0) tmem_dealloc2(tmemBase, 512);
}

}  // namespace tc2
}  // namespace rt
