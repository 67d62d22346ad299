// CTA-pair version of the NIF MLP kernel (nif_tc.cuh): two CTAs on one TPC run `tcgen05.mma.cta_group::2`, M = 256
// (128 rows = one tile per CTA), and EACH CTA STREAMS ONLY HALF OF THE B OPERAND (n/2 columns of every weight block):
// the tensor cores of both SMs read both halves. That halves the L2 -> SM weight re-streaming which bounds the
// single-CTA kernel (416 GB per 45.7 M rows at the L2 slice throughput cap), at the same MMA rate per SM
// (scripts/mma2_microbench.cu: M256 N160 K16 = 80 cycles, results verified).
//   * leader (cluster rank 0): its issuer warp issues every MMA of the pair; tcgen05.commit.cta_group::2 with
//     multicast arrives on the empty / accumulator barriers of BOTH CTAs;
//   * each CTA: own weight producer (its half of the ring), own epilogue + encode (its 128 rows in its own TMEM),
//     own activations; the epilogue threads of both CTAs arrive on the LEADER's activation barriers (mapa +
//     mbarrier.arrive.shared::cluster), the peer's otherwise idle issuer warp relays "my half of stage s has landed";
//   * everything else (four-block layer pipeline, TMEM slot rotation, in-place A operand) is the single-CTA kernel's.
#pragma once
#include "nif_tc.cuh"

namespace rt {
namespace tc {

constexpr int kPairStageBytes = (kStageK / 8) * (kHalfN / 2) * 16;  // half of the B columns per CTA
constexpr int kPairStages = 2 * kStages;                              // same ring bytes, twice the look-ahead

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

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
nif_mlp_tc_pair_kernel(const Params p, const float* __restrict__ uvDirect, const float* __restrict__ slotEscape,
                  const uint32_t* __restrict__ queue, const uint32_t* __restrict__ dCount, uint32_t directCount,
                  float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [activation planes][static planes: ones, encoded input][ring stages][barriers][tmem ptr]
  unsigned char* X = smem;
  unsigned char* S = X + (size_t)p.actPlanes * kPlaneBytes;
  unsigned char* ring = S + (size_t)kStaticPlanesMax * kPlaneBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kPairStages * kPairStageBytes);
  uint64_t* fullBar = bars;                    // [kPairStages] this CTA's half of the stage has landed
  uint64_t* emptyBar = bars + kPairStages;     // [kPairStages] the pair's MMAs that read the stage have completed
  uint64_t* peerFull = bars + 2 * kPairStages + 4;  // [kPairStages] (leader) the peer's half has landed
  uint64_t* actLoBar = bars + 2 * kPairStages;     // lo columns (+ features) of the next A operand are in place (256 arrivals)
  uint64_t* actHiBar = bars + 2 * kPairStages + 1; // hi columns are in place (256 arrivals)
  uint64_t* accBar0 = bars + 2 * kPairStages + 2;  // N0 accumulator of the current layer complete (commit)
  uint64_t* accBar1 = bars + 2 * kPairStages + 3;  // N1 accumulator complete == every MMA of the layer complete (commit)
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 3 * kPairStages + 4);

  // warp index made provably warp-uniform (shuffle from lane 0), so the role branches below are uniform branches and
  // the issuer's descriptors live in uniform registers instead of being re-broadcast per MMA
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t count = uvDirect ? directCount : min(*dCount, directCount);
  const uint32_t numTiles = (count + kRows - 1) / kRows;
  // CTA pair: rank 0 (leader) issues the MMAs for both; pair q works on tiles 2g + rank, g = q, q + numPairs, ...
  // Both CTAs run the same number of groups (a tile past the end is all-invalid rows).
  const uint32_t crank = cluster_ctarank();
  const uint32_t pairId = blockIdx.x >> 1, numPairs = gridDim.x >> 1;
  const uint32_t numGroups = (numTiles + 1u) >> 1;
  // shared::cluster addresses of the leader's barriers the peer signals
  const uint32_t actLoRemote = mapa_u32(smem_u32(actLoBar), 0), actHiRemote = mapa_u32(smem_u32(actHiBar), 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) { mbar_init(fullBar + s, 1); mbar_init(emptyBar + s, 1); mbar_init(peerFull + s, 1); }
    mbar_init(actLoBar, 2 * kEpiWarps);  // the epilogue warps of BOTH CTAs arrive on the leader's barrier
    mbar_init(actHiBar, 2 * kEpiWarps);
    mbar_init(accBar0, 1);
    mbar_init(accBar1, 1);
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
          // this CTA's half of the layer's blocks: columns [rank n/2, rank n/2 + n/2) of every block
          const unsigned char* src = reinterpret_cast<const unsigned char*>(L.wimgPair) + (size_t)crank * L.pairRankBytes;
          const uint32_t loPlanes = 2u * (uint32_t)(L.actLoSlices + L.staticSlices), hiPlanes = 2u * (uint32_t)L.actHiSlices;
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            const uint32_t planes = b < 2 ? loPlanes : hiPlanes;
            const uint32_t planeBytes = (uint32_t)((b & 1) ? L.n1 : L.n0) * 8u;  // n/2 columns x 16 B
            if (planeBytes == 0u) continue;
            for (uint32_t pl = 0; pl < planes; pl += kStageK / 8) {
              const uint32_t bytes = min((uint32_t)(kStageK / 8), planes - pl) * planeBytes;
              { NIF_PROF_T0(); mbar_wait(emptyBar + stage, phase ^ 1u); NIF_PROF_ADD(waitEmpty); }
              mbar_expect_tx(fullBar + stage, bytes);
              bulk_load(ring + (size_t)stage * kPairStageBytes, src + (size_t)pl * planeBytes, bytes, fullBar + stage);
              if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
            }
            src += (size_t)planes * planeBytes;
          }
        }
      }
      if (p.prof) p.prof[(size_t)blockIdx.x * 16 + PF_PROD_WAIT_EMPTY] = waitEmpty;
    }
  } else if (warp > kEpiWarps) {
    // ===== MMA issuer (leader CTA) / weight-arrival relay (peer CTA) =====
    if (crank != 0) {
      // The leader's MMAs read this CTA's half of B as well: tell it when each of our ring stages has landed. A remote
      // arrive is a ~500-cycle round trip, so lane s owns ring stage s and the kPairStages notifications overlap.
      uint32_t stagesPerTile = 0;
      for (int l = 0; l < p.numLayers; ++l) {
        const Layer& L = p.layers[l];
        const uint32_t loPlanes = 2u * (uint32_t)(L.actLoSlices + L.staticSlices), hiPlanes = 2u * (uint32_t)L.actHiSlices;
        const uint32_t perLo = (loPlanes + kStageK / 8 - 1) / (kStageK / 8), perHi = (hiPlanes + kStageK / 8 - 1) / (kStageK / 8);
        stagesPerTile += perLo * ((L.n0 ? 1u : 0u) + (L.n1 ? 1u : 0u)) + perHi * ((L.n0 ? 1u : 0u) + (L.n1 ? 1u : 0u));
      }
      uint32_t groupsMine = 0;
      for (uint32_t g = pairId; g < numGroups; g += numPairs) groupsMine++;
      const uint32_t total = groupsMine * stagesPerTile;
      if (lane < kPairStages) {
        const uint32_t remote = mapa_u32(smem_u32(peerFull + lane), 0);
        uint32_t phase = 0;
        for (uint32_t i = (uint32_t)lane; i < total; i += kPairStages) {
          mbar_wait(fullBar + lane, phase);
          mbar_arrive_remote(remote);
          phase ^= 1u;
        }
      }
    } else {
    // The issuing thread is a scalar instruction stream on the critical path of the tensor pipe (one MMA must be
    // issued every <= 80 cycles), so the K loop is kept to a handful of 32-bit adds per MMA: only the low word of a
    // descriptor (start address) changes, everything else is hoisted per block.
    uint32_t stage = 0, phase = 0, actLoPhase = 0, actHiPhase = 0;
    unsigned long long waitAct = 0, waitFull = 0, waitPeer = 0, mmaPhase = 0, tiles = 0;
    const long long tStart = p.prof ? clock64() : 0;
    const uint32_t xAddr = smem_u32(X), sAddr = smem_u32(S), ringAddr = smem_u32(ring);
    const uint32_t descHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, version 1 (bit 46)
    const uint32_t aLoX = ((xAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
    const uint32_t aLoS = ((sAddr >> 4) & 0x3FFFu) | ((uint32_t)(kPlaneBytes >> 4) << 16);
    constexpr uint32_t kSlicesPerStage = kStageK / 16;
    constexpr uint32_t aSliceStep = (2u * kPlaneBytes) >> 4;
    constexpr uint32_t stageStep = (uint32_t)kPairStageBytes >> 4;

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
        { const long long w0 = p.prof ? clock64() : 0; mbar_wait(peerFull + stage, phase); if (p.prof) waitPeer += (unsigned long long)(clock64() - w0); }
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
        if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
      }
    };

    for (uint32_t g = pairId; g < numGroups; g += numPairs) {
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
        mma2_commit_elect(accBar0);  // N0 of layer l complete (signalled in both CTAs)
        if (hi && n1) run_block(d1, n1, aLoX + lo * aSliceStep, hi, hi, false);
        mma2_commit_elect(accBar1);  // every MMA of layer l complete
        NIF_PROF_ADD(mmaPhase);
      }
      tiles += 1;
    }
    if (p.prof && lane == 0) {
      unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
      r[PF_TOTAL] = (unsigned long long)(clock64() - tStart);
      r[PF_MMA_WAIT_ACT] = waitAct; r[PF_MMA_WAIT_FULL] = waitFull; r[PF_MMA_ISSUE] = mmaPhase; r[PF_TILES] = tiles; r[PF_COUNT] = waitPeer;
    }
    }  // leader
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
    auto release = [&](uint32_t leaderBar) {  // generic-proxy stores (and TMEM reads) before, async-proxy MMAs after
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {  // one arrive per warp: the leader's barrier collects the 8 epilogue warps of both CTAs
        if (crank == 0) mbar_arrive(leaderBar == actLoRemote ? actLoBar : actHiBar); else mbar_arrive_remote(leaderBar);
      }
    };

    uint32_t slot = 0xFFFFFFFFu, slotNext = 0xFFFFFFFFu;
    if (pairId < numGroups) {
      slot = encode_tile(2u * pairId + crank);
      release(actLoRemote);
      release(actHiRemote);
    }
    for (uint32_t g = pairId; g < numGroups; g += numPairs) {
      const bool haveNext = g + numPairs < numGroups;
      const uint32_t tileNext = 2u * (g + numPairs) + crank;
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
            if (lastFeatLayer >= l) slotNext = encode_tile(tileNext);
            release(actLoRemote);
            release(actHiRemote);
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
          release(actLoRemote);  // next layer's B0 / B1 may start
          if (L.n1) {
            uint32_t h1[kHalfN / 4];
            drain_to_regs<kHalfN / 2>(t1 + (uint32_t)sub * (kHalfN / 2), relu, h1);  // overlaps B0 / B1 of the next layer
            const int c1 = kHalfN + sub * (kHalfN / 2);
#pragma unroll
            for (int q = 0; q < kHalfN / 16; ++q)
              *reinterpret_cast<uint4*>(xRow + (size_t)((c1 >> 3) + q) * kPlaneBytes) =
                  make_uint4(h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
          }
          release(actHiRemote);  // next layer's B2 / B3 may start
          NIF_PROF_ADD(drainCyc);
        }
        if (haveNext && l == lastFeatLayer) slotNext = encode_tile(tileNext);
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
  cluster_sync();  // neither CTA leaves (or frees TMEM) while the other may still signal it or the leader's MMAs run
  if (warp == 0) tmem_dealloc2(tmemBase, 512);
}

}  // namespace tc
}  // namespace rt
