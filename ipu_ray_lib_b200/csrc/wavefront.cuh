// Wavefront path tracer: rays stream through HBM as float4 SoA records and are regrouped between bounces.
//
// The megakernel (path_trace_kernel) runs one BVH query per lane in lockstep, so a warp works at the pace of its
// slowest traversal: with ~10 inner-node steps on average and a long tail, ncu measured ~8 of 32 lanes active in the
// traversal loop. Here the per-path state lives in HBM between bounces and three kernels split the work:
//
//   (bounce 0)   no generate kernel: the first trace and shade kernels compute the camera ray of path p = (pixel, sample)
//                themselves (per-(pixel,sample) RNG stream, jitter, first offsetRay); slot = path
//   wf_trace     persistent warps; every lane takes the next live path's slot ON ITS OWN the moment its traversal ends,
//                so lanes never wait for a neighbour's long traversal; inside, each warp iteration runs one phase
//                chosen by ballot (inner-node steps while enough lanes want one, else leaf tests, else the fetch of
//                new rays)
//   wf_shade     one thread per live path: updateHit, material, BxDF sample, roulette; survivors are compacted: their
//                records (next ray, offset applied, query constants prepared) are appended back to back to the other
//                state array with a block-aggregated append; finished paths write their throughput / escape record
//                (and the HitRecord if they are the last sample)
//
// The arithmetic per path is the megakernel's, statement for statement, so results are bit-identical; only the
// grouping of work changes. rgb is accumulated afterwards in sample order by wf_accumulate_kernel.
#pragma once
#include <cooperative_groups.h>
#include "trace_kernels.cuh"

namespace rt {

// Path state. It is indexed by SLOT, not by path: bounce b reads the records of its live paths back to back from
// st[b & 1] (slot i = the i-th survivor of the previous bounce; at bounce 0 slot = path and nothing is read, the camera
// path is computed) and writes the records of its own survivors back to back into st[(b & 1) ^ 1]. The two arrays ARE
// the queues of the wavefront: every load and store of a record is a dense, fully coalesced 16-byte access, where a
// record array indexed by path id is read through half-empty sectors from the second bounce on.
// A record is 96 B (+ 16 B normal for paths of the call's last sample, whose HitRecord is what the reference leaves in
// the ray stream) plus an 8 B hit (+ 16 B barycentrics when the scene interpolates normals). The running colour lives
// in slotColor (the array the accumulate kernel reads anyway) and is only touched on emissive hits.
struct WfState {
  float4* rayO;  // origin.xyz (offset applied), w = bounce | flags << 8 | geomID << 16
  float4* rayD;  // direction.xyz, w = primID bits
  float4* rayI;  // 1 / direction: the shading kernel prepares the next query's constants (stream_prepare)
  float4* rayS;  // RayShearParams sx, sy, sz, w = stream_prepare's flags
  float4* thr;   // throughput.xyz, w = path id p = pixel * chunk + sample (bits)
  uint4* rng;    // xoroshiro state {s0.lo, s0.hi, s1.lo, s1.hi}
  float4* nrm;   // normal.xyz; read and written only for paths of the last sample
};
struct WfBuffers {
  WfState st[2];
  float2* hitA;      // [slot] closest t, leaf reference of the winner (kRefNone = no hit)
  float4* hitB;      // [slot] b0, b1, b2; only when the scene interpolates normals (else null)
  uint32_t* counts;  // [0], [1]: live paths in st[0] / st[1]; [2]: fetch cursor of wf_trace; [3]: round cursor of wf_shade
};

struct WfArgs {
  TraceArgs t;        // scene, rays, sample range, camera, NIF slot arrays ...
  WfBuffers b;
  uint32_t numPaths;  // numRays * chunk
  uint32_t chunk;
  int chunkShift;     // log2(chunk) when chunk is a power of two (the default chunks are), else -1
  uint32_t lastSample;  // the sample whose HitRecord is left in the ray stream (last of the whole call)
  int qIn;            // state array read by this launch (bounce & 1)
  unsigned long long* phaseStats;  // optional [bounce][3][2]: warp iterations and participating lanes per phase (count builds)
};

// ------------------------------------------------------------------------------------------------
// Start of path p = (pixel idx, sample c of the chunk): per-(pixel,sample) RNG stream, jittered camera ray and the first
// offsetRay of the bounce loop (trace.cpp:126). There is no generate kernel: the bounce-0 trace kernel and the bounce-0
// shade kernel both call this (a few hundred instructions) instead of writing and re-reading 80 B per path.
__device__ __forceinline__ void wf_camera_path(const WfArgs& a, uint32_t p, V3& o, V3& d, Rng& rng) {
  const TraceArgs& t = a.t;
  const uint32_t idx = a.chunkShift >= 0 ? p >> a.chunkShift : p / a.chunk, c = p - idx * a.chunk;
  const float* tr = t.rays + (size_t)idx * TR_WORDS;
  const float row = tr[TR_ROW], col = tr[TR_COL];
  const uint32_t pixelIndex = (uint32_t)row * (uint32_t)t.imageWidth + (uint32_t)col;
  d = camera_ray(t, row, col, pixelIndex, t.firstSample + c, rng);
  o = offset_origin(mk(0.f, 0.f, 0.f), d, mk(0.f, 0.f, 1.f));
}

// ------------------------------------------------------------------------------------------------
// wf_trace: one closest-hit query per live path of the bounce (CompactBvh::intersect, include/CompactBvh.hpp:80-139,
// near-first order over the pair table; the steps themselves are rt_prims.h "Streaming traversal").
//
// Persistent 1024-thread CTAs, the pair table staged in shared memory when it fits. Every lane owns one query at a
// time and is in one of three phases, told apart by the node reference it holds: TRAV (an inner node: load its 48-byte
// pair record, test both child boxes, descend / defer / pop), LEAF (a leaf: run the primitive test, then pop) or FETCH
// (query finished: store the hit, take the next ray of the warp's batch, test the root). A warp iteration executes ONE
// phase, chosen by ballot: inner-node steps as long as at least kTravThreshold lanes want one (a single ballot on
// that path), otherwise whichever phase most lanes wait for. Lanes therefore never wait for a neighbour's long
// traversal, only for their phase to be scheduled.
enum : uint32_t { WF_TRAV = 0, WF_LEAF = 1, WF_FETCH = 2 };
// measured on the B200 (scripts/wf_kernel_times.py, 1.940 / 1.888 / 1.938 ms per launch at 8 / 12 / 16); the scheduler
// model (scripts/wf_sched_sim.cpp) has its optimum at the same place
#ifndef B200RT_TRAV_THRESHOLD
#define B200RT_TRAV_THRESHOLD 12
#endif
#ifndef B200RT_TRAV_THRESHOLD_FIRST
#define B200RT_TRAV_THRESHOLD_FIRST B200RT_TRAV_THRESHOLD
#endif
#ifndef B200RT_WF_CLAIM
#define B200RT_WF_CLAIM 128
#endif
#ifndef B200RT_SHADE_PREFETCH
#define B200RT_SHADE_PREFETCH 1
#endif
#ifndef B200RT_TRACE_PREFETCH
#define B200RT_TRACE_PREFETCH 0
#endif
#ifndef B200RT_SHADE_STREAMING_HINTS
#define B200RT_SHADE_STREAMING_HINTS 1
#endif
#if B200RT_SHADE_STREAMING_HINTS
#define WF_LD(p) __ldcs(p)
#define WF_ST(p, v) __stcs(p, v)
#else
#define WF_LD(p) (*(p))
#define WF_ST(p, v) (*(p) = (v))
#endif
// wf_trace: path records read and hits written with streaming (evict-first) hints, so that they do not displace the pair
// table and the traversal stacks from L1 (matters for the L2-resident form). A/B knob.
#ifndef B200RT_TRACE_STREAMING_HINTS
#define B200RT_TRACE_STREAMING_HINTS 0
#endif
#if B200RT_TRACE_STREAMING_HINTS
#define WFT_LD(p) __ldcs(p)
#define WFT_ST(p, v) __stcs(p, v)
#else
#define WFT_LD(p) (*(p))
#define WFT_ST(p, v) (*(p) = (v))
#endif
#ifndef B200RT_SHADE_BLOCKS
#define B200RT_SHADE_BLOCKS 6
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// The kernel's body: one bounce's queries for the CTA's warps. `pairs` / `pairsShared` = the pair table (staged copy and
// its shared-window address when kShared). Called by wf_trace_kernel (one launch per bounce) and by wf_tail_kernel.
template <bool kShared, bool kCount, bool kFirst>
__device__ __forceinline__ void wf_trace_body(const WfArgs& a, const int qIn, const uint4* pairs, const uint32_t pairsShared,
                                              unsigned long long* phaseStats) {
  const DevScene& sc = a.t.scene;
  const unsigned lane = threadIdx.x & 31, full = 0xffffffffu;
  // bounce 0 (kFirst): slot = path over all paths of the chunk and the rays are the camera rays
  const WfState& in = a.b.st[qIn];
  const uint32_t count = kFirst ? a.numPaths : __ldcg(a.b.counts + qIn);
  uint32_t* cursor = a.b.counts + 2;
  // the shading kernel that follows appends its survivors to the other array: empty it (nobody reads that counter here)
  // ... and that kernel claims its rounds from slot 0 again
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.b.counts[qIn ^ 1] = 0u; a.b.counts[3] = 0u; }
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0;
  constexpr int kTravThreshold = kFirst ? B200RT_TRAV_THRESHOLD_FIRST : B200RT_TRAV_THRESHOLD;

  StreamQuery q;  // the lane's query
  q.o = q.op = q.inv = mk(0.f, 0.f, 0.f); q.d = mk(0.f, 0.f, -1.f);
  q.sx = q.sy = q.sz = 0.f; q.permOfs = 0u; q.fast = true;
  q.hitT = __int_as_float(0x7f800000); q.hitRef = kRefNone; q.b0 = q.b1 = q.b2 = 0.f;
  q.ref = kRefNone; q.topRef = q.top2Ref = kRefNone; q.topE = q.top2E = 0.f; q.sp = 1;
  uint32_t slot = 0xFFFFFFFFu;  // the slot whose query this is; none before the first fetch
  uint2 stack[kMaxStack + 1];   // deferred children below the register-held top: {reference, entry distance}
  // slots are claimed kClaim at a time per warp: one same-address atomic per 128 rays
  constexpr uint32_t kClaim = B200RT_WF_CLAIM;
  uint32_t wNext = 0, wEnd = 0;  // warp-uniform: the unclaimed part of this warp's current batch
  unsigned phaseIters[3] = {0u, 0u, 0u}, phaseLanes[3] = {0u, 0u, 0u};  // kCount builds: scheduler statistics

  while (true) {
    bool wantT = ref_is_inner(q.ref);
    unsigned mT = __ballot_sync(full, wantT);
    int cT = __popc(mT);
    uint32_t pick = WF_TRAV;
    unsigned mF = 0u;
    int cF = 0, cL = 0;
    if (cT < kTravThreshold) {
      const unsigned mL = __ballot_sync(full, ref_is_leaf(q.ref));
      mF = __ballot_sync(full, q.ref == kRefNone);
      if (!(mT | mL | mF)) break;  // every lane is kRefDone
      cL = __popc(mL); cF = __popc(mF);
      if (!(cT >= cL && cT >= cF)) pick = cL >= cF ? WF_LEAF : WF_FETCH;
    }
    if (kCount && phaseStats && lane == 0 && pick != WF_TRAV) {
      phaseIters[pick] += 1u;
      phaseLanes[pick] += (unsigned)(pick == WF_LEAF ? cL : cF);
    }

    bool again = false;  // the node just popped is culled: keep popping
    if (pick == WF_TRAV) {
      // inner-node steps, one after the other while enough lanes want one: the loop votes for itself, so a run of steps
      // pays for one ballot each instead of a trip through the phase selection above
      do {
        if (kCount && phaseStats && lane == 0) { phaseIters[WF_TRAV] += 1u; phaseLanes[WF_TRAV] += (unsigned)cT; }
        again = false;
        if (wantT) {
          PairWords w;
          if (kShared) {
            // (kRefInner | pair) * 48 wraps to pair * 48 in 32 bits: the reference needs no mask
            const uint32_t at = pairsShared + q.ref * 48u;
            w.q0 = lds128(at); w.q1 = lds128(at + 16u); w.q2 = lds128(at + 32u);
          } else {
            w = fetch_pair<false>(pairs, ref_pair(q.ref));
          }
          if (kCount) cnt.nodeVisits += 2;
          again = stream_trav<true>(q, w, stack);
        }
        while (again) again = stream_pop(q, stack);
        wantT = ref_is_inner(q.ref);
        mT = __ballot_sync(full, wantT);
        cT = __popc(mT);
      } while (cT >= kTravThreshold);
    } else if (pick == WF_LEAF) {
      if (ref_is_leaf(q.ref)) {
        if (kCount) cnt.primTests++;
        again = stream_leaf<true>(sc, q, stack, kFirst ? nullptr : in.rayD + slot);
      }
    } else {
      // lanes that finished a query store its result and take the next slots of the warp's batch
      if (wNext == wEnd) {
        uint32_t claimBase = 0;
        if (lane == 0) claimBase = atomicAdd(cursor, kClaim);
        wNext = __shfl_sync(full, claimBase, 0);
        wEnd = wNext + kClaim;
#if B200RT_TRACE_PREFETCH
        if (!kFirst) {
#pragma unroll
          for (uint32_t k = 0; k < kClaim; k += 32u) {
            const uint32_t s = wNext + k + lane;
            if (s < count) { prefetch_l2(in.rayO + s); prefetch_l2(in.rayI + s); prefetch_l2(in.rayS + s); }
          }
        }
#endif
      }
      const uint32_t avail = wEnd - wNext;
      const uint32_t rank = (uint32_t)__popc(mF & ((1u << lane) - 1u));
      const bool served = q.ref == kRefNone && rank < avail;  // the others retry on the next fetch step
      const uint32_t qi = wNext + rank;
      wNext += min((uint32_t)cF, avail);
      if (served) {
        if (slot != 0xFFFFFFFFu) {
          // the shading kernel turns the winner's leaf reference into geomID / primID (stream_hit_ids)
          WFT_ST(a.b.hitA + slot, make_float2(q.hitT, __uint_as_float(q.hitRef)));
          if (a.b.hitB) WFT_ST(a.b.hitB + slot, make_float4(q.b0, q.b1, q.b2, 0.f));
          slot = 0xFFFFFFFFu;
        }
        if (qi >= count) {
          q.ref = kRefDone;
        } else {
          // start of CompactBvh::intersect (tMin = 0, tMax = inf as set by the bounce loop, trace.cpp:128-130); a ray
          // that misses the root holds kRefNone again: its (empty) result is stored on the next fetch step
          nClosest++;
          if (kCount) cnt.nodeVisits++;
          slot = qi;
          if (kFirst) {
            V3 o, d;
            Rng unused;
            wf_camera_path(a, slot, o, d, unused);
            stream_begin(sc, q, o, d);
          } else {
            // bounce rays come with their constants (the shading kernel made them: stream_prepare)
            const float4 ro = WFT_LD(in.rayO + slot), ri = WFT_LD(in.rayI + slot), rs = WFT_LD(in.rayS + slot);
            stream_begin_prepared(sc, q, mk(ro.x, ro.y, ro.z), mk(ri.x, ri.y, ri.z), rs.x, rs.y, rs.z, __float_as_uint(rs.w));
          }
          // the handful of queries in which a NaN could occur (a zero in the direction, coordinates beyond 2^20) take
          // the NaN-preserving steps, the whole query here and now; the loop around only ever runs the fast steps
          if (!q.fast) stream_run<false, kCount>(sc, sc.pairs, q, stack, kFirst ? nullptr : in.rayD + slot, cnt.nodeVisits, cnt.primTests);
        }
      }
    }
    // one loop for both phases, after them: the reload of the stack top that a pop issues is then not consumed until the
    // lane's next pop (inside the phases the compiler would move it into place at once and wait for it)
    while (again) again = stream_pop(q, stack);
  }
  if (kCount && phaseStats && lane == 0)
    for (int k = 0; k < 3; ++k) { atomicAdd(phaseStats + 2 * k, (unsigned long long)phaseIters[k]); atomicAdd(phaseStats + 2 * k + 1, (unsigned long long)phaseLanes[k]); }
  flush_counters(a.t.counters, nClosest, 0u, cnt, 0u, 0u);
}

// Shared-window address of the staged table, made opaque to the compiler: sm_100 materialises a shared symbol's
// address as (CTA rank in cluster << 24 | offset) with an S2R, and would redo that in every inner-node step.
template <bool kShared>
__device__ __forceinline__ uint32_t opaque_shared_base(const void* smemRaw) {
  uint32_t base = kShared ? (uint32_t)__cvta_generic_to_shared(smemRaw) : 0u;
  asm volatile("" : "+r"(base));
  return base;
}

// The L2-resident form is launched as 128-thread CTAs, eight per SM at 64 registers (B200RT_WF_L2_MINBLOCKS: A/B knob for
// leaner CTAs beside the NIF kernel).
#ifndef B200RT_WF_L2_MINBLOCKS
#define B200RT_WF_L2_MINBLOCKS 8
#endif
template <bool kShared, bool kCount, bool kFirst>
__global__ void __launch_bounds__(kShared ? 1024 : 128, kShared ? 1 : B200RT_WF_L2_MINBLOCKS) wf_trace_kernel(const WfArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const uint4* pairs = stage_pairs<kShared>(a.t, reinterpret_cast<uint4*>(smemRaw));
  wf_trace_body<kShared, kCount, kFirst>(a, a.qIn, pairs, opaque_shared_base<kShared>(smemRaw), a.phaseStats);
}

// ------------------------------------------------------------------------------------------------
// wf_shade: the rest of one bounce-loop iteration (trace.cpp:133-184) for every live path, one thread per slot:
// updateHit, material, BxDF sample, roulette. A surviving path's record (next ray with its offset applied and its query
// constants prepared) is appended to the other state array; a finished path writes its throughput / escape record
// (and the HitRecord if it belongs to the last sample). Appends are aggregated over the block: one same-address atomic
// per 256 paths and array.
// 6 blocks per SM (40 registers): the kernel waits on dependent loads (hit -> leafInfo -> vertices -> material), so it is
// occupancy that buys time here; measured 1.6 ms per launch at 1 block's worth of registers, 1.0 at 4-8.
#ifndef B200RT_SHADE_THREADS
#define B200RT_SHADE_THREADS 128
#endif
constexpr int kShadeThreads = B200RT_SHADE_THREADS;  // 256, 128 or 64: the warps of a block share its two barriers per round
constexpr int kShadeWarps = kShadeThreads / 32;
constexpr int kShadeBlocksPerSM = B200RT_SHADE_BLOCKS * (256 / kShadeThreads);
// The kernel's body for blocks of kWarps warps (2, 4 or 8: wf_shade_kernel; 32: wf_tail_kernel). sCountBuf is the
// block's [round parity][queue][warp] scratch: the next round posts into the other half.
template <bool kNif, bool kFirst, int kWarps>
__device__ __forceinline__ void wf_shade_body(const WfArgs& a, const int qIn, uint32_t (*sCountBuf)[2][kWarps < 8 ? 8 : kWarps],
                                              uint32_t* sClaim) {
  const TraceArgs& t = a.t;
  const DevScene& sc = t.scene;
  const unsigned lane = threadIdx.x & 31, full = 0xffffffffu;
  const WfState& in = a.b.st[qIn];
  const WfState& out = a.b.st[qIn ^ 1];
  const uint32_t count = kFirst ? a.numPaths : __ldcg(a.b.counts + qIn);
  uint32_t* countOut = a.b.counts + (qIn ^ 1);
  // the next trace kernel starts fetching at slot 0 again (this bounce's trace kernel is done with the cursor)
  if (blockIdx.x == 0 && threadIdx.x == 0) a.b.counts[2] = 0u;
  unsigned nSamples = 0, nEscaped = 0;
  static_assert(kWarps == 32 || kWarps == 8 || kWarps == 4 || kWarps == 2, "block-level append: 2, 4, 8 or 32 warps");
  // Whole blocks iterate together (uniform trip count) so the ballots and barriers below see converged threads. A round
  // is blockDim.x consecutive slots claimed from a grid-wide cursor (counts[3]): blocks that run beside another kernel's
  // CTA on their SM (chunk overlap: the NIF kernel of the previous chunk), or start late because of it, simply claim
  // fewer rounds. Claims run two rounds ahead (sClaim is a ring of 4) so that the next round's slots are known at the top
  // of this one for the L2 prefetch; a claim is published by the round's first barrier and read after its second.
  uint32_t* const roundCursor = a.b.counts + 3;
  if (threadIdx.x == 0) { sClaim[0] = atomicAdd(roundCursor, blockDim.x); sClaim[1] = atomicAdd(roundCursor, blockDim.x); }
  __syncthreads();
  uint32_t cur = sClaim[0], nxt = sClaim[1];
  for (uint32_t r = 0; cur < count; ++r) {
    uint32_t (*sCount)[kWarps < 8 ? 8 : kWarps] = sCountBuf[r & 1u];
    if (threadIdx.x == 0) sClaim[(r + 2u) & 3u] = nxt < count ? atomicAdd(roundCursor, blockDim.x) : 0xFFFFFFFFu;
    const uint32_t i = cur + threadIdx.x;
    const bool valid = i < count;
#if B200RT_SHADE_PREFETCH
    // the records of the block's next round are on their way from DRAM to L2 while this round is shaded
    if (nxt < count && nxt + threadIdx.x < count) {
      const uint32_t in2 = nxt + threadIdx.x;
      prefetch_l2(a.b.hitA + in2);
      if (!kFirst) { prefetch_l2(in.rayO + in2); prefetch_l2(in.rayD + in2); prefetch_l2(in.thr + in2); prefetch_l2(in.rng + in2); }
    }
#endif
    bool survive = false, lastOne = false;
    uint32_t appendSlot = 0xFFFFFFFFu, p = 0;
    // the survivor's next record, held in registers until its slot is known
    V3 o = mk(0.f, 0.f, 0.f), d = mk(0.f, 0.f, -1.f), n = mk(0.f, 0.f, 1.f), thr = mk(1.f, 1.f, 1.f);
    uint32_t bounce = 0, flags = 0, geomID = kInvalidGeom, primID = kInvalidPrim;
    Rng rng;
    rng.s0 = rng.s1 = 0ull;
    if (valid) {
      const float2 ha = WF_LD(a.b.hitA + i);
      Hit hit;
      hit.t = ha.x;
      stream_hit_ids(sc, __float_as_uint(ha.y), hit.geomID, hit.primID, hit.tri);
      hit.node = 0; hit.b0 = hit.b1 = hit.b2 = 0.f;
      if (a.b.hitB) { const float4 hb = a.b.hitB[i]; hit.b0 = hb.x; hit.b1 = hb.y; hit.b2 = hb.z; }
      if (kFirst) {
        p = i;
        wf_camera_path(a, p, o, d, rng);  // same ray, same RNG state as the trace kernel started from
      } else {
        const float4 ro = WF_LD(in.rayO + i), rd = WF_LD(in.rayD + i), th = WF_LD(in.thr + i);
        const uint4 rs = WF_LD(in.rng + i);
        p = __float_as_uint(th.w);
        o = mk(ro.x, ro.y, ro.z); d = mk(rd.x, rd.y, rd.z);
        thr = mk(th.x, th.y, th.z);
        const uint32_t packed = __float_as_uint(ro.w);
        bounce = packed & 0xffu; flags = (packed >> 8) & 0xffu; geomID = packed >> 16;
        primID = __float_as_uint(rd.w);
        rng.s0 = (uint64_t)rs.x | ((uint64_t)rs.y << 32);
        rng.s1 = (uint64_t)rs.z | ((uint64_t)rs.w << 32);
      }
      const uint32_t idx = a.chunkShift >= 0 ? p >> a.chunkShift : p / a.chunk, c = p - idx * a.chunk;
      const uint32_t s = t.firstSample + c;
      lastOne = s == a.lastSample;  // this path's HitRecord is the one left in the ray stream
      if (!kFirst && lastOne) { const float4 nn = in.nrm[i]; n = mk(nn.x, nn.y, nn.z); }
      V3 emitted = mk(0.f, 0.f, 0.f);  // thr * emission picked up at this bounce
      bool gotEmission = false, poisoned = false;
      if (bounce == 0) nSamples++;

      // ---- rest of one bounce-loop iteration (trace.cpp:133-184) ----
      bool ended = false, escaped = false;
      const float tMaxOut = hit.t;
      if (hit.geomID != kInvalidGeom) {
        geomID = hit.geomID; primID = hit.primID;
        o = o + d * hit.t;
        n = hit_normal(sc, hit, o);
        const Mat m = load_material(sc, geomID);
        if (m.emissive) { emitted = thr * m.emission; gotEmission = true; }
        if (m.type == 0) {
          const float u1 = rng_uniform(rng);
          const float u2 = rng_uniform(rng);
          d = sample_diffuse(n, u1, u2);
          thr = thr * m.albedo;
        } else if (m.type == 1) {
          d = reflect_dir(d, n);
          thr = thr * m.albedo;
        } else if (m.type == 2) {
          const float u1 = rng_uniform(rng);
          bool refracted;
          d = dielectric_dir(d, n, m.ior, u1, refracted);
          if (refracted) thr = thr * m.albedo;
        } else {
          poisoned = true;  // poisons rgb like result.rgb *= NaN (trace.cpp:167)
          flags |= kFlagError;
        }
      } else {
        flags |= kFlagEscaped;
        ended = true;
        escaped = true;
      }
      if (!ended) {
        if (bounce > t.rouletteStartDepth) {
          const float u1 = rng_uniform(rng);
          const float pr = maxc(thr);  // evaluateRoulette (geometric_sampling.hpp:56-63)
          if (pr == 0.f || u1 > pr) ended = true;
          else thr = thr * (1.f / pr);
        }
        bounce++;
        if (bounce >= t.maxPathLength) ended = true;
      }

      // running colour of the path: color = color + thr * emission, then (unknown material) color = color * NaN
      if (kFirst || gotEmission || poisoned) {
        float* sc3 = t.slotColor + 3 * (size_t)p;
        V3 color = kFirst ? mk(0.f, 0.f, 0.f) : mk(sc3[0], sc3[1], sc3[2]);
        if (gotEmission) color = color + emitted;
        if (poisoned) color = color * __int_as_float(0x7fc00000);
        sc3[0] = color.x; sc3[1] = color.y; sc3[2] = color.z;
      }
      if (!ended) {
        o = offset_origin(o, d, n);  // offsetRay at the top of the next iteration (trace.cpp:126)
        survive = true;
      } else {
        if (escaped) nEscaped++;
        if (kNif) {
          float* se = t.slotEscape + 5 * (size_t)p;
          float u = -1.f, v = 0.f;
          if (escaped) escaped_uv(d, t.hdriRotation, u, v);
          se[0] = thr.x; se[1] = thr.y; se[2] = thr.z; se[3] = u; se[4] = v;
          if (escaped) appendSlot = p;
        }
        if (lastOne) {
          // the HitRecord the reference leaves in the stream is the last sample's
          float* tr = t.rays + (size_t)idx * TR_WORDS;
          tr[TR_ORIGIN] = o.x; tr[TR_ORIGIN + 1] = o.y; tr[TR_ORIGIN + 2] = o.z;
          tr[TR_TMIN] = 0.f;
          tr[TR_DIR] = d.x; tr[TR_DIR + 1] = d.y; tr[TR_DIR + 2] = d.z;
          tr[TR_TMAX] = tMaxOut;
          tr[TR_PRIM] = __uint_as_float(primID);
          tr[TR_NORMAL] = n.x; tr[TR_NORMAL + 1] = n.y; tr[TR_NORMAL + 2] = n.z;
          tr[TR_THROUGHPUT] = thr.x; tr[TR_THROUGHPUT + 1] = thr.y; tr[TR_THROUGHPUT + 2] = thr.z;
          tr[TR_IDS] = __uint_as_float(geomID | (flags << 16));
        }
      }
    }
    // regroup: survivors go to the other state array, escaped slots to the NIF queue. Every warp posts its two counts,
    // warp 0 reserves both ranges, each warp then knows its offset.
    {
      const unsigned mS = __ballot_sync(full, survive);
      const unsigned mE = kNif ? __ballot_sync(full, appendSlot != 0xFFFFFFFFu) : 0u;
      const int warp = threadIdx.x >> 5;
      if (lane == 0) { sCount[0][warp] = (uint32_t)__popc(mS); sCount[1][warp] = (uint32_t)__popc(mE); }
      __syncthreads();
      if (warp == 0) {
        if (kWarps <= 8) {
          // exclusive scan of the (up to) 8 warp counts of each queue in lanes 0..7 / 8..15
          const int q = lane >> 3, w = lane & 7;
          uint32_t v = (lane < 16 && w < kWarps) ? sCount[q][w] : 0u, incl = v;
#pragma unroll
          for (int off = 1; off < 8; off <<= 1) {
            const uint32_t up = __shfl_up_sync(full, incl, off, 8);
            if (w >= off) incl += up;
          }
          uint32_t base = 0;
          if (lane == 7 && incl) base = atomicAdd(countOut, incl);
          if (lane == 15 && incl) base = atomicAdd(t.escapeCount, incl);
          base = __shfl_sync(full, base, (lane & 8) | 7, 32);
          if (lane < 16) sCount[q][w] = base + incl - v;
        } else {
          // 32 warps: one scan per queue over all lanes
#pragma unroll
          for (int q = 0; q < (kNif ? 2 : 1); ++q) {
            const uint32_t v = sCount[q][lane];
            uint32_t incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const uint32_t up = __shfl_up_sync(full, incl, off);
              if ((int)lane >= off) incl += up;
            }
            uint32_t base = 0;
            if (lane == 31 && incl) base = atomicAdd(q == 0 ? countOut : t.escapeCount, incl);
            base = __shfl_sync(full, base, 31);
            sCount[q][lane] = base + incl - v;
          }
        }
      }
      __syncthreads();
      if (survive) {
        // consecutive survivors of a warp write consecutive records: full sectors
        const uint32_t j = sCount[0][warp] + __popc(mS & ((1u << lane) - 1u));
        WF_ST(out.rayO + j, make_float4(o.x, o.y, o.z, __uint_as_float(bounce | (flags << 8) | (geomID << 16))));
        WF_ST(out.rayD + j, make_float4(d.x, d.y, d.z, __uint_as_float(primID)));
        WF_ST(out.thr + j, make_float4(thr.x, thr.y, thr.z, __uint_as_float(p)));
        WF_ST(out.rng + j, make_uint4((uint32_t)rng.s0, (uint32_t)(rng.s0 >> 32), (uint32_t)rng.s1, (uint32_t)(rng.s1 >> 32)));
        if (lastOne) out.nrm[j] = make_float4(n.x, n.y, n.z, 0.f);
        // constants of the next query, computed here where every lane has a ray (wf_trace's fetch phase runs at ~20 lanes)
        V3 inv;
        float sx, sy, sz;
        uint32_t qflags;
        stream_prepare(sc, o, d, inv, sx, sy, sz, qflags);
        WF_ST(out.rayI + j, make_float4(inv.x, inv.y, inv.z, 0.f));
        WF_ST(out.rayS + j, make_float4(sx, sy, sz, __uint_as_float(qflags)));
      }
      if (kNif && appendSlot != 0xFFFFFFFFu) t.escapeQueue[sCount[1][warp] + __popc(mE & ((1u << lane) - 1u))] = appendSlot;
      // no third barrier: the next round posts its counts into the other half of sCountBuf, and a warp can only get
      // two rounds ahead of another by passing that round's first barrier, which every warp must reach
    }
    cur = nxt;
    nxt = sClaim[(r + 2u) & 3u];
  }
  flush_counters(t.counters, 0u, 0u, Counters{0u, 0u}, nSamples, nEscaped);
}

template <bool kNif, bool kFirst>
__global__ void __launch_bounds__(kShadeThreads, kShadeBlocksPerSM) wf_shade_kernel(const WfArgs a) {
  __shared__ uint32_t sCountBuf[2][2][8];
  __shared__ uint32_t sClaim[4];
  wf_shade_body<kNif, kFirst, kShadeWarps>(a, a.qIn, sCountBuf, sClaim);
}

// wf_tail: the bounces from `bounceBegin` on in ONE cooperative launch (persistent 1024-thread CTAs, one per SM, the pair
// table staged once): trace phase, grid barrier, shade phase, grid barrier, until no path is left or bounceEnd is
// reached. In a render with Russian roulette the paths still alive a bounce or two after roulette starts are a fraction
// of a per cent of the chunk (bench scene: 65 K of 66 M at bounce 5), and each of their bounces otherwise costs two
// launches, a staging of the table and a pipeline drain. Same bodies, same arithmetic, same queues: results are those
// of the per-bounce launches (tests/test_gpu_parity.py runs both). The grid barrier orders the phases' global-memory
// traffic (cooperative groups' grid.sync() is a release / acquire at gpu scope), and the path records are never read
// through the non-coherent path.
template <bool kShared, bool kCount, bool kNif>
__global__ void __launch_bounds__(1024) wf_tail_kernel(const WfArgs a, const uint32_t bounceBegin, const uint32_t bounceEnd) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint32_t sCountBuf[2][2][32];
  __shared__ uint32_t sClaim[4];
  const uint4* pairs = stage_pairs<kShared>(a.t, reinterpret_cast<uint4*>(smemRaw));
  const uint32_t pairsShared = opaque_shared_base<kShared>(smemRaw);
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  for (uint32_t b = bounceBegin; b < bounceEnd; ++b) {
    const int qIn = (int)(b & 1u);
    wf_trace_body<kShared, kCount, false>(a, qIn, pairs, pairsShared, nullptr);
    grid.sync();
    wf_shade_body<kNif, false, 32>(a, qIn, sCountBuf, sClaim);
    grid.sync();
    if (__ldcg(a.b.counts + (qIn ^ 1)) == 0u) break;  // nobody survived this bounce (every thread reads the same value)
  }
}

// rgb += colour_s (+ throughput_s * env_s when an environment light is loaded), s in chunk order.
//
// The slot arrays are [pixel][sample]: a thread that walks its pixel's samples reads with a stride of chunk * 12 B
// between lanes. So a block first turns the slots of its kAccPixels pixels into the two addends of every sample
// (colour, throughput * env) with fully coalesced loads, parks them in shared memory, and only then does each thread
// add up its pixel's row -- in the same order and with the same operations as the sequential loop.
constexpr int kAccPixels = 48;   // 48 x (32 x 7 + 1) floats = 43 KB of static shared memory
constexpr int kAccSamples = 32;  // samples staged per pass (chunks may be much longer than that on small ray streams)
__global__ void __launch_bounds__(256) wf_accumulate_kernel(float* rays, uint32_t numRays, uint32_t chunk, const float* slotColor,
                                                           const float* slotEscape, const float* slotEnv) {
  // [kAccPixels][kAccSamples * 7 + 1]: per sample colour.xyz, (throughput * env).xyz, lit flag
  __shared__ float accSmem[kAccPixels * (kAccSamples * 7 + 1)];
  constexpr uint32_t rowWords = kAccSamples * 7u + 1u;  // odd stride: the per-pixel walk below is bank-conflict free
  const uint32_t pix0 = blockIdx.x * kAccPixels;
  const uint32_t pixels = min((uint32_t)kAccPixels, numRays - pix0);
  float* tr = rays + (size_t)(pix0 + threadIdx.x) * TR_WORDS;
  V3 rgb = mk(0.f, 0.f, 0.f);
  if (threadIdx.x < pixels) rgb = mk(tr[TR_RGB], tr[TR_RGB + 1], tr[TR_RGB + 2]);
  for (uint32_t c0 = 0; c0 < chunk; c0 += kAccSamples) {
    const uint32_t cs = min((uint32_t)kAccSamples, chunk - c0);
    for (uint32_t i = threadIdx.x; i < pixels * cs; i += blockDim.x) {
      const uint32_t pl = i / cs, c = i - pl * cs;
      const size_t slot = (size_t)(pix0 + pl) * chunk + c0 + c;
      float* dst = accSmem + pl * rowWords + c * 7u;
      const float* col = slotColor + 3 * slot;
      dst[0] = col[0]; dst[1] = col[1]; dst[2] = col[2];
      float ex = 0.f, ey = 0.f, ez = 0.f;
      bool lit = false;
      if (slotEnv) {
        const float* se = slotEscape + 5 * slot;
        if (se[3] >= 0.f) {
          const float* env = slotEnv + 3 * slot;
          const V3 e = mk(se[0], se[1], se[2]) * mk(env[2], env[1], env[0]);
          ex = e.x; ey = e.y; ez = e.z;
          lit = true;
        }
      }
      // a sample without an environment term must not add anything (not even +0: -0 + 0 would flip a sign bit)
      dst[3] = ex; dst[4] = ey; dst[5] = ez; dst[6] = lit ? 1.f : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < pixels) {
      const float* row = accSmem + threadIdx.x * rowWords;
      for (uint32_t c = 0; c < cs; ++c) {
        const float* a = row + c * 7u;
        rgb = rgb + mk(a[0], a[1], a[2]);
        if (a[6] != 0.f) rgb = rgb + mk(a[3], a[4], a[5]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < pixels) { tr[TR_RGB] = rgb.x; tr[TR_RGB + 1] = rgb.y; tr[TR_RGB + 2] = rgb.z; }
}


}  // namespace rt
