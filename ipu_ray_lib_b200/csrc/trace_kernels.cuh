// The two render kernels of the trace path.
//
//   shadow_trace_kernel : traceShadowRay over a ray stream (include/Render.hpp:37-72; IPU vertex
//                         ShadowTrace, codelets/TraceCodelets.cpp:269-316)
//   path_trace_kernel   : per-sample camera ray + the bounce loop (trace.cpp:115-188; IPU vertex
//                         PathTrace + sampleCameraRays, codelets/TraceCodelets.cpp:142-264)
//
// Both are persistent: the grid is sized to the SM count, every warp pulls 32-ray chunks of the
// TraceResult stream from a global counter until the stream is exhausted. In the path tracer each
// lane owns one pixel and regenerates its next sample the moment its current path ends, so lanes
// stay busy through the bounce loop regardless of path length, while rgb is still accumulated in
// sample order (bit-identical to the sequential reference loop).
#pragma once
#include "async_copy.cuh"
#include "rt_device.cuh"

namespace rt {

// Word offsets inside the 84-byte TraceResult (include/embree_utils/geometry.hpp:227-260).
enum : int {
  TR_RGB = 0, TR_ROW = 3, TR_COL = 4, TR_ORIGIN = 5, TR_TMIN = 8, TR_DIR = 9, TR_TMAX = 12,
  TR_PRIM = 13, TR_NORMAL = 14, TR_THROUGHPUT = 17, TR_IDS = 20, TR_WORDS = 21
};
constexpr uint32_t kFlagError = 1u, kFlagEscaped = 2u;

struct DeviceCounters {
  unsigned long long closest, occlusion, nodeVisits, primTests, samples, escaped;
};

struct TraceArgs {
  DevScene scene;
  float* rays;              // TraceResult[n] as words
  uint32_t numRays;
  uint32_t* workCounter;    // persistent-warp chunk counter (zeroed before launch)
  DeviceCounters* counters;
  uint32_t nodeBytes;       // size of the node array when staged into shared memory (reference-order / megakernel paths)
  // shadow trace
  float lightX, lightY, lightZ, ambient;
  // path trace
  float imageWidth, imageHeight, tanTheta, antiAlias;
  uint32_t maxPathLength, rouletteStartDepth;
  uint32_t firstSample, endSample;
  unsigned long long rngKey;  // splitmix64(rngSeed)
  // NIF wavefront outputs (path trace with an environment light); null otherwise
  float* slotColor;         // [numRays][samplesPerChunk][3]
  float* slotEscape;        // [numRays][samplesPerChunk][5] = throughput.xyz, u, v (u < 0: not escaped)
  uint32_t* escapeQueue;    // compacted slot indices of escaped samples
  uint32_t* escapeCount;
  float hdriRotation;       // radians
  int travThreshold;        // state machine: run inner-node steps while at least this many lanes want one
  // primary-ray pre-pass (null = off): closest hit of every camera ray of the chunk, traced warp-coherently
  uint4* primA;             // [numRays][chunk] = {t bits, geomID, primID, global triangle index}
  float4* primB;            // [numRays][chunk] = {b0, b1, b2, -}; only when the scene interpolates normals
};

__device__ __forceinline__ void flush_counters(DeviceCounters* out, unsigned closest, unsigned occl, const Counters& c,
                                               unsigned samples, unsigned escaped) {
  // warp-reduce then one atomic per counter per warp
  unsigned v[6] = {closest, occl, c.nodeVisits, c.primTests, samples, escaped};
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(out);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    unsigned long long x = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(dst + k, x);
  }
}

template <bool kShared>
__device__ __forceinline__ const uint2* stage_nodes(const TraceArgs& a, uint2* smem) {
  if (!kShared) return a.scene.nodes;
  // cooperative 16-byte copy of the whole node array into shared memory (array is 8 B aligned and a
  // multiple of 8 B; the device copy is allocated 16 B aligned and padded to 16 B)
  const uint4* src = reinterpret_cast<const uint4*>(a.scene.nodes);
  uint4* dst = reinterpret_cast<uint4*>(smem);
  const uint32_t n16 = (a.nodeBytes + 15u) / 16u;
  for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
  return smem;
}

// Same for the pair table (rt_prims.h) that the streaming kernels traverse: 48 B per inner node.
template <bool kShared>
__device__ __forceinline__ const uint4* stage_pairs(const TraceArgs& a, uint4* smem) {
  if (!kShared) return a.scene.pairs;
  const uint32_t n16 = a.scene.numPairs * 3u;
  for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) smem[i] = __ldg(a.scene.pairs + i);
  __syncthreads();
  return smem;
}

template <bool kShared, bool kOrdered, bool kCount>
__device__ __forceinline__ void closest_hit(const DevScene& sc, const uint2* nodes, V3 o, V3 d, float tMin, float tMax,
                                            Hit& h, Counters& c) {
  if (kOrdered) closest_hit_ordered<kShared, kCount>(sc, nodes, o, d, tMin, tMax, h, c);
  else closest_hit_ref_order<kShared, kCount>(sc, nodes, o, d, tMin, tMax, h, c);
}

// -------------------------------------------------------------------------------------------------
// traceShadowRay for one ray (include/Render.hpp:37-72): closest hit, then on a hit updateHit, the shadow ray towards
// the light and Lambert + ambient. kOrdered = near-first over the pair table; otherwise the reference's visiting order
// over the caller's own node array (identical node / primitive counters).
struct ShadowOut {
  bool hit;
  V3 color, o, n;
  float t;
  uint32_t primID, geomID;
};
template <bool kShared, bool kOrdered, bool kCount>
__device__ __forceinline__ ShadowOut shadow_one(const TraceArgs& a, const void* table, V3 o, V3 d, float tMin, float tMax,
                                                Counters& cnt, unsigned& nOccl) {
  const DevScene& sc = a.scene;
  ShadowOut r;
  Hit h;
  if (kOrdered) {
    const uint4* pairs = static_cast<const uint4*>(table);
    uint2 stack[kMaxStack];
    PairHit ph;
    pair_closest_hit<kShared, kCount>(sc, pairs, o, d, tMin, tMax, ph, stack, cnt.nodeVisits, cnt.primTests);
    h.t = ph.t; h.geomID = ph.geomID; h.b0 = ph.b0; h.b1 = ph.b1; h.b2 = ph.b2; h.node = 0;
    hit_ids(sc, ph, h.primID, h.tri);
  } else {
    closest_hit_ref_order<kShared, kCount>(sc, static_cast<const uint2*>(table), o, d, tMin, tMax, h, cnt);
  }
  r.hit = h.geomID != kInvalidGeom;
  r.t = h.t; r.primID = h.primID; r.geomID = h.geomID;
  r.o = o; r.n = mk(0.f, 0.f, 1.f); r.color = mk(0.f, 0.f, 0.f);
  if (r.hit) {
    // updateHit (Render.hpp:15-23)
    r.o = o + d * h.t;
    r.n = hit_normal(sc, h, r.o);
    const Mat m = load_material(sc, h.geomID);
    // shadow ray towards the light (Render.hpp:50-60); tMax is the distance BEFORE the offset
    const V3 lightOffset = mk(a.lightX, a.lightY, a.lightZ) - r.o;
    const V3 sd = normalized(lightOffset);
    const V3 so = offset_origin(r.o, sd, r.n);
    const float sMax = sqrtf(norm2(lightOffset));
    r.color = m.albedo * a.ambient;
    nOccl++;
    bool occluded;
    if (kOrdered) {
      uint32_t stack[kMaxStack];
      occluded = pair_any_hit<kShared, kCount>(sc, static_cast<const uint4*>(table), so, sd, 0.f, sMax, stack, cnt.nodeVisits, cnt.primTests);
    } else {
      occluded = any_hit<kShared, kCount>(sc, static_cast<const uint2*>(table), so, sd, 0.f, sMax, cnt);
    }
    if (!occluded) r.color = r.color + m.albedo * dot(sd, r.n);
  }
  return r;
}

template <bool kShared, bool kOrdered, bool kCount>
__global__ void __launch_bounds__(768) shadow_trace_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const void* table = kOrdered ? static_cast<const void*>(stage_pairs<kShared>(a, reinterpret_cast<uint4*>(smemRaw)))
                               : static_cast<const void*>(stage_nodes<kShared>(a, reinterpret_cast<uint2*>(smemRaw)));
  const unsigned lane = threadIdx.x & 31;
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0, nOccl = 0;

  while (true) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.workCounter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.numRays) break;
    const uint32_t idx = base + lane;
    if (idx >= a.numRays) continue;
    float* tr = a.rays + (size_t)idx * TR_WORDS;

    const V3 o = mk(tr[TR_ORIGIN], tr[TR_ORIGIN + 1], tr[TR_ORIGIN + 2]);
    const V3 d = mk(tr[TR_DIR], tr[TR_DIR + 1], tr[TR_DIR + 2]);
    const float tMin = tr[TR_TMIN], tMax = tr[TR_TMAX];
    nClosest++;
    const ShadowOut r = shadow_one<kShared, kOrdered, kCount>(a, table, o, d, tMin, tMax, cnt, nOccl);
    if (r.hit) {
      tr[TR_RGB] = r.color.x; tr[TR_RGB + 1] = r.color.y; tr[TR_RGB + 2] = r.color.z;
      tr[TR_ORIGIN] = r.o.x; tr[TR_ORIGIN + 1] = r.o.y; tr[TR_ORIGIN + 2] = r.o.z;
      tr[TR_TMAX] = r.t;
      tr[TR_PRIM] = __uint_as_float(r.primID);
      tr[TR_NORMAL] = r.n.x; tr[TR_NORMAL + 1] = r.n.y; tr[TR_NORMAL + 2] = r.n.z;
      const uint32_t ids = __float_as_uint(tr[TR_IDS]);
      tr[TR_IDS] = __uint_as_float((ids & 0xffff0000u) | r.geomID);
    } else {
      const uint32_t ids = __float_as_uint(tr[TR_IDS]);
      tr[TR_IDS] = __uint_as_float(ids | (kFlagEscaped << 16));
    }
  }
  flush_counters(a.counters, nClosest, nOccl, cnt, 0u, 0u);
}

// -------------------------------------------------------------------------------------------------
// shadow_stream_kernel: the same single-pass render with the ray stream moved by the TMA engine.
//
// The caller's TraceResult stream is AoS, 84 B per ray. A tile = 32 consecutive rays = 2688 contiguous, 16-byte
// aligned bytes, which is exactly what a 1-D bulk copy moves. Per persistent CTA (one per SM, the pair table staged in
// shared memory beside the rings):
//   * warp 0 is the PRODUCER: it claims tile ids from the global counter and keeps a ring of kLoadSlots tiles in flight
//     with cp.async.bulk global -> shared, each completing on the slot's mbarrier (the ring is the multi-buffering: up
//     to 8 tiles = 21 KB per SM are on their way while the other warps trace);
//   * warps 1..24 are CONSUMERS: each takes the next landed tile, every lane copies ITS ray's 21 words out of the slot
//     (word stride 21 is odd: conflict-free; this is the AoS -> per-lane "SoA in registers" transposition), the slot
//     goes straight back to the producer, the warp traces its 32 rays (closest hit + shadow ray), writes the 32 finished
//     records into a store slot and one lane sends it home with cp.async.bulk shared -> global.
// HBM therefore sees only full, aligned 2688-byte bursts in both directions, issued asynchronously to the traversal,
// instead of 21 strided word loads and 13 strided word stores per lane.
constexpr int kStreamTileRays = 32;
constexpr uint32_t kStreamTileBytes = kStreamTileRays * TR_WORDS * 4u;  // 2688
constexpr int kLoadSlots = 8, kStoreSlots = 6;
constexpr int kStreamThreads = 800;  // 24 consumer warps + the producer warp (80 registers x 800 threads fit the file)
constexpr uint32_t kStreamNoTile = 0xFFFFFFFFu;
struct StreamCtl {
  uint64_t full[kLoadSlots];    // tile landed (transaction bytes)
  uint64_t empty[kLoadSlots];   // tile copied into registers by its consumer
  uint32_t tileOf[kLoadSlots];  // tile id held by the slot, kStreamNoTile = end of stream
  uint32_t storeGen[kStoreSlots];  // completed uses of each store slot
  uint32_t published;           // load tickets the producer has issued
  uint32_t loadTicket, storeTicket;
};
constexpr uint32_t kStreamRingBytes = (kLoadSlots + kStoreSlots) * kStreamTileBytes + 256u;
static_assert(sizeof(StreamCtl) <= 256, "control block");

template <bool kShared, bool kCount>
__global__ void __launch_bounds__(kStreamThreads) shadow_stream_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const uint32_t pairBytes = kShared ? a.scene.numPairs * 48u : 0u;
  unsigned char* loadRing = smemRaw + pairBytes;
  unsigned char* storeRing = loadRing + kLoadSlots * kStreamTileBytes;
  StreamCtl* ctl = reinterpret_cast<StreamCtl*>(storeRing + kStoreSlots * kStreamTileBytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLoadSlots; ++i) { ac::mbar_init(&ctl->full[i], 1); ac::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < kStoreSlots; ++i) ctl->storeGen[i] = 0u;
    ctl->published = 0u; ctl->loadTicket = 0u; ctl->storeTicket = 0u;
    ac::fence_barrier_init();
  }
  const uint4* pairs = stage_pairs<kShared>(a, reinterpret_cast<uint4*>(smemRaw));
  __syncthreads();
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t numTiles = (a.numRays + kStreamTileRays - 1) / kStreamTileRays;
  const uint32_t numConsumers = (blockDim.x >> 5) - 1u;
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0, nOccl = 0;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ticket = 0;
      auto publish = [&](uint32_t t) {
        const uint32_t slot = ticket % kLoadSlots, phase = (ticket / kLoadSlots) & 1u;
        ac::mbar_wait(&ctl->empty[slot], phase ^ 1u);  // the slot's previous tile has been taken
        ctl->tileOf[slot] = t;
        const bool whole = t != kStreamNoTile && (size_t)(t + 1u) * kStreamTileRays <= a.numRays;
        if (whole) {
          ac::mbar_expect_tx(&ctl->full[slot], kStreamTileBytes);
          ac::bulk_load(loadRing + slot * kStreamTileBytes, a.rays + (size_t)t * kStreamTileRays * TR_WORDS, kStreamTileBytes,
                        &ctl->full[slot]);
        } else {
          ac::mbar_arrive(&ctl->full[slot]);  // end marker, or the ragged last tile (read with plain loads)
        }
        ++ticket;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&ctl->published) = ticket;
      };
      while (true) {
        const uint32_t base = atomicAdd(a.workCounter, 4u);
        if (base >= numTiles) break;
        const uint32_t end = min(base + 4u, numTiles);
        for (uint32_t t = base; t < end; ++t) publish(t);
      }
      for (uint32_t c = 0; c < numConsumers; ++c) publish(kStreamNoTile);
    }
  } else {
    bool stored = false;
    uint32_t pendingSlot = kStreamNoTile, pendingGen = 0;  // lane 0: store slot whose bulk store may still be reading it
    auto retire_store = [&]() {
      if (lane == 0 && pendingSlot != kStreamNoTile) {
        ac::bulk_wait_read_all();
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&ctl->storeGen[pendingSlot]) = pendingGen + 1u;
        pendingSlot = kStreamNoTile;
      }
    };
    while (true) {
      uint32_t ticket = 0;
      if (lane == 0) ticket = atomicAdd(&ctl->loadTicket, 1u);
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      const uint32_t slot = ticket % kLoadSlots, phase = (ticket / kLoadSlots) & 1u;
      // only wait on the slot's barrier once OUR ticket is the one in flight there (a parity bit cannot tell the phases
      // of tickets k and k + 2 * kLoadSlots apart, and more consumers than slots may be queueing)
      while (*reinterpret_cast<volatile uint32_t*>(&ctl->published) <= ticket) __nanosleep(32);
      ac::mbar_wait(&ctl->full[slot], phase);
      const uint32_t tile = ctl->tileOf[slot];
      if (tile == kStreamNoTile) {
        __syncwarp();
        if (lane == 0) ac::mbar_arrive(&ctl->empty[slot]);
        retire_store();
        break;
      }
      const uint32_t first = tile * kStreamTileRays;
      const uint32_t inTile = min((uint32_t)kStreamTileRays, a.numRays - first);
      const bool whole = inTile == kStreamTileRays;
      const bool mine = lane < inTile;
      float w[TR_WORDS];
      if (whole) {
        const float* src = reinterpret_cast<const float*>(loadRing + slot * kStreamTileBytes) + lane * TR_WORDS;
#pragma unroll
        for (int i = 0; i < TR_WORDS; ++i) w[i] = src[i];
      } else {
        const float* src = a.rays + (size_t)(first + (mine ? lane : 0u)) * TR_WORDS;
#pragma unroll
        for (int i = 0; i < TR_WORDS; ++i) w[i] = src[i];
      }
      __syncwarp();
      if (lane == 0) ac::mbar_arrive(&ctl->empty[slot]);  // slot back to the producer before the tracing starts
      retire_store();  // the previous tile's bulk store has had the whole load phase to drain its slot

      if (mine) {
        const V3 o = mk(w[TR_ORIGIN], w[TR_ORIGIN + 1], w[TR_ORIGIN + 2]);
        const V3 d = mk(w[TR_DIR], w[TR_DIR + 1], w[TR_DIR + 2]);
        nClosest++;
        const ShadowOut r = shadow_one<kShared, true, kCount>(a, pairs, o, d, w[TR_TMIN], w[TR_TMAX], cnt, nOccl);
        const uint32_t ids = __float_as_uint(w[TR_IDS]);
        if (r.hit) {
          w[TR_RGB] = r.color.x; w[TR_RGB + 1] = r.color.y; w[TR_RGB + 2] = r.color.z;
          w[TR_ORIGIN] = r.o.x; w[TR_ORIGIN + 1] = r.o.y; w[TR_ORIGIN + 2] = r.o.z;
          w[TR_TMAX] = r.t;
          w[TR_PRIM] = __uint_as_float(r.primID);
          w[TR_NORMAL] = r.n.x; w[TR_NORMAL + 1] = r.n.y; w[TR_NORMAL + 2] = r.n.z;
          w[TR_IDS] = __uint_as_float((ids & 0xffff0000u) | r.geomID);
        } else {
          w[TR_IDS] = __uint_as_float(ids | (kFlagEscaped << 16));
        }
      }
      __syncwarp();
      if (whole) {
        uint32_t st = 0;
        if (lane == 0) st = atomicAdd(&ctl->storeTicket, 1u);
        st = __shfl_sync(0xffffffffu, st, 0);
        const uint32_t sslot = st % kStoreSlots, gen = st / kStoreSlots;
        // the slot's previous bulk store must have finished reading it
        while (*reinterpret_cast<volatile uint32_t*>(&ctl->storeGen[sslot]) != gen) __nanosleep(32);
        float* dst = reinterpret_cast<float*>(storeRing + sslot * kStreamTileBytes) + lane * TR_WORDS;
#pragma unroll
        for (int i = 0; i < TR_WORDS; ++i) dst[i] = w[i];
        ac::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ac::bulk_store(a.rays + (size_t)first * TR_WORDS, storeRing + sslot * kStreamTileBytes, kStreamTileBytes);
          ac::bulk_commit();
          pendingSlot = sslot; pendingGen = gen;  // released once the copy has read the slot (retire_store)
          stored = true;
        }
      } else if (mine) {
        float* dst = a.rays + (size_t)(first + lane) * TR_WORDS;
#pragma unroll
        for (int i = 0; i < TR_WORDS; ++i) dst[i] = w[i];
      }
    }
    if (stored) ac::bulk_wait_all();  // this lane's bulk stores have reached global memory
  }
  __syncwarp();
  flush_counters(a.counters, nClosest, nOccl, cnt, 0u, 0u);
}

// -------------------------------------------------------------------------------------------------
// Equirectangular (u,v) of an escaped direction (PreProcessEscapedRays, codelets/TraceCodelets.cpp:337-348).
__device__ __forceinline__ void escaped_uv(V3 d, float rotation, float& u, float& v) {
  const float theta = acosf(d.y);
  float phi = atan2f(d.z, d.x) + rotation;
  const float twoPi = 6.28318530717958647692f;
  if (phi < 0.f) phi += twoPi;
  else if (phi > twoPi) phi -= twoPi;
  u = theta * 0.31830988618379067154f;
  v = phi * 0.15915494309189533577f;
}

// sampleCameraRays (codelets/TraceCodelets.cpp:142-164) with the per-(pixel,sample) stream: seeds `rng`, consumes the
// two jitter draws and returns the camera-ray direction. Shared by the pre-pass and the path tracer so that both see the
// same ray bit for bit.
__device__ __forceinline__ V3 camera_ray(const TraceArgs& a, float row, float col, uint32_t pixelIndex, uint32_t s, Rng& rng) {
  rng_seed_stream(rng, a.rngKey, pixelIndex, s);
  const uint64_t ra = rng_next(rng), rb = rng_next(rng);
  float g0, g1;
  gaussian_pair(ra, rb, g0, g1);
  const float pu = row + a.antiAlias * g0;
  const float pv = col + a.antiAlias * g1;
  return pixel_to_ray_dir(pv, pu, a.imageWidth, a.imageHeight, a.tanTheta);
}

// Primary-ray pre-pass. In the path tracer below every lane runs its own path, so a lane's camera ray is traced next to
// 31 unrelated bounce rays and the warp pays for the longest of them. Camera rays of neighbouring pixels are the one
// coherent population of the whole render: here a warp traces the same sample of 32 adjacent pixels together (~31 of 32
// lanes active) and parks the hit; the path tracer then starts each path from the parked hit instead of a traversal.
// Same ray, same traversal code, same hit -- results do not change, about a third of the path tracer's queries go away.
template <bool kShared, bool kOrdered, bool kCount>
__global__ void __launch_bounds__(768) primary_hit_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const uint2* nodes = stage_nodes<kShared>(a, reinterpret_cast<uint2*>(smemRaw));
  const DevScene& sc = a.scene;
  const unsigned lane = threadIdx.x & 31;
  const float inf = __int_as_float(0x7f800000);
  const uint32_t chunk = a.endSample - a.firstSample;
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0;
  while (true) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.workCounter + 1, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.numRays) break;
    const uint32_t idx = base + lane;
    if (idx >= a.numRays) continue;
    const float* tr = a.rays + (size_t)idx * TR_WORDS;
    const float row = tr[TR_ROW], col = tr[TR_COL];
    const uint32_t pixelIndex = (uint32_t)row * (uint32_t)a.imageWidth + (uint32_t)col;
    for (uint32_t s = a.firstSample; s < a.endSample; ++s) {
      Rng rng;
      const V3 d = camera_ray(a, row, col, pixelIndex, s, rng);
      const V3 o = offset_origin(mk(0.f, 0.f, 0.f), d, mk(0.f, 0.f, 1.f));
      Hit h;
      nClosest++;
      closest_hit<kShared, kOrdered, kCount>(sc, nodes, o, d, 0.f, inf, h, cnt);
      const size_t slot = (size_t)idx * chunk + (s - a.firstSample);
      a.primA[slot] = make_uint4(__float_as_uint(h.t), h.geomID, h.primID, h.tri);
      if (a.primB) a.primB[slot] = make_float4(h.b0, h.b1, h.b2, 0.f);
    }
  }
  flush_counters(a.counters, nClosest, 0u, cnt, 0u, 0u);
}

template <bool kShared, bool kOrdered, bool kCount, bool kNif>
__global__ void __launch_bounds__(768) path_trace_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const uint2* nodes = stage_nodes<kShared>(a, reinterpret_cast<uint2*>(smemRaw));
  const DevScene& sc = a.scene;
  const unsigned lane = threadIdx.x & 31;
  const float inf = __int_as_float(0x7f800000);
  const uint32_t chunk = a.endSample - a.firstSample;
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0, nSamples = 0, nEscaped = 0;

  while (true) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.workCounter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.numRays) break;
    const uint32_t idx = base + lane;
    if (idx >= a.numRays) continue;
    float* tr = a.rays + (size_t)idx * TR_WORDS;

    const float row = tr[TR_ROW], col = tr[TR_COL];
    const uint32_t pixelIndex = (uint32_t)row * (uint32_t)a.imageWidth + (uint32_t)col;
    V3 rgb = mk(tr[TR_RGB], tr[TR_RGB + 1], tr[TR_RGB + 2]);

    // per-path state (HitRecord fields + loop variables)
    V3 o, d, n, thr, color;
    float tMaxOut = inf;
    uint32_t primID = kInvalidPrim, geomID = kInvalidGeom, flags = 0;
    Rng rng;
    uint32_t s = a.firstSample;
    uint32_t bounce = 0;
    bool live = false;  // the lane holds a ray (o, d, n) that still has to be traced

    // One bounce of the loop in trace.cpp:125-184, from the hit `h` of the ray (o, d) onwards: hit point, emission,
    // BxDF sample, roulette, path end. Leaves the next ray in (o, d, n) with live = true, or ends the path (s++).
    auto shade = [&](const Hit& h) {
      bool ended = false, escaped = false;
      tMaxOut = h.t;
      if (h.geomID != kInvalidGeom) {
        geomID = h.geomID; primID = h.primID;
        o = o + d * h.t;
        n = hit_normal(sc, h, o);
        const Mat m = load_material(sc, geomID);
        if (m.emissive) color = color + thr * m.emission;
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
          // result.rgb *= NaN (trace.cpp:167). With an environment light rgb is summed afterwards from the slots, so
          // the poison has to travel in the path's colour (as in wf_shade_kernel)
          if (kNif) color = color * __int_as_float(0x7fc00000);
          else rgb = rgb * __int_as_float(0x7fc00000);
          flags |= kFlagError;
        }
      } else {
        flags |= kFlagEscaped;
        ended = true;
        escaped = true;
      }
      if (!ended) {
        if (bounce > a.rouletteStartDepth) {
          const float u1 = rng_uniform(rng);
          const float p = maxc(thr);  // evaluateRoulette (geometric_sampling.hpp:56-63)
          if (p == 0.f || u1 > p) ended = true;
          else thr = thr * (1.f / p);
        }
        bounce++;
        if (bounce >= a.maxPathLength) ended = true;
      }
      live = !ended;
      if (ended) {
        if (escaped) nEscaped++;
        if (kNif) {
          // wavefront hand-off: the environment light is evaluated by the NIF kernel afterwards and
          // rgb is accumulated in sample order by accumulate_kernel
          const size_t slot = (size_t)idx * chunk + (s - a.firstSample);
          float* sc3 = a.slotColor + 3 * slot;
          sc3[0] = color.x; sc3[1] = color.y; sc3[2] = color.z;
          float* se = a.slotEscape + 5 * slot;
          float u = -1.f, v = 0.f;
          if (escaped) escaped_uv(d, a.hdriRotation, u, v);
          se[0] = thr.x; se[1] = thr.y; se[2] = thr.z; se[3] = u; se[4] = v;
          // compact escaped slots (warp-aggregated append)
          const unsigned active = __activemask();
          const unsigned mask = __ballot_sync(active, escaped);
          if (escaped) {
            const int leader = __ffs(mask) - 1;
            uint32_t qbase = 0;
            if ((int)lane == leader) qbase = atomicAdd(a.escapeCount, (uint32_t)__popc(mask));
            qbase = __shfl_sync(mask, qbase, leader);
            a.escapeQueue[qbase + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)slot;
          }
        } else {
          rgb = rgb + color;  // result.rgb += color (trace.cpp:187)
        }
        s++;
      }
    };

    while (true) {
      // Regenerate: lanes whose path ended start their next sample(s) until they hold a ray to trace. With the
      // pre-pass the camera ray's hit is already parked, so its bounce is shaded right here and the lane joins the
      // traversal below with its first BOUNCE ray -- every trip through the traversal does real work in every lane.
      while (!live && s != a.endSample) {
        d = camera_ray(a, row, col, pixelIndex, s, rng);
        o = mk(0.f, 0.f, 0.f);
        n = mk(0.f, 0.f, 1.f);
        primID = kInvalidPrim; geomID = kInvalidGeom; flags = 0;
        thr = mk(1.f, 1.f, 1.f);
        color = mk(0.f, 0.f, 0.f);
        bounce = 0;
        nSamples++;
        if (a.primA != nullptr) {
          o = offset_origin(o, d, n);
          const size_t slot = (size_t)idx * chunk + (s - a.firstSample);
          const uint4 pa = a.primA[slot];
          Hit h;
          h.t = __uint_as_float(pa.x); h.geomID = pa.y; h.primID = pa.z; h.tri = pa.w; h.node = 0;
          h.b0 = h.b1 = h.b2 = 0.f;
          if (a.primB) { const float4 pb = a.primB[slot]; h.b0 = pb.x; h.b1 = pb.y; h.b2 = pb.z; }
          shade(h);
        } else {
          live = true;
        }
      }
      if (!live) break;

      // ---- one iteration of the bounce loop (trace.cpp:125-184) ----
      o = offset_origin(o, d, n);
      Hit h;
      nClosest++;
      closest_hit<kShared, kOrdered, kCount>(sc, nodes, o, d, 0.f, inf, h, cnt);
      shade(h);
    }

    // write back: rgb running sum + the HitRecord of the last sample (what the reference leaves behind)
    if (!kNif) { tr[TR_RGB] = rgb.x; tr[TR_RGB + 1] = rgb.y; tr[TR_RGB + 2] = rgb.z; }
    if (a.endSample > a.firstSample) {
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
  flush_counters(a.counters, nClosest, 0u, cnt, nSamples, nEscaped);
}

// Bare-ray queries for the parity tests (b200rt_intersect / b200rt_occluded).
struct QueryHit {
  float t;
  uint32_t geomID, primID;
  float nx, ny, nz;
};
template <bool kOrdered>
__global__ void intersect_kernel(DevScene sc, const float* rays, uint32_t n, QueryHit* out, DeviceCounters* counters) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  Counters cnt = {0u, 0u};
  unsigned nq = 0;
  if (idx < n) {
    const float* r = rays + 8 * (size_t)idx;
    const V3 o = mk(r[0], r[1], r[2]), d = mk(r[4], r[5], r[6]);
    Hit h;
    nq = 1;
    if (kOrdered) {  // near-first over the pair table, as the render kernels traverse
      uint2 stack[kMaxStack];
      PairHit ph;
      pair_closest_hit<false, true>(sc, sc.pairs, o, d, r[3], r[7], ph, stack, cnt.nodeVisits, cnt.primTests);
      h.t = ph.t; h.geomID = ph.geomID; h.b0 = ph.b0; h.b1 = ph.b1; h.b2 = ph.b2; h.node = 0;
      hit_ids(sc, ph, h.primID, h.tri);
    } else {
      closest_hit_ref_order<false, true>(sc, sc.nodes, o, d, r[3], r[7], h, cnt);
    }
    QueryHit q;
    q.t = h.t; q.geomID = h.geomID; q.primID = h.primID; q.nx = q.ny = q.nz = 0.f;
    if (h.geomID != kInvalidGeom) {
      const V3 nn = hit_normal(sc, h, o + d * h.t);
      q.nx = nn.x; q.ny = nn.y; q.nz = nn.z;
    }
    out[idx] = q;
  }
  flush_counters(counters, nq, 0u, cnt, 0u, 0u);
}
__global__ void occluded_kernel(DevScene sc, const float* rays, uint32_t n, unsigned char* out) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float* r = rays + 8 * (size_t)idx;
  Counters cnt = {0u, 0u};
  out[idx] = any_hit<false, false>(sc, sc.nodes, mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3], r[7], cnt) ? 1 : 0;
}

}  // namespace rt
