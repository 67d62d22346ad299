// path_trace_sm_kernel — the path tracer as a warp-scheduled state machine.
//
// Same per-lane arithmetic as path_trace_kernel (bit-identical results), different control flow. The plain
// kernel nests "for each bounce { traverse; shade }" per lane; inside a warp that executes the leaf test, the
// inner-node step and the shading code every iteration for whichever lanes happen to need them, and a whole warp
// waits for its slowest traversal: ncu measured 8.5 of 32 lanes active per issued instruction. Here each lane
// carries an explicit phase
//     TRAV  expand one inner node (fetch + slab-test both children, descend to the nearer, defer the farther)
//     LEAF  intersect one leaf primitive, then pop
//     SHADE closest hit known: updateHit, material, BxDF sample, roulette -> next ray or end of path
//     REGEN start the next sample of the lane's pixel, or fetch the lane's next pixel
// and every warp iteration runs ONE phase, chosen by ballot (inner-node steps while enough lanes want them,
// otherwise the phase with the most waiting lanes). Lanes never wait for another lane's traversal to finish;
// rare, expensive phases (leaf tests, shading) run when many lanes have queued up for them.
// Work distribution is per lane: a lane that finishes its pixel's samples takes the next unclaimed ray index.
#pragma once
#include "trace_kernels.cuh"

namespace rt {

enum : int { PH_TRAV = 0, PH_LEAF = 1, PH_SHADE = 2, PH_REGEN = 3, PH_IDLE = 4 };

template <bool kShared, bool kCount, bool kNif>
__global__ void __launch_bounds__(768) path_trace_sm_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const uint2* nodes = stage_nodes<kShared>(a, reinterpret_cast<uint2*>(smemRaw));
  const DevScene& sc = a.scene;
  const unsigned lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  const uint32_t chunk = a.endSample - a.firstSample;
  Counters cnt = {0u, 0u};
  unsigned nClosest = 0, nSamples = 0, nEscaped = 0;

  // ---- per-lane path state (HitRecord fields + loop variables) ----
  V3 o = mk(0.f, 0.f, 0.f), d = mk(0.f, 0.f, -1.f), n = mk(0.f, 0.f, 1.f);
  V3 thr = mk(1.f, 1.f, 1.f), color = mk(0.f, 0.f, 0.f), rgb = mk(0.f, 0.f, 0.f);
  float tMaxOut = inf, row = 0.f, col = 0.f;
  uint32_t primID = kInvalidPrim, geomID = kInvalidGeom, flags = 0;
  uint32_t pixelIndex = 0, idx = 0xFFFFFFFFu, s = a.endSample, bounce = 0;
  Rng rng;
  rng.s0 = rng.s1 = 0;
  // ---- per-lane traversal state ----
  V3 inv = mk(0.f, 0.f, 0.f);
  Shear sh;
  sh.kz = 2; sh.sx = sh.sy = sh.sz = 0.f;
  Hit hit;
  hit.t = inf; hit.geomID = kInvalidGeom; hit.primID = kInvalidPrim; hit.tri = 0; hit.node = 0;
  hit.b0 = hit.b1 = hit.b2 = 0.f;
  uint32_t cur = 0, meta = 0, curGeom = kInvalidGeom;
  uint32_t stackIdx[kMaxStack];
  float stackEnter[kMaxStack];
  int sp = 0;
  int phase = PH_REGEN;

  // pop the next deferred node that can still contain a closer hit; sets the phase
  auto pop_next = [&]() {
    bool found = false;
    while (sp > 0) {
      --sp;
      if (!(stackEnter[sp] > hit.t)) { found = true; break; }
    }
    if (!found) { phase = PH_SHADE; return; }
    cur = stackIdx[sp];
    const uint2* p = nodes + 3u * cur;
    if (kShared) { meta = p[1].y; curGeom = p[2].y >> 16; }
    else { meta = __ldg(p + 1).y; curGeom = __ldg(p + 2).y >> 16; }
    phase = curGeom != kInvalidGeom ? PH_LEAF : PH_TRAV;
  };

  // offsetRay + start of CompactBvh::intersect for the lane's current ray (o, d)
  auto begin_query = [&]() {
    o = offset_origin(o, d, n);
    nClosest++;
    inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
    sh = make_shear(d);
    hit.t = inf; hit.geomID = kInvalidGeom; hit.primID = kInvalidPrim; hit.tri = 0; hit.node = 0;
    hit.b0 = hit.b1 = hit.b2 = 0.f;
    sp = 0;
    const NodeWords w = fetch_node<kShared>(nodes, 0);
    if (kCount) cnt.nodeVisits++;
    float enter;
    if (!slab_test(w, o, inv, 0.f, hit.t, enter)) { phase = PH_SHADE; return; }
    cur = 0; meta = w.b.y; curGeom = w.c.y >> 16;
    phase = curGeom != kInvalidGeom ? PH_LEAF : PH_TRAV;
  };

  while (true) {
    const unsigned mT = __ballot_sync(full, phase == PH_TRAV);
    const unsigned mL = __ballot_sync(full, phase == PH_LEAF);
    const unsigned mS = __ballot_sync(full, phase == PH_SHADE);
    const unsigned mR = __ballot_sync(full, phase == PH_REGEN);
    if (!(mT | mL | mS | mR)) break;
    const int cT = __popc(mT), cL = __popc(mL), cS = __popc(mS), cR = __popc(mR);
    int pick;
    if (cT >= a.travThreshold || (cT >= cL && cT >= cS && cT >= cR)) pick = PH_TRAV;
    else if (cL >= cS && cL >= cR) pick = PH_LEAF;
    else if (cS >= cR) pick = PH_SHADE;
    else pick = PH_REGEN;

    if (pick == PH_TRAV) {
      if (phase == PH_TRAV) {
        // ---- expand one inner node (see closest_hit_ordered) ----
        const uint32_t c0 = cur + 1, c1 = meta;
        const NodeWords w0 = fetch_node<kShared>(nodes, c0);
        const NodeWords w1 = fetch_node<kShared>(nodes, c1);
        if (kCount) cnt.nodeVisits += 2;
        float e0, e1;
        const bool h0 = slab_test(w0, o, inv, 0.f, hit.t, e0);
        const bool h1 = slab_test(w1, o, inv, 0.f, hit.t, e1);
        if (h0 && h1) {
          const bool firstNear = !(e1 < e0);  // ties go to the first child, like pre-order
          stackIdx[sp] = firstNear ? c1 : c0;
          stackEnter[sp] = firstNear ? e1 : e0;
          sp++;
          cur = firstNear ? c0 : c1;
          meta = firstNear ? w0.b.y : w1.b.y;
          curGeom = (firstNear ? w0.c.y : w1.c.y) >> 16;
          if (curGeom != kInvalidGeom) phase = PH_LEAF;
        } else if (h0) {
          cur = c0; meta = w0.b.y; curGeom = w0.c.y >> 16;
          if (curGeom != kInvalidGeom) phase = PH_LEAF;
        } else if (h1) {
          cur = c1; meta = w1.b.y; curGeom = w1.c.y >> 16;
          if (curGeom != kInvalidGeom) phase = PH_LEAF;
        } else {
          pop_next();
        }
      }
    } else if (pick == PH_LEAF) {
      if (phase == PH_LEAF) {
        if (kCount) cnt.primTests++;
        const LeafResult r = leaf_test(sc, curGeom, meta, o, d, 0.f, sh);
        if (r.t > 0.f && (r.t < hit.t || (r.t == hit.t && hit.geomID != kInvalidGeom && cur < hit.node))) {
          hit.t = r.t; hit.geomID = curGeom; hit.primID = sc.geoms[curGeom].type == 0 ? meta : 0u;
          hit.tri = r.tri; hit.node = cur; hit.b0 = r.b0; hit.b1 = r.b1; hit.b2 = r.b2;
        }
        pop_next();
      }
    } else if (pick == PH_SHADE) {
      uint32_t appendSlot = 0xFFFFFFFFu;  // NIF wavefront: slot of a sample that escaped in this step
      if (phase == PH_SHADE) {
        // ---- rest of one bounce-loop iteration (trace.cpp:133-184) ----
        bool ended = false, escaped = false;
        tMaxOut = hit.t;
        if (hit.geomID != kInvalidGeom) {
          geomID = hit.geomID; primID = hit.primID;
          o = o + d * hit.t;
          n = hit_normal(sc, hit, o);
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
            rgb = rgb * __int_as_float(0x7fc00000);  // result.rgb *= NaN (trace.cpp:167)
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
        if (!ended) {
          begin_query();
        } else {
          if (escaped) nEscaped++;
          if (kNif) {
            const size_t slot = (size_t)idx * chunk + (s - a.firstSample);
            float* sc3 = a.slotColor + 3 * slot;
            sc3[0] = color.x; sc3[1] = color.y; sc3[2] = color.z;
            float* se = a.slotEscape + 5 * slot;
            float u = -1.f, v = 0.f;
            if (escaped) escaped_uv(d, a.hdriRotation, u, v);
            se[0] = thr.x; se[1] = thr.y; se[2] = thr.z; se[3] = u; se[4] = v;
            if (escaped) appendSlot = (uint32_t)slot;
          } else {
            rgb = rgb + color;  // result.rgb += color (trace.cpp:187)
          }
          s++;
          phase = PH_REGEN;
        }
      }
      if (kNif) {
        // compact the escaped slots of this step (warp-aggregated append; all lanes are converged here)
        const unsigned mask = __ballot_sync(full, appendSlot != 0xFFFFFFFFu);
        if (mask) {
          const int leader = __ffs(mask) - 1;
          uint32_t qbase = 0;
          if ((int)lane == leader) qbase = atomicAdd(a.escapeCount, (uint32_t)__popc(mask));
          qbase = __shfl_sync(full, qbase, leader);
          if (appendSlot != 0xFFFFFFFFu) a.escapeQueue[qbase + __popc(mask & ((1u << lane) - 1u))] = appendSlot;
        }
      }
    } else {
      // lanes that have finished their pixel claim the next ray indices together (all lanes are converged here)
      const unsigned want = __ballot_sync(full, phase == PH_REGEN && s == a.endSample);
      uint32_t claimBase = 0;
      if (want) {
        const int leader = __ffs(want) - 1;
        if ((int)lane == leader) claimBase = atomicAdd(a.workCounter, (uint32_t)__popc(want));
        claimBase = __shfl_sync(full, claimBase, leader);
      }
      if (phase == PH_REGEN) {
        if (s == a.endSample) {
          // pixel finished: write back (rgb running sum + the HitRecord of the last sample), take the next ray
          if (idx != 0xFFFFFFFFu) {
            float* tr = a.rays + (size_t)idx * TR_WORDS;
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
          idx = claimBase + (uint32_t)__popc(want & ((1u << lane) - 1u));
          if (idx >= a.numRays || a.endSample == a.firstSample) {
            idx = 0xFFFFFFFFu;
            phase = PH_IDLE;
          } else {
            const float* tr = a.rays + (size_t)idx * TR_WORDS;
            row = tr[TR_ROW]; col = tr[TR_COL];
            pixelIndex = (uint32_t)row * (uint32_t)a.imageWidth + (uint32_t)col;
            rgb = mk(tr[TR_RGB], tr[TR_RGB + 1], tr[TR_RGB + 2]);
            s = a.firstSample;
          }
        }
        if (phase == PH_REGEN) {
          // sampleCameraRays (codelets/TraceCodelets.cpp:142-164) with the per-(pixel,sample) stream
          rng_seed_stream(rng, a.rngKey, pixelIndex, s);
          const uint64_t ra = rng_next(rng), rb = rng_next(rng);
          float g0, g1;
          gaussian_pair(ra, rb, g0, g1);
          const float pu = row + a.antiAlias * g0;
          const float pv = col + a.antiAlias * g1;
          d = pixel_to_ray_dir(pv, pu, a.imageWidth, a.imageHeight, a.tanTheta);
          o = mk(0.f, 0.f, 0.f);
          n = mk(0.f, 0.f, 1.f);
          primID = kInvalidPrim; geomID = kInvalidGeom; flags = 0;
          thr = mk(1.f, 1.f, 1.f);
          color = mk(0.f, 0.f, 0.f);
          bounce = 0;
          nSamples++;
          begin_query();
        }
      }
    }
  }
  flush_counters(a.counters, nClosest, 0u, cnt, nSamples, nEscaped);
}

}  // namespace rt
