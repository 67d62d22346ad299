// Device-side building blocks of the trace path: compact-BVH node fetch + slab test, the three
// primitive tests, closest-hit / any-hit traversal, ray offsetting and the BxDFs.
//
// Numerical contract: compiled with --fmad=false, IEEE div/sqrt, no FTZ, so each function below
// produces the same bits as the reference function it cites when given the same inputs. Min/max
// are written as the reference's ternaries (NOT fminf/fmaxf) to keep its NaN/inf behaviour
// (SURVEY.md §7 "NaN/inf semantics in the slab test").
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "rt_prims.h"

namespace rt {

struct Hit {
  float t;          // closest t so far (starts at ray tMax)
  uint32_t geomID;  // kInvalidGeom when nothing hit
  uint32_t primID;  // primitive id as the reference reports it (triangle index within its mesh, or 0)
  uint32_t tri;     // global triangle index (meshes only)
  uint32_t node;    // BVH node index of the winning leaf (tie-break key)
  float b0, b1, b2; // barycentrics of the winning triangle
};

struct Counters {
  uint32_t nodeVisits, primTests;
};

// ---------------------------------------------------------------------------------------------
// Node access. A node is {f32 min[3]; u32 primOrSecond; f16 d[3]; u16 geomID}.
struct NodeWords {
  uint2 a, b, c;  // a = (min_x, min_y)  b = (min_z, primOrSecond)  c = (dx|dy<<16, dz|geomID<<16)
};

template <bool kShared>
__device__ __forceinline__ NodeWords fetch_node(const uint2* __restrict__ nodes, uint32_t idx) {
  NodeWords w;
  const uint2* p = nodes + 3u * idx;
  if (kShared) {
    w.a = p[0]; w.b = p[1]; w.c = p[2];
  } else {
    w.a = __ldg(p); w.b = __ldg(p + 1); w.c = __ldg(p + 2);
  }
  return w;
}

// CompactBVH2Node::intersect + intersectRaySlab x3 (src/CompactBVH2Node.cpp:5-22,
// include/CompactBVH2Node.hpp:36-48). Evaluated without the per-axis early-outs: t0 only grows and
// t1 only shrinks, so the final comparison equals the early-out result for every input incl. NaN/inf.
// Returns the accumulated entry distance in `enter`.
__device__ __forceinline__ bool slab_test(const NodeWords& w, V3 o, V3 inv, float tMin, float tLimit, float& enter) {
  return slab_exact(__uint_as_float(w.a.x), __uint_as_float(w.a.y), __uint_as_float(w.b.x), w.c.x, w.c.y, o, inv, tMin, tLimit, enter);
}

// One leaf: dispatch on the geometry type (primLookup + virtual Primitive::intersect,
// codelets/TraceCodelets.cpp:127-140). Returns the t the reference's Intersection would carry:
// +inf for a missed triangle (Mesh.hpp:90), 0 for a missed sphere/disc (Intersection::Failed()).
struct LeafResult {
  float t;
  uint32_t tri;
  float b0, b1, b2;
};
__device__ __forceinline__ LeafResult leaf_test(const DevScene& sc, uint32_t geomID, uint32_t primID, V3 o, V3 d,
                                                float tMin, const Shear& sh) {
  LeafResult r;
  r.tri = 0; r.b0 = r.b1 = r.b2 = 0.f;
  const GeomEntry g = sc.geoms[geomID];
  if (g.type == 0) {
    r.tri = g.first + primID;
    const float4* tv = sc.triVerts + 3u * r.tri;
    const float4 a = __ldg(tv), b = __ldg(tv + 1), c = __ldg(tv + 2);
    const float t = tri_test(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), o, sh, r.b0, r.b1, r.b2);
    r.t = t > 0.f ? t : __int_as_float(0x7f800000);  // Mesh.hpp:93: only t>0 && t<inf replaces the inf default
  } else if (g.type == 1) {
    r.t = sphere_test(__ldg(sc.spheres + g.first), o, d, tMin);
  } else {
    r.t = disc_test(sc.discs + 7u * g.first, o, d);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Closest hit, REFERENCE ORDER: pre-order DFS, first child first, no near/far ordering; a hit is
// accepted iff t > tMin && t < closest (CompactBvh.hpp:80-139). The "push second, push first, pop
// first" of the reference is folded into "push second, continue with first".
template <bool kShared, bool kCount>
__device__ __forceinline__ void closest_hit_ref_order(const DevScene& sc, const uint2* __restrict__ nodes, V3 o, V3 d,
                                                      float tMin, float tMax, Hit& hit, Counters& cnt) {
  const V3 inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  hit.t = tMax; hit.geomID = kInvalidGeom; hit.primID = kInvalidPrim; hit.tri = 0; hit.node = 0;
  hit.b0 = hit.b1 = hit.b2 = 0.f;
  uint32_t stack[kMaxStack];
  int sp = 0;
  uint32_t cur = 0;
  while (true) {
    const NodeWords w = fetch_node<kShared>(nodes, cur);
    if (kCount) cnt.nodeVisits++;
    float enter;
    bool descend = false;
    if (slab_test(w, o, inv, tMin, hit.t, enter)) {
      const uint32_t geomID = w.c.y >> 16;
      if (geomID != kInvalidGeom) {
        if (kCount) cnt.primTests++;
        const LeafResult r = leaf_test(sc, geomID, w.b.y, o, d, tMin, sh);
        if (r.t > tMin && r.t < hit.t) {
          hit.t = r.t; hit.geomID = geomID; hit.primID = sc.geoms[geomID].type == 0 ? w.b.y : 0u;
          hit.tri = r.tri; hit.node = cur; hit.b0 = r.b0; hit.b1 = r.b1; hit.b2 = r.b2;
        }
      } else {
        stack[sp++] = w.b.y;  // second child waits
        cur = cur + 1;        // first child is the next node in the array
        descend = true;
      }
    }
    if (!descend) {
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
}

// Closest hit, NEAR-FIRST ORDER. Same node tests, same primitive tests, same acceptance window;
// only the visiting order changes (nearer child first, farther child deferred with its entry
// distance and re-checked against the shrunken closest-t when popped, which is exactly the
// reference's pop-time slab test because enter <= boxExit was already established). Equal-t ties are
// resolved to the LOWEST leaf node index, i.e. to the leaf the reference's pre-order walk meets
// first, so the reported primitive is the reference's.
template <bool kShared, bool kCount>
__device__ __forceinline__ void closest_hit_ordered(const DevScene& sc, const uint2* __restrict__ nodes, V3 o, V3 d,
                                                    float tMin, float tMax, Hit& hit, Counters& cnt) {
  const V3 inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  hit.t = tMax; hit.geomID = kInvalidGeom; hit.primID = kInvalidPrim; hit.tri = 0; hit.node = 0;
  hit.b0 = hit.b1 = hit.b2 = 0.f;
  // one 8-byte local-memory record per deferred node: {node index, entry distance}
  uint2 stack[kMaxStack];
  int sp = 0;

  NodeWords w = fetch_node<kShared>(nodes, 0);
  if (kCount) cnt.nodeVisits++;
  float enter;
  if (!slab_test(w, o, inv, tMin, hit.t, enter)) return;
  constexpr uint32_t kDone = 0xFFFFFFFFu;
  uint32_t cur = 0;
  uint32_t meta = w.b.y, geomID = w.c.y >> 16;  // of `cur`

  // pop the next deferred node that can still contain a closer hit (the reference's pop-time slab test)
  auto pop = [&]() {
    cur = kDone;
    while (sp > 0) {
      const uint2 e = stack[--sp];
      if (!(__uint_as_float(e.y) > hit.t)) { cur = e.x; break; }
    }
    if (cur != kDone) {
      const uint2* p = nodes + 3u * cur;
      if (kShared) { meta = p[1].y; geomID = p[2].y >> 16; }
      else { meta = __ldg(p + 1).y; geomID = __ldg(p + 2).y >> 16; }
    }
  };

  // "while-while" (Aila & Laine): every lane first descends inner nodes until it HOLDS a leaf (or is done); the
  // leaf test then runs for all lanes of the warp that hold one, instead of for the one or two lanes that happen
  // to reach a leaf in a given iteration (ncu: the triangle test ran with 3-6 of 32 lanes active before).
  while (cur != kDone) {
    while (cur != kDone && geomID == kInvalidGeom) {
      const uint32_t c0 = cur + 1, c1 = meta;
      const NodeWords w0 = fetch_node<kShared>(nodes, c0);
      const NodeWords w1 = fetch_node<kShared>(nodes, c1);
      if (kCount) cnt.nodeVisits += 2;
      float e0, e1;
      const bool h0 = slab_test(w0, o, inv, tMin, hit.t, e0);
      const bool h1 = slab_test(w1, o, inv, tMin, hit.t, e1);
      if (h0 && h1) {
        const bool firstNear = !(e1 < e0);  // ties go to the first child, like pre-order
        stack[sp++] = make_uint2(firstNear ? c1 : c0, __float_as_uint(firstNear ? e1 : e0));
        cur = firstNear ? c0 : c1;
        meta = firstNear ? w0.b.y : w1.b.y;
        geomID = (firstNear ? w0.c.y : w1.c.y) >> 16;
      } else if (h0) {
        cur = c0; meta = w0.b.y; geomID = w0.c.y >> 16;
      } else if (h1) {
        cur = c1; meta = w1.b.y; geomID = w1.c.y >> 16;
      } else {
        pop();
      }
    }
    if (cur == kDone) break;
    if (kCount) cnt.primTests++;
    const LeafResult r = leaf_test(sc, geomID, meta, o, d, tMin, sh);
    if (r.t > tMin && (r.t < hit.t || (r.t == hit.t && hit.geomID != kInvalidGeom && cur < hit.node))) {
      hit.t = r.t; hit.geomID = geomID; hit.primID = sc.geoms[geomID].type == 0 ? meta : 0u;
      hit.tri = r.tri; hit.node = cur; hit.b0 = r.b0; hit.b1 = r.b1; hit.b2 = r.b2;
    }
    pop();
  }
}

// Any hit (CompactBvh::occluded, CompactBvh.hpp:33-78): the node window stays [tMin, tMax], a
// primitive occludes iff tMin < t < tMax. The answer does not depend on visiting order.
template <bool kShared, bool kCount>
__device__ __forceinline__ bool any_hit(const DevScene& sc, const uint2* __restrict__ nodes, V3 o, V3 d, float tMin,
                                        float tMax, Counters& cnt) {
  const V3 inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  uint32_t stack[kMaxStack];
  int sp = 0;
  uint32_t cur = 0;
  while (true) {
    const NodeWords w = fetch_node<kShared>(nodes, cur);
    if (kCount) cnt.nodeVisits++;
    float enter;
    bool descend = false;
    if (slab_test(w, o, inv, tMin, tMax, enter)) {
      const uint32_t geomID = w.c.y >> 16;
      if (geomID != kInvalidGeom) {
        if (kCount) cnt.primTests++;
        const LeafResult r = leaf_test(sc, geomID, w.b.y, o, d, tMin, sh);
        if (r.t > tMin && r.t < tMax) return true;
      } else {
        stack[sp++] = w.b.y;
        cur = cur + 1;
        descend = true;
      }
    }
    if (!descend) {
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// Shading-side helpers.

// Primitive::normal at the updated hit point (see prim_normal in rt_prims.h).
__device__ __forceinline__ V3 hit_normal(const DevScene& sc, const Hit& h, V3 hitPoint) {
  return prim_normal(sc, h.geomID, h.tri, h.b0, h.b1, h.b2, hitPoint);
}

// offsetRay (Render.hpp:29-33)
__device__ __forceinline__ V3 offset_origin(V3 o, V3 d, V3 n) {
  const float m = (1.f + maxc(vabs(o))) * kRayEpsilon * copysignf(1.f, dot(n, d));
  return o + n * m;
}

struct Mat {
  V3 albedo;
  float ior;
  V3 emission;
  int type;
  bool emissive;
};
__device__ __forceinline__ Mat load_material(const DevScene& sc, uint32_t geomID) {
  const float* m = sc.materials + 9u * __ldg(sc.matIDs + geomID);
  Mat r;
  r.albedo = mk(__ldg(m), __ldg(m + 1), __ldg(m + 2));
  r.ior = __ldg(m + 3);
  r.emission = mk(__ldg(m + 4), __ldg(m + 5), __ldg(m + 6));
  r.type = __float_as_int(__ldg(m + 7));
  r.emissive = (__float_as_uint(__ldg(m + 8)) & 0xffu) != 0u;
  return r;
}

// sampleDiffuse (BxDF.hpp:11-30) = orthonormalSystem (geometry.hpp:147-159) +
// cosineSampleHemisphere / sampleDiscConcentric (geometric_sampling.hpp:8-45).
__device__ __forceinline__ V3 sample_diffuse(V3 n, float u1, float u2) {
  V3 xb;
  const V3 a = vabs(n), sq = n * n;
  if (a.x > a.y) {
    const float il = 1.f / sqrtf(sq.x + sq.z);
    xb = mk(-n.z * il, 0.f, n.x * il);
  } else {
    const float il = 1.f / sqrtf(sq.y + sq.z);
    xb = mk(0.f, n.z * il, -n.y * il);
  }
  const V3 yb = cross(n, xb);
  const float ux = 2.f * u1 - 1.f, uy = 2.f * u2 - 1.f;
  float px, py;
  if (ux == 0.f && uy == 0.f) {
    px = ux; py = uy;
  } else {
    float r, th;
    if (fabsf(ux) > fabsf(uy)) { r = ux; th = 0.78539816339744830962f * (uy / ux); }
    else { r = uy; th = 1.57079632679489661923f - 0.78539816339744830962f * (ux / uy); }
    float s, c;
    sincos_tbl(th, s, c);
    px = r * c; py = r * s;
  }
  const float zz = 1.f - px * px - py * py;
  const float pz = sqrtf(0.f < zz ? zz : 0.f);
  const V3 wi = mk(px, py, pz);
  return mk(dot(mk(xb.x, yb.x, n.x), wi), dot(mk(xb.y, yb.y, n.y), wi), dot(mk(xb.z, yb.z, n.z), wi));
}

// reflect (BxDF.hpp:33-37)
__device__ __forceinline__ V3 reflect_dir(V3 d, V3 n) {
  const float c = dot(d, n);
  return normalized(d - n * (c * 2.f));
}

// dielectric = schlick + refract + reflect (BxDF.hpp:39-75)
__device__ __forceinline__ V3 dielectric_dir(V3 d, V3 n, float ri, float u1, bool& refracted) {
  if (dot(n, d) > 0.f) n = -n; else ri = 1.f / ri;
  const float ndotr = dot(n, d);
  const float cost1 = -ndotr;
  const float cost2 = 1.f - ri * ri * (1.f - cost1 * cost1);
  bool doRefract = false;
  if (cost2 > 0.f) {
    float r0 = (1.f - ri) / (1.f + ri);
    r0 = r0 * r0;
    const float base = 1.f - cost1;
    const float base2 = base * base;
    const float base5 = base2 * base * base2;
    doRefract = u1 > r0 + (1.f - r0) * base5;
  }
  refracted = doRefract;
  if (doRefract) {
    const V3 rPerp = (d + n * cost1) * ri;
    const V3 rPar = n * -sqrtf(fabsf(1.f - norm2(rPerp)));
    return rPerp + rPar;
  }
  return reflect_dir(d, n);
}

}  // namespace rt
