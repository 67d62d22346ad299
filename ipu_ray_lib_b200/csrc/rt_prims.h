// Host+device building blocks of one BVH query: the slab test, the three primitive tests, and the
// derived two-children ("pair") node table that the streaming kernels traverse.
//
// Everything here is RT_HD: nvcc (--fmad=false, IEEE div/sqrt, no FTZ) and g++ (-ffp-contract=off)
// produce the same bits, so the traversal core is checked on the CPU against the oracle
// (tests/host_pair_check.cpp) before it ever runs on a GPU.
//
// Numerical contract: each function produces the same bits as the reference function it cites when
// given the same inputs.
#pragma once
#include <stdint.h>
#include <vector_types.h>
#include <vector_functions.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#include "rt_math.h"

#if defined(__CUDA_ARCH__)
#define RT_LDG(p) __ldg(p)
#else
#define RT_LDG(p) (*(p))
#endif

namespace rt {

constexpr uint32_t kInvalidGeom = 0xFFFFu;
constexpr uint32_t kInvalidPrim = 0xFFFFFFFFu;
constexpr int kMaxStack = 64;  // the builder (like Embree's, bvh.hpp:52) bounds depth at 64

// geomID -> what to intersect. Built on the host at scene creation from GeomRef[] + MeshInfo[]
// (include/Scene.hpp:27-32, include/Mesh.hpp:15-20) so a leaf needs one lookup instead of two.
struct alignas(8) GeomEntry {
  uint32_t type;   // 0 mesh, 1 sphere, 2 disc
  uint32_t first;  // mesh: global index of its first triangle; sphere/disc: index into that array
};

// Read-only scene view handed to every kernel by value.
struct DevScene {
  const uint2* nodes;        // CompactBVH2Node[], 24 B each, UNCHANGED from the caller (read as 3 x 8 B)
  const GeomEntry* geoms;    // [num_geometry]
  const float4* triVerts;    // [3][num_tris][3]  p0,p1,p2 gathered from Triangle[] + Vec3fa[] (w unused); copy 0 as
                             // given, copies 1 and 2 with the components rotated for kz = 0 / kz = 1 (see tri_test_fast)
  const float4* triNormals;  // [num_tris][3]  vertex normals, or nullptr when the scene has none
  const float4* triFaceNormals;  // [num_tris] normalized(cross(p1 - p0, p2 - p0)), computed once at scene creation (host and
                             // device arithmetic agree bit for bit, see rt_math.h); nullptr = compute from the vertices
  const float4* spheres;     // {x,y,z,radius}
  const float* discs;        // {nx,ny,nz,r,cx,cy,cz}
  const uint32_t* matIDs;    // [num_geometry]
  const float* materials;    // Material[], 9 words each (36 B)
  uint32_t numNodes;
  uint32_t numMaterials;
  // derived pair table (see "Pair table" below)
  const uint4* pairs;        // [numPairs][3]
  const uint32_t* leafOrig;  // [numPairs][2] reference node index of each child (equal-t tie-breaks only)
  uint32_t numPairs;
  uint32_t rootRef;          // root as a child reference: pair index 0, or the leaf reference when the tree is one leaf
  uint32_t rootGeom;         // geomID of the root (kInvalidGeom when it is an inner node)
  uint32_t boundsFinite;     // 1 = every node bound is finite: the NaN-free fast slab test is allowed
  // per-primitive table of the streaming kernels, indexed by prim_slot(leaf reference)
  const uint4* leafInfo;     // {geomID, primID as the reference reports it, reference node index of its (first) leaf, -}
  uint32_t numTris, numSpheres;
  uint32_t trisBounded;      // 1 = every triangle vertex is finite and |coordinate| < 2^20: tri_test_fast is allowed
};

RT_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } b;
  b.u = u;
  return b.f;
#endif
}
RT_HD uint32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } b;
  b.f = f;
  return b.u;
#endif
}

RT_HD float half_bits_to_float(uint32_t h16) {
#if defined(__CUDA_ARCH__)
  return __half2float(__ushort_as_half((unsigned short)h16));
#else
  union { _Float16 h; uint16_t u; } b;
  b.u = (uint16_t)h16;
  return (float)b.h;  // exact widening
#endif
}

// ---------------------------------------------------------------------------------------------
// One axis of intersectRaySlab (include/CompactBVH2Node.hpp:36-48), the reference's ternaries kept
// so that NaN/inf behave exactly as there.
RT_HD void slab_axis_exact(float mn, float ext, float o, float inv, float& t0, float& t1) {
  const float mx = mn + ext;
  float tmin = (mn - o) * inv, tmax = (mx - o) * inv;
  if (tmin > tmax) { const float s = tmin; tmin = tmax; tmax = s; }
  tmax *= kSlabGuard;
  t0 = tmin > t0 ? tmin : t0;
  t1 = tmax < t1 ? tmax : t1;
}
// CompactBVH2Node::intersect (src/CompactBVH2Node.cpp:5-22) on unpacked fields, evaluated without the
// per-axis early-outs: t0 only grows and t1 only shrinks, so the final comparison equals the early-out
// result for every input incl. NaN/inf. Returns the accumulated entry distance in `enter`.
RT_HD bool slab_exact(float mnx, float mny, float mnz, uint32_t extXY, uint32_t extZ, V3 o, V3 inv, float tMin,
                      float tLimit, float& enter) {
  float t0 = tMin, t1 = tLimit;
  slab_axis_exact(mnx, half_bits_to_float(extXY & 0xffffu), o.x, inv.x, t0, t1);
  slab_axis_exact(mny, half_bits_to_float(extXY >> 16), o.y, inv.y, t0, t1);
  slab_axis_exact(mnz, half_bits_to_float(extZ & 0xffffu), o.z, inv.z, t0, t1);
  enter = t0;
  return !(t0 > t1);
}

// The same test for the common case in which NO NaN can occur: every node bound finite, the ray origin
// finite, 1/d finite on all three axes (so d != 0 and not denormal-small) and tMin/tLimit not NaN. Then
// (bound - o) * inv is finite or +-inf, the two plane distances of an axis are ordered, and
//   swap-if-greater        == (min, max)
//   t0 = tmin > t0 ? ..    == max(tmin, t0)       (t0, t1 are never NaN)
//   !(t0 > t1)             == t0 <= t1
// up to the sign of a zero, which only ever feeds comparisons. min/max map to FMNMX/FMNMX3 on sm_100a:
// 32 instructions per box instead of 46 with the NaN-preserving select chains.
RT_HD bool slab_fast(float mnx, float mny, float mnz, uint32_t extXY, uint32_t extZ, V3 o, V3 inv, float tMin,
                     float tLimit, float& enter) {
  const float mxx = mnx + half_bits_to_float(extXY & 0xffffu);
  const float mxy = mny + half_bits_to_float(extXY >> 16);
  const float mxz = mnz + half_bits_to_float(extZ & 0xffffu);
  const float ax = (mnx - o.x) * inv.x, bx = (mxx - o.x) * inv.x;
  const float ay = (mny - o.y) * inv.y, by = (mxy - o.y) * inv.y;
  const float az = (mnz - o.z) * inv.z, bz = (mxz - o.z) * inv.z;
  const float hx = fmaxf(ax, bx) * kSlabGuard, hy = fmaxf(ay, by) * kSlabGuard, hz = fmaxf(az, bz) * kSlabGuard;
  const float t0 = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tMin);
  const float t1 = fminf(fminf(fminf(hx, hy), hz), tLimit);
  enter = t0;
  return t0 <= t1;
}
RT_HD bool is_finite_f(float x) { return fabsf(x) < __builtin_huge_valf(); }  // false for NaN and +-inf
// Whether a query may use slab_fast (see its comment).
RT_HD bool fast_slab_ok(const DevScene& sc, V3 o, V3 inv, float tMin, float tMax) {
  return sc.boundsFinite != 0u && is_finite_f(inv.x) && is_finite_f(inv.y) && is_finite_f(inv.z) && is_finite_f(o.x) &&
         is_finite_f(o.y) && is_finite_f(o.z) && tMin == tMin && tMax == tMax;
}

// ---------------------------------------------------------------------------------------------
// Per-ray constants of the triangle test: RayShearParams (src/Primitives.cpp:5-22). The reference
// rebuilds them at every leaf (include/Mesh.hpp:89); they depend on the ray only, so they are hoisted.
struct Shear {
  int kz;  // iz; ix = (kz+1)%3, iy = (kz+2)%3
  float sx, sy, sz;
};
RT_HD Shear make_shear(V3 d) {
  Shear s;
  s.kz = maxi(d);
  const int kx = s.kz == 2 ? 0 : s.kz + 1;
  const int ky = kx == 2 ? 0 : kx + 1;
  const float dx = comp(d, kx), dy = comp(d, ky), dz = comp(d, s.kz);
  s.sx = -dx / dz;
  s.sy = -dy / dz;
  s.sz = 1.f / dz;
  return s;
}
RT_HD V3 permute(V3 v, int kz) {
  // (c[ix], c[iy], c[iz]) for the cyclic permutation selected by kz
  return kz == 0 ? mk(v.y, v.z, v.x) : (kz == 1 ? mk(v.z, v.x, v.y) : v);
}

// TriangleMesh::intersectTriangle (src/Mesh.cpp:6-104) with tFar = +inf, the only value the callers
// use (include/Mesh.hpp:90-92), and ALLOW_DOUBLE_FALLBACK off (CMakeLists.txt:13). Returns t (0 = miss).
RT_HD float tri_test(V3 p0, V3 p1, V3 p2, V3 o, const Shear& sh, float& b0, float& b1, float& b2) {
  V3 p0t = permute(p0 - o, sh.kz);
  V3 p1t = permute(p1 - o, sh.kz);
  V3 p2t = permute(p2 - o, sh.kz);
  p0t.x += sh.sx * p0t.z; p0t.y += sh.sy * p0t.z;
  p1t.x += sh.sx * p1t.z; p1t.y += sh.sy * p1t.z;
  p2t.x += sh.sx * p2t.z; p2t.y += sh.sy * p2t.z;
  const float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  const float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  const float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return 0.f;
  const float det = e0 + e1 + e2;
  if (det == 0) return 0.f;
  p0t.z *= sh.sz; p1t.z *= sh.sz; p2t.z *= sh.sz;
  const float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  const float inf = __builtin_huge_valf();
  if (det < 0.f && (tScaled >= 0.f || tScaled < inf * det)) return 0.f;
  else if (det > 0.f && (tScaled <= 0.f || tScaled > inf * det)) return 0.f;
  const float invDet = 1 / det;
  b0 = e0 * invDet; b1 = e1 * invDet; b2 = e2 * invDet;
  const float t = tScaled * invDet;
  // PBRT-style conservative error bound (Mesh.cpp:85-101); maxc() keeps the reference's chain.
  const float maxZt = maxc(vabs(mk(p0t.z, p1t.z, p2t.z)));
  const float deltaZ = kGamma3 * maxZt;
  const float maxXt = maxc(vabs(mk(p0t.x, p1t.x, p2t.x)));
  const float maxYt = maxc(vabs(mk(p0t.y, p1t.y, p2t.y)));
  const float deltaX = kGamma5 * (maxXt + maxZt);
  const float deltaY = kGamma5 * (maxYt + maxZt);
  const float deltaE = 2 * (kGamma2 * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  const float maxE = maxc(vabs(mk(e0, e1, e2)));
  const float deltaT = 3 * (kGamma3 * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
  if (t <= deltaT) return 0.f;
  return t;
}

// Sphere::intersect (src/Primitives.cpp:24-46). Returns t (0 = miss).
RT_HD float sphere_test(float4 s, V3 o, V3 d, float tMin) {
  const V3 f = mk(s.x, s.y, s.z) - o;
  const float radius2 = s.w * s.w;
  const float rd2 = 1.f / norm2(d);
  const float tca = dot(f, d) * rd2;
  if (tca < 0.f) return 0.f;
  const V3 l = f - d * tca;
  const float l2 = norm2(l);
  if (l2 > radius2) return 0.f;
  const float td = sqrtf(radius2 - l2) * rd2;
  float t0 = tca - td, t1 = tca + td;
  if (t0 > t1) { const float s2 = t0; t0 = t1; t1 = s2; }
  if (t0 < tMin) {
    t0 = t1;
    if (t0 < tMin) return 0.f;
  }
  return t0;
}

// Disc::intersect (src/Primitives.cpp:48-67), including its abs(c.n) plane offset. Returns t (0 = miss).
RT_HD float disc_test(const float* __restrict__ p, V3 o, V3 d) {
  const V3 n = mk(RT_LDG(p + 0), RT_LDG(p + 1), RT_LDG(p + 2));
  const float r = RT_LDG(p + 3);
  const V3 c = mk(RT_LDG(p + 4), RT_LDG(p + 5), RT_LDG(p + 6));
  const float angle = dot(n, d);
  if (angle != 0.f) {
    const float dd = fabsf(dot(c, n));
    const float t = -(dot(n, o) + dd) / angle;
    if (t > kMachineEps) {
      const V3 hp = o + d * t;
      const float d2 = norm2(hp - c);
      if (d2 < r * r) return t;
    }
  }
  return 0.f;
}

// =============================================================================================
// Pair table. The caller's CompactBVH2Node[] (24 B nodes, first child = index + 1, second child =
// secondChildIndex; include/CompactBVH2Node.hpp:52-85) is uploaded unchanged and stays the
// interface. For traversal, scene creation derives one 48-byte record per INNER node holding BOTH of
// its children, numbered in pre-order over inner nodes (root = pair 0):
//
//   word  0..2  L.min.xyz (f32)          word  6..8   R.min.xyz (f32)
//   word  3     L.ref                    word  9      R.ref
//   word  4     L.dx | L.dy << 16 (f16)  word 10      R.dx | R.dy << 16
//   word  5     L.dz | L.geomID << 16    word 11      R.dz | R.geomID << 16
//
// where ref = kRefInner | pair index of the child if it is an inner node (geomID == 0xFFFF), else the leaf
// reference  type << 30 | index  (type 0: global triangle index; 1: sphere index; 2: disc index) — the
// GeomRef/MeshInfo lookup of primLookup (codelets/TraceCodelets.cpp:127-140) folded in at build time.
// One unsigned comparison therefore tells a leaf (ref < kRefInner) from an inner node.
// Bounds are the node's own bits (fp32 min, fp16 extents): the box arithmetic is the reference's.
// A traversal step loads one aligned 48-byte record (3 x LDS.128 / LDG.128) instead of two 24-byte
// nodes from unrelated addresses (6 x 8 B), and a leaf test needs no geometry-table lookup.
// A child is identified by key = pair * 2 + side; leafOrig[key] is its index in the caller's array,
// needed only when two leaves report exactly the same t (the reference keeps the one it meets first
// in pre-order = the lower index).
struct PairWords {
  uint4 q0, q1, q2;
};
constexpr uint32_t kLeafIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kRefInner = 0xC0000000u;  // child reference of an inner node: kRefInner | pair index
constexpr uint32_t kRefNone = 0xFFFFFFFFu;   // "no node": end of a query / no hit yet (never a valid reference)
constexpr uint32_t kRefDone = 0xFFFFFFFEu;   // streaming kernels: the lane has run out of rays
RT_HD bool ref_is_leaf(uint32_t ref) { return ref < kRefInner; }
RT_HD bool ref_is_inner(uint32_t ref) { return ref - kRefInner < 0x3FFFFFFEu; }
RT_HD uint32_t ref_pair(uint32_t ref) { return ref & kLeafIndexMask; }

template <bool kShared>
RT_HD PairWords fetch_pair(const uint4* __restrict__ pairs, uint32_t idx) {
  PairWords w;
  const uint4* p = pairs + 3u * idx;
#if defined(__CUDA_ARCH__)
  if (kShared) { w.q0 = p[0]; w.q1 = p[1]; w.q2 = p[2]; }
  else { w.q0 = __ldg(p); w.q1 = __ldg(p + 1); w.q2 = __ldg(p + 2); }
#else
  w.q0 = p[0]; w.q1 = p[1]; w.q2 = p[2];
#endif
  return w;
}
// ref / geomID of child `side` of pair `idx` (what a pop needs).
template <bool kShared>
RT_HD void fetch_child_ref(const uint4* __restrict__ pairs, uint32_t key, uint32_t& ref, uint32_t& geom) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(pairs + 3u * (key >> 1)) + 6u * (key & 1u);
#if defined(__CUDA_ARCH__)
  if (kShared) { ref = w[3]; geom = w[5] >> 16; }
  else { ref = __ldg(w + 3); geom = __ldg(w + 5) >> 16; }
#else
  ref = w[3]; geom = w[5] >> 16;
#endif
}

// Both box tests of one traversal step. kFast selects slab_fast (see fast_slab_ok).
template <bool kFast>
RT_HD void pair_slabs(const PairWords& w, V3 o, V3 inv, float tMin, float tLimit, bool& h0, bool& h1, float& e0, float& e1) {
  const float lx = bits_f(w.q0.x), ly = bits_f(w.q0.y), lz = bits_f(w.q0.z);
  const float rx = bits_f(w.q1.z), ry = bits_f(w.q1.w), rz = bits_f(w.q2.x);
  if (kFast) {
    h0 = slab_fast(lx, ly, lz, w.q1.x, w.q1.y, o, inv, tMin, tLimit, e0);
    h1 = slab_fast(rx, ry, rz, w.q2.z, w.q2.w, o, inv, tMin, tLimit, e1);
  } else {
    h0 = slab_exact(lx, ly, lz, w.q1.x, w.q1.y, o, inv, tMin, tLimit, e0);
    h1 = slab_exact(rx, ry, rz, w.q2.z, w.q2.w, o, inv, tMin, tLimit, e1);
  }
}

// The root is tested from the caller's own node 0 (once per query).
RT_HD bool root_slab(const DevScene& sc, V3 o, V3 inv, float tMin, float tLimit, bool fast) {
  const uint2 a = RT_LDG(sc.nodes), b = RT_LDG(sc.nodes + 1), c = RT_LDG(sc.nodes + 2);
  float enter;
  const float mx = bits_f(a.x), my = bits_f(a.y), mz = bits_f(b.x);
  return fast ? slab_fast(mx, my, mz, c.x, c.y, o, inv, tMin, tLimit, enter)
              : slab_exact(mx, my, mz, c.x, c.y, o, inv, tMin, tLimit, enter);
}

// One leaf: the t the reference's Intersection would carry (primLookup + virtual Primitive::intersect,
// codelets/TraceCodelets.cpp:127-140): +inf for a missed triangle (Mesh.hpp:90-93: only t > 0 replaces the
// inf default), 0 for a missed sphere/disc (Intersection::Failed()).
RT_HD float leaf_eval(const DevScene& sc, uint32_t ref, V3 o, V3 d, float tMin, const Shear& sh, float& b0, float& b1,
                      float& b2) {
  const uint32_t type = ref >> 30, index = ref & kLeafIndexMask;
  if (type == 0u) {
    const float4* tv = sc.triVerts + 3u * index;
    const float4 a = RT_LDG(tv), b = RT_LDG(tv + 1), c = RT_LDG(tv + 2);
    const float t = tri_test(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), o, sh, b0, b1, b2);
    return t > 0.f ? t : __builtin_huge_valf();
  }
  b0 = b1 = b2 = 0.f;
  if (type == 1u) return sphere_test(RT_LDG(sc.spheres + index), o, d, tMin);
  return disc_test(sc.discs + 7u * index, o, d);
}

// Primitive::normal at the updated hit point (Render.hpp:15-23 -> Mesh.hpp:106-121 /
// Primitives.hpp:49-51 / :73). For meshes the reference computes the normal inside intersect() for
// every accepted candidate; it is a pure function of the winning triangle, so it is computed once.
RT_HD V3 prim_normal(const DevScene& sc, uint32_t geomID, uint32_t tri, float b0, float b1, float b2, V3 hitPoint) {
  const uint2 gw = RT_LDG(reinterpret_cast<const uint2*>(sc.geoms) + geomID);
  GeomEntry g;
  g.type = gw.x; g.first = gw.y;
  if (g.type == 0u) {
    if (sc.triNormals == nullptr) {
      if (sc.triFaceNormals != nullptr) { const float4 fn = RT_LDG(sc.triFaceNormals + tri); return mk(fn.x, fn.y, fn.z); }
      const float4* tv = sc.triVerts + 3u * tri;
      const float4 a = RT_LDG(tv), b = RT_LDG(tv + 1), c = RT_LDG(tv + 2);
      const V3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
      return normalized(cross(p1 - p0, p2 - p0));
    }
    const float4* tn = sc.triNormals + 3u * tri;
    const float4 a = RT_LDG(tn), b = RT_LDG(tn + 1), c = RT_LDG(tn + 2);
    return normalized((mk(a.x, a.y, a.z) * b0 + mk(b.x, b.y, b.z) * b1) + mk(c.x, c.y, c.z) * b2);
  }
  if (g.type == 1u) {
    const float4 s = RT_LDG(sc.spheres + g.first);
    return normalized(hitPoint - mk(s.x, s.y, s.z));
  }
  const float* p = sc.discs + 7u * g.first;
  return mk(RT_LDG(p), RT_LDG(p + 1), RT_LDG(p + 2));
}

// Closest hit of a query, as the streaming kernels keep it.
struct PairHit {
  float t;          // closest t so far (starts at the ray's tMax)
  uint32_t geomID;  // kInvalidGeom when nothing was hit
  uint32_t ref;     // leaf reference of the winner (type << 30 | index)
  uint32_t key;     // pair * 2 + side of the winner (tie-break handle)
  float b0, b1, b2; // barycentrics of the winning triangle
};
// Acceptance of CompactBvh::intersect (CompactBvh.hpp:124): tMin < t < closest; an equal t only replaces
// the current winner if this leaf precedes it in the reference's pre-order walk.
RT_HD bool accept_hit(const DevScene& sc, float t, float tMin, const PairHit& h, uint32_t key) {
  if (!(t > tMin)) return false;
  if (t < h.t) return true;
  return t == h.t && h.geomID != kInvalidGeom && RT_LDG(sc.leafOrig + key) < RT_LDG(sc.leafOrig + h.key);
}
// primID / global triangle index as the reference reports them (triangle index within its mesh, or 0).
RT_HD void hit_ids(const DevScene& sc, const PairHit& h, uint32_t& primID, uint32_t& tri) {
  primID = kInvalidPrim; tri = 0u;
  if (h.geomID == kInvalidGeom) return;
  const uint32_t type = h.ref >> 30, index = h.ref & kLeafIndexMask;
  if (type == 0u) { tri = index; primID = index - RT_LDG(&sc.geoms[h.geomID].first); }
  else primID = 0u;
}

// Closest hit, near-first order over the pair table, one query per caller (lockstep "while-while" on a GPU).
// Same node tests, same primitive tests, same acceptance window as CompactBvh::intersect
// (include/CompactBvh.hpp:80-139); only the visiting order differs (nearer child first, the farther one
// deferred with its entry distance and re-checked against the shrunken closest-t when popped: exactly the
// reference's pop-time slab test because enter <= boxExit was already established).
// `stack` is caller storage of kMaxStack entries {key, enter bits}.
template <bool kShared, bool kCount>
RT_HD void pair_closest_hit(const DevScene& sc, const uint4* __restrict__ pairs, V3 o, V3 d, float tMin, float tMax,
                            PairHit& hit, uint2* stack, uint32_t& nodeVisits, uint32_t& primTests) {
  const V3 inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  const bool fast = fast_slab_ok(sc, o, inv, tMin, tMax);
  hit.t = tMax; hit.geomID = kInvalidGeom; hit.ref = 0u; hit.key = 0u; hit.b0 = hit.b1 = hit.b2 = 0.f;
  if (kCount) nodeVisits++;
  if (!root_slab(sc, o, inv, tMin, hit.t, fast)) return;
  int sp = 0;
  uint32_t ref = sc.rootRef, geom = sc.rootGeom, key = 0u;  // current child: pair index (inner) or leaf reference
  bool done = false;
  auto pop = [&]() {
    done = true;
    while (sp > 0) {
      const uint2 e = stack[--sp];
      if (!(bits_f(e.y) > hit.t)) { key = e.x; done = false; break; }
    }
    if (!done) fetch_child_ref<kShared>(pairs, key, ref, geom);
  };
  while (!done) {
    while (!done && geom == kInvalidGeom) {
      const PairWords w = fetch_pair<kShared>(pairs, ref_pair(ref));
      if (kCount) nodeVisits += 2;
      bool h0, h1;
      float e0, e1;
      if (fast) pair_slabs<true>(w, o, inv, tMin, hit.t, h0, h1, e0, e1);
      else pair_slabs<false>(w, o, inv, tMin, hit.t, h0, h1, e0, e1);
      if (h0 | h1) {
        const bool goL = h0 && (!h1 || !(e1 < e0));  // ties go to the first child, like pre-order
        const uint32_t k0 = ref_pair(ref) * 2u;
        if (h0 && h1) stack[sp++] = make_uint2(goL ? k0 + 1u : k0, f_bits(goL ? e1 : e0));
        key = goL ? k0 : k0 + 1u;
        ref = goL ? w.q0.w : w.q2.y;
        geom = (goL ? w.q1.y : w.q2.w) >> 16;
      } else {
        pop();
      }
    }
    if (done) break;
    if (kCount) primTests++;
    float b0, b1, b2;
    const float t = leaf_eval(sc, ref, o, d, tMin, sh, b0, b1, b2);
    if (accept_hit(sc, t, tMin, hit, key)) {
      hit.t = t; hit.geomID = geom; hit.ref = ref; hit.key = key; hit.b0 = b0; hit.b1 = b1; hit.b2 = b2;
    }
    pop();
  }
}

// Any hit (CompactBvh::occluded, CompactBvh.hpp:33-78): the node window stays [tMin, tMax], a primitive
// occludes iff tMin < t < tMax. The answer does not depend on the visiting order.
template <bool kShared, bool kCount>
RT_HD bool pair_any_hit(const DevScene& sc, const uint4* __restrict__ pairs, V3 o, V3 d, float tMin, float tMax,
                        uint32_t* stack, uint32_t& nodeVisits, uint32_t& primTests) {
  const V3 inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  const bool fast = fast_slab_ok(sc, o, inv, tMin, tMax);
  if (kCount) nodeVisits++;
  if (!root_slab(sc, o, inv, tMin, tMax, fast)) return false;
  int sp = 0;
  uint32_t ref = sc.rootRef, geom = sc.rootGeom;
  while (true) {
    bool needPop = false;
    if (geom == kInvalidGeom) {
      const PairWords w = fetch_pair<kShared>(pairs, ref_pair(ref));
      if (kCount) nodeVisits += 2;
      bool h0, h1;
      float e0, e1;
      if (fast) pair_slabs<true>(w, o, inv, tMin, tMax, h0, h1, e0, e1);
      else pair_slabs<false>(w, o, inv, tMin, tMax, h0, h1, e0, e1);
      if (h0 | h1) {
        if (h0 && h1) stack[sp++] = ref_pair(ref) * 2u + 1u;
        const bool goL = h0;
        ref = goL ? w.q0.w : w.q2.y;
        geom = (goL ? w.q1.y : w.q2.w) >> 16;
      } else {
        needPop = true;
      }
    } else {
      if (kCount) primTests++;
      float b0, b1, b2;
      const float t = leaf_eval(sc, ref, o, d, tMin, sh, b0, b1, b2);
      if (t > tMin && t < tMax) return true;
      needPop = true;
    }
    if (needPop) {
      if (sp == 0) return false;
      fetch_child_ref<kShared>(pairs, stack[--sp], ref, geom);
    }
  }
}


// =============================================================================================
// Streaming traversal: the per-lane state machine of wf_trace_kernel (wavefront.cuh), one step at a time.
//
// The kernel keeps one query per lane and advances it by ONE step per warp iteration: an inner-node step
// (stream_trav: both child boxes of a pair record), a leaf step (stream_leaf: one primitive test) or the start of
// the next query (stream_begin). The steps are written here, RT_HD, so that tests/host_pair_check.cpp runs the very
// same code on the CPU against the oracle. Answers are those of CompactBvh::intersect (include/CompactBvh.hpp:80-139)
// with the bounce loop's window tMin = 0, tMax = inf (trace.cpp:128-130); the visiting order is near-first as in
// pair_closest_hit above.
//
// What keeps the steps short:
//  * the node a lane holds is ONE word, the child reference itself (inner: kRefInner | pair, leaf: type << 30 | index,
//    kRefNone: query finished), so descending is a select and the phase of a lane is a comparison;
//  * deferred children are stacked as {reference, entry distance}; the two top entries live in registers, so a pop is
//    a few moves plus a reload from local memory that is only consumed two pops later;
//  * geomID / primID of the winner are looked up once per query from leafInfo (by the shading kernel), equal-t ties
//    compare the reference node indices stored there;
//  * triangles of a NaN-free query (fast_query_ok) are tested by tri_test_fast on vertex copies whose components are
//    already rotated for the ray's kz: no per-test permutation, no divergent min/max chains.

// Index of a leaf reference's primitive in DevScene::leafInfo: triangles, then spheres, then discs.
RT_HD uint32_t prim_slot(const DevScene& sc, uint32_t ref) {
  const uint32_t type = ref >> 30, index = ref & kLeafIndexMask;
  return index + (type == 0u ? 0u : (type == 1u ? sc.numTris : sc.numTris + sc.numSpheres));
}

RT_HD float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }

// TriangleMesh::intersectTriangle (src/Mesh.cpp:6-104) for the case in which no NaN can occur before the final
// comparisons: triangle vertices, ray origin and the three shear constants all finite and below 2^20 in magnitude.
// Then every translated/sheared coordinate is below 2^42 and every edge function below 2^85 (finite), so the
// reference's Vec3fa::maxi()/operator[] chains over |values| -- which select the MINIMUM component, see rt_math.h --
// equal fminf(fminf(a, b), c) (all inputs non-NaN and non-negative). Everything else is the statement sequence of
// tri_test above. a, b, c are the triangle's vertices with their components already rotated by the ray's kz
// (DevScene::triVerts copy kz + 1, or copy 0 for kz = 2) and op is the ray origin rotated the same way:
// a - op == permute(p0 - o, kz) component by component. Returns t (0 = miss).
RT_HD float tri_test_fast(V3 a, V3 b, V3 c, V3 op, float sx, float sy, float sz, float& b0, float& b1, float& b2) {
  V3 p0t = a - op, p1t = b - op, p2t = c - op;
  p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
  p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
  p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
  const float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  const float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  const float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return 0.f;
  const float det = e0 + e1 + e2;
  if (det == 0) return 0.f;
  p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
  const float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  const float inf = __builtin_huge_valf();
  if (det < 0.f && (tScaled >= 0.f || tScaled < inf * det)) return 0.f;
  else if (det > 0.f && (tScaled <= 0.f || tScaled > inf * det)) return 0.f;
  const float invDet = 1 / det;
  b0 = e0 * invDet; b1 = e1 * invDet; b2 = e2 * invDet;
  const float t = tScaled * invDet;
  const float maxZt = min3f(fabsf(p0t.z), fabsf(p1t.z), fabsf(p2t.z));
  const float deltaZ = kGamma3 * maxZt;
  const float maxXt = min3f(fabsf(p0t.x), fabsf(p1t.x), fabsf(p2t.x));
  const float maxYt = min3f(fabsf(p0t.y), fabsf(p1t.y), fabsf(p2t.y));
  const float deltaX = kGamma5 * (maxXt + maxZt);
  const float deltaY = kGamma5 * (maxYt + maxZt);
  const float deltaE = 2 * (kGamma2 * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  const float maxE = min3f(fabsf(e0), fabsf(e1), fabsf(e2));
  const float deltaT = 3 * (kGamma3 * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
  if (t <= deltaT) return 0.f;
  return t;
}

constexpr float kFastBound = 1048576.f;  // 2^20, see tri_test_fast
RT_HD bool below_bound(float x) { return fabsf(x) < kFastBound; }  // false for NaN and +-inf

struct StreamQuery {
  V3 o, d, inv;       // ray as given (d only where the caller keeps it, see stream_leaf), 1/d
  V3 op;              // origin rotated by kz (tri_test_fast)
  float sx, sy, sz;   // RayShearParams (src/Primitives.cpp:5-22)
  uint32_t permOfs;   // offset, in float4s, of the triangle-vertex copy rotated for this ray's kz
  bool fast;          // no NaN can occur in this query's box and triangle tests: slab_fast / tri_test_fast apply
  float hitT;         // closest t so far
  uint32_t hitRef;    // leaf reference of the winner, kRefNone = nothing hit
  float b0, b1, b2;   // barycentrics of the winning triangle
  uint32_t ref;       // node held: inner (ref_is_inner), leaf (ref_is_leaf) or kRefNone = query finished
  uint32_t topRef;    // the two top entries of the stack of deferred children live in registers
  float topE;
  uint32_t top2Ref;   // ... and the entry below it: a pop's reload from memory is then only consumed two pops later
  float top2E;
  int sp;             // entries below the top that live in `stack`
};

// Per-ray constants of a query, computed where the ray is made (the shading kernel for bounce rays, stream_begin for
// camera rays): 1/d, RayShearParams, and `flags` = rotation of the triangle copies (bits 0-1: 0 for kz = 2, else
// kz + 1) | kStreamFast (slab_fast / tri_test_fast apply: no NaN can occur in this query's box and triangle tests) |
// kStreamRootHit (the ray passes the root's box test, tMin = 0, tMax = inf).
constexpr uint32_t kStreamFast = 4u, kStreamRootHit = 8u;
RT_HD void stream_prepare(const DevScene& sc, V3 o, V3 d, V3& inv, float& sx, float& sy, float& sz, uint32_t& flags) {
  inv = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
  const Shear sh = make_shear(d);
  sx = sh.sx; sy = sh.sy; sz = sh.sz;
  const bool fast = sc.boundsFinite != 0u && sc.trisBounded != 0u && is_finite_f(inv.x) && is_finite_f(inv.y) && is_finite_f(inv.z) &&
                    below_bound(o.x) && below_bound(o.y) && below_bound(o.z) && below_bound(sx) && below_bound(sy) && below_bound(sz);
  flags = (sh.kz == 2 ? 0u : (uint32_t)sh.kz + 1u) | (fast ? kStreamFast : 0u);
  if (root_slab(sc, o, inv, 0.f, __builtin_huge_valf(), fast)) flags |= kStreamRootHit;
}

// Start of CompactBvh::intersect (tMin = 0, tMax = inf) for a ray whose constants are known. q.d is NOT set: only
// sphere and disc tests read the direction, and the kernels fetch it there (see stream_leaf's `lazyD`).
// `stack` needs kMaxStack + 1 entries.
RT_HD void stream_begin_prepared(const DevScene& sc, StreamQuery& q, V3 o, V3 inv, float sx, float sy, float sz, uint32_t flags) {
  q.o = o; q.inv = inv;
  q.sx = sx; q.sy = sy; q.sz = sz;
  const uint32_t perm = flags & 3u;
  q.op = perm == 0u ? o : (perm == 1u ? mk(o.y, o.z, o.x) : mk(o.z, o.x, o.y));  // permute(o, kz)
  q.permOfs = perm * sc.numTris * 3u;
  q.fast = (flags & kStreamFast) != 0u;
  q.hitT = __builtin_huge_valf(); q.hitRef = kRefNone; q.b0 = q.b1 = q.b2 = 0.f;
  // bottom of the stack: popping it ends the query (its entry distance, -inf, is never beyond the closest hit)
  q.topRef = kRefNone; q.topE = -__builtin_huge_valf(); q.sp = 1;
  q.top2Ref = kRefNone; q.top2E = -__builtin_huge_valf();
  q.ref = (flags & kStreamRootHit) ? sc.rootRef : kRefNone;
}
RT_HD void stream_begin(const DevScene& sc, StreamQuery& q, V3 o, V3 d) {
  V3 inv;
  float sx, sy, sz;
  uint32_t flags;
  stream_prepare(sc, o, d, inv, sx, sy, sz, flags);
  q.d = d;
  stream_begin_prepared(sc, q, o, inv, sx, sy, sz, flags);
}

// Next deferred child. Returns true when that child's entry distance lies beyond the closest hit (the reference's
// pop-time slab test would reject it): the caller pops again.
RT_HD bool stream_pop(StreamQuery& q, const uint2* stack) {
  const float e = q.topE;
  q.ref = q.topRef;
  const uint2 below = stack[--q.sp];
  q.topRef = q.top2Ref; q.topE = q.top2E;
  q.top2Ref = below.x; q.top2E = bits_f(below.y);
  return e > q.hitT;
}

// One inner-node step on the pair record `w` of q.ref. kFast must equal q.fast (the kernels run their NaN-free
// queries, all but a handful, through the <true> instantiation only). Returns stream_pop's "pop again".
template <bool kFast>
RT_HD bool stream_trav(StreamQuery& q, const PairWords& w, uint2* stack) {
  bool h0, h1;
  float e0, e1;
  pair_slabs<kFast>(w, q.o, q.inv, 0.f, q.hitT, h0, h1, e0, e1);
  // the second child is entered first only if it is hit and the first child is not, or is hit farther away
  // (ties go to the first child, like pre-order). Near / far are selected BEFORE the branches on purpose: with the
  // selects written inside `if (h0 && h1)` nvcc 12.9 folded them to "first child near" in the <true> instantiation
  // (found on the GPU as lost subtrees; tests/test_gpu_parity.py pins it).
  // = h1 && (!h0 || e1 < e0), written as ONE unordered compare: a missed first child counts as NaN-far away (entry
  // distances themselves are never NaN: they start at tMin and are only replaced by values that compare greater)
  const float e0x = h0 ? e0 : bits_f(0x7fc00000u);
  const bool rightFirst = h1 && !(e0x <= e1);
  const uint32_t nearRef = rightFirst ? w.q2.y : w.q0.w;
  const uint32_t farRef = rightFirst ? w.q0.w : w.q2.y;
  const float farE = rightFirst ? e0 : e1;
  if (h0 && h1) {
    stack[q.sp++] = make_uint2(q.top2Ref, f_bits(q.top2E));
    q.top2Ref = q.topRef; q.top2E = q.topE;
    q.topRef = farRef;
    q.topE = farE;
  }
  if (h0 || h1) {
    q.ref = nearRef;
    return false;
  }
  return stream_pop(q, stack);
}

// One leaf step on q.ref: primitive test (primLookup + Primitive::intersect, codelets/TraceCodelets.cpp:127-140) and
// the acceptance of CompactBvh.hpp:124 (0 < t < closest; an equal t replaces the winner only if this primitive's leaf
// precedes the winner's in the reference's pre-order walk). lazyD: where the ray's direction {x, y, z, -} is kept when
// q.d is not (nullptr = use q.d). Returns stream_pop's "pop again".
template <bool kFast>
RT_HD bool stream_leaf(const DevScene& sc, StreamQuery& q, const uint2* stack, const float4* lazyD = nullptr) {
  const uint32_t type = q.ref >> 30, index = q.ref & kLeafIndexMask;
  float t, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (type == 0u) {
    if (kFast) {
      const float4* tv = sc.triVerts + (q.permOfs + 3u * index);
      const float4 a = RT_LDG(tv), b = RT_LDG(tv + 1), c = RT_LDG(tv + 2);
      t = tri_test_fast(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), q.op, q.sx, q.sy, q.sz, b0, b1, b2);
    } else {
      const float4* tv = sc.triVerts + 3u * index;
      const float4 a = RT_LDG(tv), b = RT_LDG(tv + 1), c = RT_LDG(tv + 2);
      Shear sh;
      sh.kz = q.permOfs == 0u ? 2 : (q.permOfs == sc.numTris * 3u ? 0 : 1);
      sh.sx = q.sx; sh.sy = q.sy; sh.sz = q.sz;
      t = tri_test(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), q.o, sh, b0, b1, b2);
    }
    t = t > 0.f ? t : __builtin_huge_valf();  // Mesh.hpp:90-93: only t > 0 replaces the inf default
  } else {
    V3 d = q.d;
    if (lazyD) { const float4 dd = *lazyD; d = mk(dd.x, dd.y, dd.z); }  // path state: never through the non-coherent path
    if (type == 1u) t = sphere_test(RT_LDG(sc.spheres + index), q.o, d, 0.f);
    else t = disc_test(sc.discs + 7u * index, q.o, d);
  }
  if (t > 0.f) {
    bool accept = t < q.hitT;
    if (!accept && t == q.hitT && q.hitRef != kRefNone)
      accept = RT_LDG(&sc.leafInfo[prim_slot(sc, q.ref)].z) < RT_LDG(&sc.leafInfo[prim_slot(sc, q.hitRef)].z);
    if (accept) { q.hitT = t; q.hitRef = q.ref; q.b0 = b0; q.b1 = b1; q.b2 = b2; }
  }
  return stream_pop(q, stack);
}

// A whole query, one step after the other (host checks; the kernels' handful of queries that are not NaN-free).
template <bool kFast, bool kCount>
RT_HD void stream_run(const DevScene& sc, const uint4* __restrict__ pairs, StreamQuery& q, uint2* stack, const float4* lazyD,
                      uint32_t& nodeVisits, uint32_t& primTests) {
  while (q.ref != kRefNone) {
    bool again;
    if (ref_is_inner(q.ref)) {
      if (kCount) nodeVisits += 2;
      again = stream_trav<kFast>(q, fetch_pair<false>(pairs, ref_pair(q.ref)), stack);
    } else {
      if (kCount) primTests++;
      again = stream_leaf<kFast>(sc, q, stack, lazyD);
    }
    while (again) again = stream_pop(q, stack);
  }
}

// geomID / primID / global triangle index of a finished query's winner, as the reference reports them.
RT_HD void stream_hit_ids(const DevScene& sc, uint32_t hitRef, uint32_t& geomID, uint32_t& primID, uint32_t& tri) {
  geomID = kInvalidGeom; primID = kInvalidPrim; tri = 0u;
  if (hitRef == kRefNone) return;
  const uint4 info = RT_LDG(sc.leafInfo + prim_slot(sc, hitRef));
  geomID = info.x; primID = info.y;
  tri = (hitRef >> 30) == 0u ? (hitRef & kLeafIndexMask) : 0u;
}

}  // namespace rt
