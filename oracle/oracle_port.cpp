// oracle_port.cpp — CPU restatement ("port") of the reference's trace path.
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h): the checker for the CUDA path, never the product.
// Plain scalar C++ with no dependency on /root/reference, so it also builds on the GPU box. Every
// function cites the reference lines it restates. It is pinned against the reference's own
// compiled kernel sources (oracle/_ref, built from /root/reference by oracle/Makefile):
// tests/test_oracle_pinning.py requires bit-identical ray streams and known-answer values from both,
// and tests/golden/ holds outputs of the reference build for machines without it.
// Build flags (oracle/Makefile): -ffp-contract=off, no -march, no fast-math — the reference CPU
// path is FMA-free IEEE fp32.
#define ORC_PREFIX orc_
#include "oracle_api.h"

#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// Vec3fa (include/embree_utils/geometry.hpp:27-163)
struct F3 {
  float x, y, z;
};
inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
inline F3 add(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline F3 sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline F3 scale(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
inline F3 mul(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline F3 neg(F3 a) { return f3(-a.x, -a.y, -a.z); }
inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float sqnorm(F3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline F3 cross(F3 a, F3 v) { return f3(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x); }
inline F3 normalized(F3 a) { return scale(a, 1.f / sqrtf(sqnorm(a))); }  // :139
inline F3 absv(F3 a) { return f3(std::abs(a.x), std::abs(a.y), std::abs(a.z)); }
inline float at(F3 a, unsigned i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// :115-121. The comparison chain is restated verbatim; note that it yields the index of the
// smallest signed component although the reference names it maxi().
inline unsigned maxi(F3 a) {
  if (a.x < a.y) { return a.x < a.z ? 0 : 2; }
  return a.y < a.z ? 1 : 2;
}
inline float maxc(F3 a) { return at(a, maxi(a)); }  // :123-125

// include/precision_utils.hpp:19-26
constexpr float machineEpsilon = std::numeric_limits<float>::epsilon() * .5f;
constexpr float gammaf(int i) { return (machineEpsilon * i) / (1 - machineEpsilon * i); }
constexpr float rayEpsilon = machineEpsilon * 1500.f;
const float kInf = std::numeric_limits<float>::infinity();

// ------------------------------------------------------------------------------------------------
// fp16 <-> fp32 (include/precision_utils.hpp:28-47; half = IEEE binary16, RNE)
inline float halfToFloat(uint16_t b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }
inline uint16_t floatToHalf(float f) { _Float16 h = (_Float16)f; uint16_t b; std::memcpy(&b, &h, 2); return b; }
inline uint16_t roundToHalfNotSmaller(float f) {
  uint16_t h = floatToHalf(f);
  if (halfToFloat(h) < f) h = (uint16_t)(h + 1);
  return h;
}

// ------------------------------------------------------------------------------------------------
// sincos (ext/math/sincos.cpp:236-355, ACC5 + ABSERR + MOD360, flg = 0). The table is sin(i deg).
const float* sinTable() {
  static float tbl[92];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 92; ++i) tbl[i] = (float)std::sin((double)i * 3.14159265358979323846264338327950288 / 180.0);
    init = true;
  }
  return tbl;
}
inline void sincosRef(float x, float& s, float& c) {
  const float* tbl = sinTable();
  x = x * float(180.0 / 3.14159265358979323846264338327950288);
  int xsign = 1;
  if (x < 0.f) { xsign = -1; x = -x; }
  x = x - 360.f * std::floor(x / 360.f);
  int ix = x + .5f;
  float z = x - ix;
  int ssign, csign;
  if (ix <= 180) { ssign = 1; csign = 1; } else { ssign = -1; csign = -1; ix -= 180; }
  if (ix > 90) { csign = -csign; ix = 180 - ix; }
  float sx = tbl[ix];
  if (ssign < 0) sx = -sx;
  float cx = tbl[90 - ix];
  if (csign < 0) cx = -cx;
  float sz = 1.74531263774940077459e-2f * z;
  float cz = 1.f - 1.52307909153324666207e-4f * z * z;
  float y = sx * cz + cx * sz;
  if (xsign < 0) y = -y;
  s = y;
  c = cx * cz - sx * sz;
}

// ------------------------------------------------------------------------------------------------
// xoroshiro128** (include/xoshiro.hpp:18-80) + the build-defined streams / Gaussian
inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
inline uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9;
  z = (z ^ (z >> 27)) * 0x94d049bb133111eb;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t s[2];
};
inline void seedRng(Rng& r, uint64_t seed) { r.s[0] = splitmix64(seed); r.s[1] = splitmix64(r.s[0]); }
inline uint64_t nextRng(Rng& r) {
  const uint64_t s0 = r.s[0];
  uint64_t s1 = r.s[1];
  const uint64_t result = rotl(s0 * 5, 7) * 9;
  s1 ^= s0;
  r.s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16);
  r.s[1] = rotl(s1, 37);
  return result;
}
inline float uniform01(Rng& r) {  // :67-80, through double exactly as the reference
  const uint64_t x = nextRng(r);
  union { uint64_t i; double d; } u;
  u.i = UINT64_C(0x3FF) << 52 | x >> 12;
  return (float)(u.d - 1.0);
}
// per-(pixel,sample) stream — build-defined, see ipu_ray_lib_b200/csrc/rt_math.h
inline void seedStream(Rng& r, uint64_t key, uint32_t pixelIndex, uint32_t sample) {
  const uint64_t id = ((uint64_t)pixelIndex << 32) | (uint64_t)sample;
  seedRng(r, splitmix64(id ^ key));
}
inline float detLog(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  int e = (int)((u >> 23) & 0xffu) - 126;
  u = (u & 0x007fffffu) | 0x3f000000u;
  float m;
  std::memcpy(&m, &u, 4);
  if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.f; } else { m = m - 1.f; }
  const float z = m * m;
  float p = 7.0376836292e-2f;
  p = p * m - 1.1514610310e-1f;
  p = p * m + 1.1676998740e-1f;
  p = p * m - 1.2420140846e-1f;
  p = p * m + 1.4249322787e-1f;
  p = p * m - 1.6668057665e-1f;
  p = p * m + 2.0000714765e-1f;
  p = p * m - 2.4999993993e-1f;
  p = p * m + 3.3333331174e-1f;
  float y = p * m * z;
  const float fe = (float)e;
  y = y + -2.12194440e-4f * fe;
  y = y + -0.5f * z;
  float r = m + y;
  r = r + 0.693359375f * fe;
  return r;
}
inline void gaussianPair(uint64_t a, uint64_t b, float& g0, float& g1) {
  const float u1 = (float)((uint32_t)(a >> 40) + 1u) * 5.9604644775390625e-08f;
  const float u2 = (float)((uint32_t)(b >> 40)) * 5.9604644775390625e-08f;
  const float rad = sqrtf(-2.f * detLog(u1));
  float s, c;
  sincosRef(6.28318530717958647692f * u2, s, c);
  g0 = rad * c;
  g1 = rad * s;
}

// ------------------------------------------------------------------------------------------------
// Scene view over the C-ABI arrays
#pragma pack(push, 1)
struct Node {  // include/CompactBVH2Node.hpp:52-85
  float min_x, min_y, min_z;
  uint32_t primOrSecond;
  uint16_t dx, dy, dz;
  uint16_t geomID;
};
struct GeomRefRec { uint16_t index; uint8_t type; uint8_t pad; };
struct MeshInfoRec { uint32_t firstIndex, firstVertex, numTriangles, numVertices; };
struct MaterialRec { float albedo[3]; float ior; float emission[3]; int32_t type; uint8_t emissive; uint8_t pad[3]; };
#pragma pack(pop)
static_assert(sizeof(Node) == 24 && sizeof(MaterialRec) == 36, "layouts");

struct TraceResultRec {  // include/embree_utils/geometry.hpp:212-260
  float rgb[3];
  float row, col;
  float origin[3]; float tMin; float direction[3]; float tMax;
  uint32_t primID; float normal[3]; float throughput[3];
  uint16_t geomID; uint16_t flags;
};
static_assert(sizeof(TraceResultRec) == 84, "TraceResult");
struct RayRec { float origin[3]; float tMin; float direction[3]; float tMax; };

constexpr uint16_t kInvalidGeom = 0xFFFF;
constexpr uint32_t kInvalidPrim = 0xFFFFFFFFu;
constexpr uint16_t kError = 1, kEscaped = 2;

struct Scene {
  const b200rt_scene_desc& d;
  const Node* nodes;
  const GeomRefRec* geom;
  const MeshInfoRec* info;
  const uint16_t* tris;
  const float* verts;
  const float* normals;
  const MaterialRec* materials;
  explicit Scene(const b200rt_scene_desc& desc)
      : d(desc), nodes((const Node*)desc.bvh_nodes), geom((const GeomRefRec*)desc.geometry),
        info((const MeshInfoRec*)desc.mesh_info), tris((const uint16_t*)desc.mesh_tris),
        verts((const float*)desc.mesh_verts), normals((const float*)desc.mesh_normals),
        materials((const MaterialRec*)desc.materials) {}
};

struct Counters {
  uint64_t closest = 0, occl = 0, nodes = 0, prims = 0, samples = 0, escaped = 0;
  void flush(uint64_t* out) const {
    if (!out) return;
    const uint64_t v[6] = {closest, occl, nodes, prims, samples, escaped};
    for (int i = 0; i < 6; ++i) {
#pragma omp atomic
      out[i] += v[i];
    }
  }
};

// ------------------------------------------------------------------------------------------------
// intersectRaySlab (include/CompactBVH2Node.hpp:36-48) and CompactBVH2Node::intersect
// (src/CompactBVH2Node.cpp:5-22)
inline bool raySlab(float invDir, float org, float slabMin, float slabMax, float& t0, float& t1) {
  float tmin = (slabMin - org) * invDir;
  float tmax = (slabMax - org) * invDir;
  if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
  tmax *= 1 + 2 * gammaf(3);
  t0 = tmin > t0 ? tmin : t0;
  t1 = tmax < t1 ? tmax : t1;
  if (t0 > t1) return false;
  return true;
}
inline bool nodeIntersect(const Node& n, F3 o, F3 inv, float& t0, float& t1) {
  float max_x = n.min_x + halfToFloat(n.dx);
  if (raySlab(inv.x, o.x, n.min_x, max_x, t0, t1)) {
    float max_y = n.min_y + halfToFloat(n.dy);
    if (raySlab(inv.y, o.y, n.min_y, max_y, t0, t1)) {
      float max_z = n.min_z + halfToFloat(n.dz);
      if (raySlab(inv.z, o.z, n.min_z, max_z, t0, t1)) return true;
    }
  }
  return false;
}

// Result of Primitive::intersect (include/Intersection.hpp): t, the primID the primitive reports,
// whether a primitive pointer was set (operator bool), plus what normal() needs later.
struct PrimHit {
  float t;
  uint32_t primID;
  bool valid;     // prim != nullptr
  F3 normal;      // meshes: computed inside intersect (Mesh.hpp:93-98)
};

// RayShearParams (src/Primitives.cpp:5-22)
struct ShearParams {
  F3 o, dir;
  unsigned ix, iy, iz;
  float sx, sy, sz;
};
inline ShearParams makeShear(F3 origin, F3 direction) {
  ShearParams p;
  p.o = origin;
  p.iz = maxi(direction);
  p.ix = p.iz + 1; if (p.ix == 3) p.ix = 0;
  p.iy = p.ix + 1; if (p.iy == 3) p.iy = 0;
  p.dir = f3(at(direction, p.ix), at(direction, p.iy), at(direction, p.iz));
  p.sx = -p.dir.x / p.dir.z;
  p.sy = -p.dir.y / p.dir.z;
  p.sz = 1.f / p.dir.z;
  return p;
}
inline F3 permuted(F3 v, const ShearParams& p) { return f3(at(v, p.ix), at(v, p.iy), at(v, p.iz)); }

inline F3 vertexOf(const Scene& sc, const MeshInfoRec& mi, uint16_t v) {
  const float* p = sc.verts + 3 * ((size_t)mi.firstVertex + v);
  return f3(p[0], p[1], p[2]);
}
inline F3 normalOf(const Scene& sc, const MeshInfoRec& mi, uint16_t v) {
  const float* p = sc.normals + 3 * ((size_t)mi.firstVertex + v);
  return f3(p[0], p[1], p[2]);
}

// TriangleMesh::intersect(primID, ray) = intersectTriangle + computeNormal
// (include/Mesh.hpp:88-121, src/Mesh.cpp:6-104; ALLOW_DOUBLE_FALLBACK = 0)
inline PrimHit meshIntersect(const Scene& sc, const MeshInfoRec& mi, uint32_t primID, F3 org, F3 dir) {
  PrimHit res{kInf, kInvalidPrim, false, f3(0, 0, 0)};
  const ShearParams tf = makeShear(org, dir);
  const uint16_t* tri = sc.tris + 3 * ((size_t)mi.firstIndex + primID);
  const F3 p0 = vertexOf(sc, mi, tri[0]), p1 = vertexOf(sc, mi, tri[1]), p2 = vertexOf(sc, mi, tri[2]);
  F3 p0t = permuted(sub(p0, tf.o), tf), p1t = permuted(sub(p1, tf.o), tf), p2t = permuted(sub(p2, tf.o), tf);
  p0t.x += tf.sx * p0t.z; p0t.y += tf.sy * p0t.z;
  p1t.x += tf.sx * p1t.z; p1t.y += tf.sy * p1t.z;
  p2t.x += tf.sx * p2t.z; p2t.y += tf.sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return res;
  float det = e0 + e1 + e2;
  if (det == 0) return res;
  p0t.z *= tf.sz; p1t.z *= tf.sz; p2t.z *= tf.sz;
  float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  const float tFar = kInf;
  if (det < 0.f && (tScaled >= 0.f || tScaled < tFar * det)) return res;
  else if (det > 0.f && (tScaled <= 0.f || tScaled > tFar * det)) return res;
  float invDet = 1 / det;
  float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
  float t = tScaled * invDet;
  float maxZt = maxc(absv(f3(p0t.z, p1t.z, p2t.z)));
  float deltaZ = gammaf(3) * maxZt;
  float maxXt = maxc(absv(f3(p0t.x, p1t.x, p2t.x)));
  float maxYt = maxc(absv(f3(p0t.y, p1t.y, p2t.y)));
  float deltaX = gammaf(5) * (maxXt + maxZt);
  float deltaY = gammaf(5) * (maxYt + maxZt);
  float deltaE = 2 * (gammaf(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  float maxE = maxc(absv(f3(e0, e1, e2)));
  float deltaT = 3 * (gammaf(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * std::abs(invDet);
  if (t <= deltaT) t = 0.f;
  if (t > 0.f && t < kInf) {  // Mesh.hpp:93
    res.t = t;
    res.primID = primID;
    res.valid = true;
    if (sc.d.num_normals == 0) {
      res.normal = normalized(cross(sub(p1, p0), sub(p2, p0)));
    } else {
      const F3 n0 = normalOf(sc, mi, tri[0]), n1 = normalOf(sc, mi, tri[1]), n2 = normalOf(sc, mi, tri[2]);
      res.normal = normalized(add(add(scale(n0, b0), scale(n1, b1)), scale(n2, b2)));
    }
  }
  return res;
}

// Sphere::intersect (src/Primitives.cpp:24-46)
inline PrimHit sphereIntersect(const float* s, F3 org, F3 dir, float tMin) {
  PrimHit fail{0.f, kInvalidPrim, false, f3(0, 0, 0)};
  const F3 centre = f3(s[0], s[1], s[2]);
  const float radius2 = s[3] * s[3];
  F3 f = sub(centre, org);
  float rd2 = 1.f / sqnorm(dir);
  float tca = dot(f, dir) * rd2;
  if (tca < 0.f) return fail;
  F3 l = sub(f, scale(dir, tca));
  float l2 = sqnorm(l);
  if (l2 > radius2) return fail;
  float td = sqrtf(radius2 - l2) * rd2;
  float t0 = tca - td, t1 = tca + td;
  if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
  if (t0 < tMin) {
    t0 = t1;
    if (t0 < tMin) return fail;
  }
  return PrimHit{t0, 0u, true, f3(0, 0, 0)};
}

// Disc::intersect (src/Primitives.cpp:48-67)
inline PrimHit discIntersect(const float* p, F3 org, F3 dir) {
  const F3 n = f3(p[0], p[1], p[2]), c = f3(p[4], p[5], p[6]);
  const float r2 = p[3] * p[3];
  float angle = dot(n, dir);
  if (angle != 0.f) {
    float dd = std::abs(dot(c, n));
    float t = -(dot(n, org) + dd) / angle;
    if (t > machineEpsilon) {
      const F3 hp = add(org, scale(dir, t));
      float d2 = sqnorm(sub(hp, c));
      if (d2 < r2) return PrimHit{t, 0u, true, f3(0, 0, 0)};
    }
  }
  return PrimHit{0.f, kInvalidPrim, false, f3(0, 0, 0)};
}

inline PrimHit primIntersect(const Scene& sc, uint16_t geomID, uint32_t primID, F3 org, F3 dir, float tMin) {
  const GeomRefRec g = sc.geom[geomID];  // primLookup, codelets/TraceCodelets.cpp:127-140
  if (g.type == 0) return meshIntersect(sc, sc.info[g.index], primID, org, dir);
  if (g.type == 1) return sphereIntersect(sc.d.spheres + 4 * (size_t)g.index, org, dir, tMin);
  return discIntersect(sc.d.discs + 7 * (size_t)g.index, org, dir);
}

struct Closest {
  float t;
  uint16_t geomID;
  uint32_t primID;
  bool valid;
  F3 normal;  // mesh normal recorded at intersection time
};

// CompactBvh::intersect (include/CompactBvh.hpp:80-139)
Closest bvhIntersect(const Scene& sc, F3 org, F3 dir, float tMin, float tMax, Counters& cnt) {
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  const F3 inv = f3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
  Closest best{tMax, kInvalidGeom, kInvalidPrim, false, f3(0, 0, 0)};
  while (sp > 0) {
    const uint32_t cur = stack[--sp];
    const Node& node = sc.nodes[cur];
    float t0 = tMin, t1 = best.t;
    cnt.nodes += 1;
    if (nodeIntersect(node, org, inv, t0, t1)) {
      if (node.geomID != kInvalidGeom) {
        cnt.prims += 1;
        const PrimHit h = primIntersect(sc, node.geomID, node.primOrSecond, org, dir, tMin);
        if (h.t > tMin && h.t < best.t) {
          best.t = h.t; best.geomID = node.geomID; best.primID = h.primID; best.valid = h.valid; best.normal = h.normal;
        }
      } else {
        stack[sp++] = node.primOrSecond;
        stack[sp++] = cur + 1;
      }
    }
  }
  return best;
}

// CompactBvh::occluded (include/CompactBvh.hpp:33-78)
bool bvhOccluded(const Scene& sc, F3 org, F3 dir, float tMin, float tMax, Counters& cnt) {
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  const F3 inv = f3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
  while (sp > 0) {
    const uint32_t cur = stack[--sp];
    const Node& node = sc.nodes[cur];
    float t0 = tMin, t1 = tMax;
    cnt.nodes += 1;
    if (nodeIntersect(node, org, inv, t0, t1)) {
      if (node.geomID != kInvalidGeom) {
        cnt.prims += 1;
        const PrimHit h = primIntersect(sc, node.geomID, node.primOrSecond, org, dir, tMin);
        if (h.t > tMin && h.t < tMax) return true;
      } else {
        stack[sp++] = node.primOrSecond;
        stack[sp++] = cur + 1;
      }
    }
  }
  return false;
}

// Primitive::normal (Mesh.hpp:48-51, Primitives.hpp:49-51, :73) at the advanced origin
inline F3 primNormal(const Scene& sc, const Closest& c, F3 hitPoint) {
  const GeomRefRec g = sc.geom[c.geomID];
  if (g.type == 0) return c.normal;
  if (g.type == 1) {
    const float* s = sc.d.spheres + 4 * (size_t)g.index;
    return normalized(sub(hitPoint, f3(s[0], s[1], s[2])));
  }
  const float* p = sc.d.discs + 7 * (size_t)g.index;
  return f3(p[0], p[1], p[2]);
}

// offsetRay (include/Render.hpp:29-33)
inline F3 offsetRay(F3 origin, F3 direction, F3 n) {
  const float m = (1.f + maxc(absv(origin))) * rayEpsilon * std::copysign(1.f, dot(n, direction));
  return add(origin, scale(n, m));
}

// pixelToRayDir (include/Render.hpp:74-85)
inline F3 pixelToRayDir(float x, float y, float w, float h, float tanTheta) {
  const float aspect = w / h;
  x = (x / w) - .5f;
  y = (y / h) - .5f;
  return normalized(f3(2.f * x * aspect * tanTheta, -2.f * y * tanTheta, -1.f));
}

// sampleDiscConcentric / cosineSampleHemisphere (include/geometric_sampling.hpp:8-45)
inline void sampleDiscConcentric(float u1, float u2, float& ox, float& oy) {
  float ux = 2.f * u1 - 1.f, uy = 2.f * u2 - 1.f;
  if (ux == 0.f && uy == 0.f) { ox = ux; oy = uy; return; }
  float r, th;
  if (std::abs(ux) > std::abs(uy)) { r = ux; th = (float)(3.14159265358979323846264338327950288 / 4.0) * (uy / ux); }
  else { r = uy; th = (float)(3.14159265358979323846264338327950288 / 2.0) - (float)(3.14159265358979323846264338327950288 / 4.0) * (ux / uy); }
  float s, c;
  sincosRef(th, s, c);
  ox = r * c;
  oy = r * s;
}
// sampleDiffuse (include/BxDF.hpp:11-30) with orthonormalSystem (geometry.hpp:147-159)
inline F3 sampleDiffuse(F3 n, float u1, float u2) {
  F3 v2;
  const F3 a = absv(n), sq = mul(n, n);
  if (a.x > a.y) { float il = 1.f / sqrtf(sq.x + sq.z); v2 = f3(-n.z * il, 0.f, n.x * il); }
  else { float il = 1.f / sqrtf(sq.y + sq.z); v2 = f3(0.f, n.z * il, -n.y * il); }
  const F3 xB = v2, yB = cross(n, v2);
  float x, y;
  sampleDiscConcentric(u1, u2, x, y);
  float z = sqrtf(std::max(0.f, 1.f - x * x - y * y));
  const F3 wi = f3(x, y, z);
  return f3(dot(f3(xB.x, yB.x, n.x), wi), dot(f3(xB.y, yB.y, n.y), wi), dot(f3(xB.z, yB.z, n.z), wi));
}
inline F3 reflectDir(F3 d, F3 n) {  // BxDF.hpp:33-37
  float c = dot(d, n);
  return normalized(sub(d, scale(n, c * 2.f)));
}
inline float schlick(float cosTheta, float ri) {  // BxDF.hpp:39-46
  float r0 = (1.f - ri) / (1.f + ri);
  r0 = r0 * r0;
  float base = 1.f - cosTheta, base2 = base * base, base5 = base2 * base * base2;
  return r0 + (1.f - r0) * base5;
}
inline F3 refractDir(F3 dir, F3 n, float ndotr, float ri) {  // BxDF.hpp:48-55
  const float cosTheta = -ndotr;
  F3 rPerp = scale(add(dir, scale(n, cosTheta)), ri);
  F3 rPar = scale(n, -sqrtf(std::abs(1.f - sqnorm(rPerp))));
  return add(rPerp, rPar);
}
inline F3 dielectric(F3 dir, F3 n, float ri, float u1, bool& refracted) {  // BxDF.hpp:57-75
  if (dot(n, dir) > 0.f) n = neg(n); else ri = 1.f / ri;
  const float ndotr = dot(n, dir);
  const float cost1 = -ndotr;
  const float cost2 = 1.f - ri * ri * (1.f - cost1 * cost1);
  if (cost2 > 0.f && u1 > schlick(cost1, ri)) { refracted = true; return refractDir(dir, n, ndotr, ri); }
  refracted = false;
  return reflectDir(dir, n);
}
inline bool evaluateRoulette(float u1, F3& thr) {  // geometric_sampling.hpp:56-63
  const float p = maxc(thr);
  if (p == 0.f || u1 > p) return true;
  thr = scale(thr, 1.f / p);
  return false;
}

inline float fovTan(float fov) { float s, c; sincosRef(fov / 2.f, s, c); return s / c; }
inline F3 ld3(const float* p) { return f3(p[0], p[1], p[2]); }
inline void st3(float* p, F3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// ------------------------------------------------------------------------------------------------
// NIF forward on the CPU: encode (src/neural_networks/NifModel.cpp:186-219, host twin :417-433),
// Dense layers with the auto-detected concat (:296-327), decode (:222-246 / :493-510).
// Numerics mirror the B200 kernel's contract: fp16 features/weights/layer outputs, fp32 accumulation.
// Weights widened to fp32 once per call (exact), so the per-sample loop is a plain fp32 GEMV.
struct NifHost {
  const b200rt_nif_desc& d;
  std::vector<std::vector<float>> w, b;
  explicit NifHost(const b200rt_nif_desc& nif) : d(nif), w(nif.num_layers), b(nif.num_layers) {
    for (uint32_t l = 0; l < nif.num_layers; ++l) {
      const b200rt_nif_layer& L = nif.layers[l];
      w[l].resize((size_t)L.in_features * L.out_features);
      for (size_t i = 0; i < w[l].size(); ++i) w[l][i] = halfToFloat(L.kernel_f16[i]);
      if (L.bias_f16) {
        b[l].resize(L.out_features);
        for (uint32_t i = 0; i < L.out_features; ++i) b[l][i] = halfToFloat(L.bias_f16[i]);
      }
    }
  }
};

// halfChunk > 0 emulates the IPU's matmul option partialsType = half (src/IpuScene.cpp:256-262): products of
// `halfChunk` consecutive k are summed exactly-ish (fp32, one AMP pass) and the running partial is ROUNDED TO FP16 after
// every such group and again after the bias. poplibs does not publish its accumulation order, so this is a model of
// that option, not a pinned restatement; it bounds how far fp16 partials can sit from the fp32-accumulate contract.
void nifForward(const NifHost& host, float u, float v, float out[3], int halfChunk = 0) {
  const b200rt_nif_desc& nif = host.d;
  const int E = (int)nif.embedding_dimension, F = 4 * E;
  std::vector<float> feat((size_t)F);
  const float un = (u - 1.f) * 2.f, vn = (v - 1.f) * 2.f;
  float coeff = 1.f;
  for (int j = 0; j < E; ++j, coeff *= 2.f) {
    const float au = halfToFloat(floatToHalf(un * coeff));
    const float av = halfToFloat(floatToHalf(vn * coeff));
    feat[(size_t)j] = halfToFloat(floatToHalf(std::sin(au)));
    feat[(size_t)(j + E)] = halfToFloat(floatToHalf(std::sin(av)));
    feat[(size_t)(j + 2 * E)] = halfToFloat(floatToHalf(std::cos(au)));
    feat[(size_t)(j + 3 * E)] = halfToFloat(floatToHalf(std::cos(av)));
  }
  std::vector<float> x = feat, y;
  for (uint32_t l = 0; l < nif.num_layers; ++l) {
    const b200rt_nif_layer& L = nif.layers[l];
    if (x.size() != L.in_features) x.insert(x.end(), feat.begin(), feat.end());  // concat(x, input), :303-309
    y.assign(L.out_features, 0.f);
    const float* W = host.w[l].data();
    // y[n] = sum_k x[k] * W[k][n], k ascending for every n (loop order chosen so the compiler vectorises over n)
    if (halfChunk <= 0) {
      for (uint32_t k = 0; k < L.in_features; ++k) {
        const float xk = x[k];
        const float* row = W + (size_t)k * L.out_features;
        for (uint32_t n = 0; n < L.out_features; ++n) y[n] += xk * row[n];
      }
    } else {
      std::vector<float> part(L.out_features);
      for (uint32_t k0 = 0; k0 < L.in_features; k0 += (uint32_t)halfChunk) {
        std::fill(part.begin(), part.end(), 0.f);
        for (uint32_t k = k0; k < std::min<uint32_t>(L.in_features, k0 + (uint32_t)halfChunk); ++k) {
          const float xk = x[k];
          const float* row = W + (size_t)k * L.out_features;
          for (uint32_t n = 0; n < L.out_features; ++n) part[n] += xk * row[n];
        }
        for (uint32_t n = 0; n < L.out_features; ++n) y[n] = halfToFloat(floatToHalf(y[n] + part[n]));
      }
    }
    for (uint32_t n = 0; n < L.out_features; ++n) {
      float acc = y[n];
      if (L.bias_f16) acc += host.b[l][n];
      if (L.relu) acc = acc > 0.f ? acc : 0.f;
      y[n] = halfToFloat(floatToHalf(acc));
    }
    x.swap(y);
  }
  for (int c = 0; c < 3; ++c) {
    float val = x[(size_t)c] * nif.max + nif.mean[c];
    if (nif.log_tone_map) val = std::exp(val);
    out[c] = val;
  }
}

// PreProcessEscapedRays (codelets/TraceCodelets.cpp:337-348)
inline void dirToUv(F3 d, float rotation, float& u, float& v) {
  const float TwoPi = (float)(2.0 * 3.14159265358979323846264338327950288);
  const float InvPi = (float)(1.0 / 3.14159265358979323846264338327950288);
  const float Inv2Pi = (float)(1.0 / (2.0 * 3.14159265358979323846264338327950288));
  float theta = acosf(d.y);
  float phi = atan2f(d.z, d.x) + rotation;
  if (phi < 0.f) phi += TwoPi;
  else if (phi > TwoPi) phi -= TwoPi;
  u = theta * InvPi;
  v = phi * Inv2Pi;
}

int threadsOr(int t) { return t > 0 ? t : omp_get_max_threads(); }

}  // namespace

// Serialiser<16> restated (include/serialisation/Serialiser.hpp:24-64, serialisation.hpp:21-52): every object is
// padded to its own alignment relative to a 16-byte aligned base; an array is a u32 count followed by its raw elements.
namespace {
struct BlobWriter {
  std::vector<uint8_t> bytes;
  void pad(size_t align) {
    const size_t rem = (16 + bytes.size()) % align;
    if (rem) bytes.resize(bytes.size() + (align - rem));
  }
  template <class T> void scalar(const T& v) {
    pad(alignof(T));
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    bytes.insert(bytes.end(), p, p + sizeof(T));
  }
  void array(const void* data, uint32_t count, size_t elemSize, size_t elemAlign) {
    scalar(count);
    pad(elemAlign);
    const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
    bytes.insert(bytes.end(), p, p + (size_t)count * elemSize);
  }
};
}  // namespace

extern "C" {

size_t orc_serialise_scene(const b200rt_scene_desc* d, uint8_t* out, size_t cap) {
  BlobWriter w;
  w.array(d->geometry, d->num_geometry, 4, 2);       // GeomRef {u16, u8, u8}
  w.array(d->mesh_info, d->num_meshes, 16, 4);       // MeshInfo
  w.array(d->mesh_tris, d->num_tris, 6, 2);          // Triangle {u16 x 3}
  w.array(d->mesh_verts, d->num_verts, 12, 4);       // Vec3fa
  w.array(d->mesh_normals, d->num_normals, 12, 4);
  w.array(d->mat_ids, d->num_mat_ids, 4, 4);
  w.array(d->materials, d->num_materials, 36, 4);    // Material
  w.array(d->bvh_nodes, d->num_bvh_nodes, 24, 4);    // CompactBVH2Node {f32 x 3, u32, f16 x 3, u16}
  w.scalar(d->max_leaf_depth);
  w.scalar(d->image_width);
  w.scalar(d->image_height);
  w.scalar(d->fov_radians);
  w.scalar(d->anti_alias_scale);
  w.scalar(d->max_path_length);
  w.scalar(d->roulette_start_depth);
  w.scalar(d->samples_per_pixel);
  if (out && cap >= w.bytes.size()) std::memcpy(out, w.bytes.data(), w.bytes.size());
  return w.bytes.size();
}

const char* orc_kind(void) { return "port"; }

// traceShadowRay (include/Render.hpp:37-72) over the stream (trace.cpp:246-255)
int orc_shadow_trace(const b200rt_scene_desc* d, void* raysV, size_t n, const float lp[3], float ambient, int threads,
                     uint64_t* counters) {
  const Scene sc(*d);
  auto* rays = (TraceResultRec*)raysV;
  const F3 light = f3(lp[0], lp[1], lp[2]);
#pragma omp parallel num_threads(threadsOr(threads))
  {
    Counters cnt;
#pragma omp for schedule(dynamic, 512)
    for (long long i = 0; i < (long long)n; ++i) {
      TraceResultRec& r = rays[i];
      F3 org = ld3(r.origin);
      const F3 dir = ld3(r.direction);
      cnt.closest += 1;
      const Closest c = bvhIntersect(sc, org, dir, r.tMin, r.tMax, cnt);
      if (c.valid) {
        // updateHit (Render.hpp:15-23)
        r.geomID = c.geomID;
        r.primID = c.primID;
        r.tMax = c.t;
        org = add(org, scale(dir, c.t));
        st3(r.origin, org);
        const F3 nrm = primNormal(sc, c, org);
        st3(r.normal, nrm);
        const MaterialRec& m = sc.materials[sc.d.mat_ids[c.geomID]];
        const F3 lightOffset = sub(light, org);
        const F3 sdir = normalized(lightOffset);
        const F3 sorg = offsetRay(org, sdir, nrm);
        const float sMax = std::sqrt(sqnorm(lightOffset));
        const F3 albedo = ld3(m.albedo);
        F3 color = scale(albedo, ambient);
        cnt.occl += 1;
        if (!bvhOccluded(sc, sorg, sdir, 0.f, sMax, cnt)) color = add(color, scale(albedo, dot(sdir, nrm)));
        st3(r.rgb, color);
      } else {
        r.flags |= kEscaped;
      }
    }
    cnt.flush(counters);
  }
  return 0;
}

// pathTrace (trace.cpp:115-188) with per-sample camera rays (codelets/TraceCodelets.cpp:142-164) and,
// when a NIF is given, the escaped-ray environment lookup (codelets/TraceCodelets.cpp:321-382).
int orc_path_trace(const b200rt_scene_desc* d, void* raysV, size_t n, uint32_t firstSample, uint32_t numSamples,
                   const b200rt_nif_desc* nif, float hdriRotationDegrees, int threads, uint64_t* counters) {
  const Scene sc(*d);
  auto* rays = (TraceResultRec*)raysV;
  const float tanTheta = fovTan(d->fov_radians);
  const uint64_t key = splitmix64(d->rng_seed);
  const uint32_t imgW = (uint32_t)d->image_width;
  const float rotation = (hdriRotationDegrees / 360.f) * (float)(2.0 * M_PI);  // src/IpuScene.cpp:641
  std::unique_ptr<NifHost> nifHost;
  if (nif) nifHost.reset(new NifHost(*nif));
#pragma omp parallel num_threads(threadsOr(threads))
  {
    Counters cnt;
#pragma omp for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)n; ++i) {
      TraceResultRec& r = rays[i];
      const uint32_t row = (uint32_t)r.row, col = (uint32_t)r.col;
      F3 rgb = ld3(r.rgb);
      for (uint32_t s = firstSample; s < firstSample + numSamples; ++s) {
        Rng rng;
        seedStream(rng, key, row * imgW + col, s);
        const uint64_t a = nextRng(rng), b = nextRng(rng);
        float g0, g1;
        gaussianPair(a, b, g0, g1);
        const float pu = r.row + d->anti_alias_scale * g0;
        const float pv = r.col + d->anti_alias_scale * g1;
        // HitRecord(origin, dir) (geometry.hpp:237-243)
        F3 org = f3(0.f, 0.f, 0.f);
        F3 dir = pixelToRayDir(pv, pu, d->image_width, d->image_height, tanTheta);
        F3 nrm = f3(0.f, 0.f, 1.f);
        r.primID = kInvalidPrim; r.geomID = kInvalidGeom; r.flags = 0;
        r.tMin = 0.f; r.tMax = kInf;
        cnt.samples += 1;

        F3 thr = f3(1.f, 1.f, 1.f), color = f3(0.f, 0.f, 0.f);
        bool escaped = false;
        for (uint32_t bounce = 0; bounce < d->max_path_length; ++bounce) {
          org = offsetRay(org, dir, nrm);
          r.tMin = 0.f;
          r.tMax = kInf;
          cnt.closest += 1;
          const Closest c = bvhIntersect(sc, org, dir, 0.f, kInf, cnt);
          if (c.valid) {
            r.geomID = c.geomID; r.primID = c.primID; r.tMax = c.t;
            org = add(org, scale(dir, c.t));
            nrm = primNormal(sc, c, org);
            const MaterialRec& m = sc.materials[sc.d.mat_ids[c.geomID]];
            if (m.emissive) color = add(color, mul(thr, ld3(m.emission)));
            if (m.type == 0) {
              const float u1 = uniform01(rng);
              const float u2 = uniform01(rng);
              dir = sampleDiffuse(nrm, u1, u2);
              thr = mul(thr, ld3(m.albedo));
            } else if (m.type == 1) {
              dir = reflectDir(dir, nrm);
              thr = mul(thr, ld3(m.albedo));
            } else if (m.type == 2) {
              const float u1 = uniform01(rng);
              bool refracted;
              dir = dielectric(dir, nrm, m.ior, u1, refracted);
              if (refracted) thr = mul(thr, ld3(m.albedo));
            } else {
              rgb = scale(rgb, std::numeric_limits<float>::quiet_NaN());
              r.flags |= kError;
            }
          } else {
            r.flags |= kEscaped;
            escaped = true;
            break;
          }
          if (bounce > d->roulette_start_depth) {
            const float u1 = uniform01(rng);
            if (evaluateRoulette(u1, thr)) break;
          }
        }
        rgb = add(rgb, color);  // result.rgb += color
        if (escaped) {
          cnt.escaped += 1;
          if (nif) {  // PreProcess -> NIF -> PostProcessEscapedRays
            float u, v, bgr[3];
            dirToUv(dir, rotation, u, v);
            nifForward(*nifHost, u, v, bgr);
            rgb = add(rgb, mul(thr, f3(bgr[2], bgr[1], bgr[0])));
          }
        }
        st3(r.origin, org); st3(r.direction, dir); st3(r.normal, nrm); st3(r.throughput, thr);
      }
      st3(r.rgb, rgb);
    }
    cnt.flush(counters);
  }
  return 0;
}

int orc_intersect(const b200rt_scene_desc* d, const void* raysV, size_t n, b200rt_hit* out, int threads, uint64_t* counters) {
  const Scene sc(*d);
  auto* rays = (const RayRec*)raysV;
#pragma omp parallel num_threads(threadsOr(threads))
  {
    Counters cnt;
#pragma omp for schedule(dynamic, 512)
    for (long long i = 0; i < (long long)n; ++i) {
      const F3 org = ld3(rays[i].origin), dir = ld3(rays[i].direction);
      cnt.closest += 1;
      const Closest c = bvhIntersect(sc, org, dir, rays[i].tMin, rays[i].tMax, cnt);
      b200rt_hit h;
      if (c.valid) {
        const F3 nrm = primNormal(sc, c, add(org, scale(dir, c.t)));
        h.t = c.t; h.geom_id = c.geomID; h.prim_id = c.primID;
        h.normal[0] = nrm.x; h.normal[1] = nrm.y; h.normal[2] = nrm.z;
      } else {
        h.t = rays[i].tMax; h.geom_id = 0xFFFFu; h.prim_id = 0xFFFFFFFFu;
        h.normal[0] = h.normal[1] = h.normal[2] = 0.f;
      }
      out[i] = h;
    }
    cnt.flush(counters);
  }
  return 0;
}

int orc_occluded(const b200rt_scene_desc* d, const void* raysV, size_t n, uint8_t* out, int threads) {
  const Scene sc(*d);
  auto* rays = (const RayRec*)raysV;
#pragma omp parallel for schedule(dynamic, 512) num_threads(threadsOr(threads))
  for (long long i = 0; i < (long long)n; ++i) {
    Counters cnt;
    out[i] = bvhOccluded(sc, ld3(rays[i].origin), ld3(rays[i].direction), rays[i].tMin, rays[i].tMax, cnt) ? 1 : 0;
  }
  return 0;
}

void orc_sincos(const float* x, size_t n, float* s, float* c) {
  for (size_t i = 0; i < n; ++i) sincosRef(x[i], s[i], c[i]);
}
void orc_uniform_stream(uint64_t seed, size_t n, float* out) {
  Rng r; seedRng(r, seed);
  for (size_t i = 0; i < n; ++i) out[i] = uniform01(r);
}
void orc_raw_stream(uint64_t seed, size_t n, uint64_t* out) {
  Rng r; seedRng(r, seed);
  for (size_t i = 0; i < n; ++i) out[i] = nextRng(r);
}
void orc_sample_diffuse(const float* nrm, const float* u12, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) st3(out + 3 * i, sampleDiffuse(ld3(nrm + 3 * i), u12[2 * i], u12[2 * i + 1]));
}
void orc_dielectric(const float* dirs, const float* nrm, const float* iorU1, size_t n, float* out, uint8_t* refr) {
  for (size_t i = 0; i < n; ++i) {
    bool refracted;
    st3(out + 3 * i, dielectric(ld3(dirs + 3 * i), ld3(nrm + 3 * i), iorU1[2 * i], iorU1[2 * i + 1], refracted));
    refr[i] = refracted ? 1 : 0;
  }
}
void orc_reflect(const float* dirs, const float* nrm, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) st3(out + 3 * i, reflectDir(ld3(dirs + 3 * i), ld3(nrm + 3 * i)));
}
void orc_offset_ray(const float* org, const float* dirs, const float* nrm, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) st3(out + 3 * i, offsetRay(ld3(org + 3 * i), ld3(dirs + 3 * i), ld3(nrm + 3 * i)));
}
void orc_pixel_to_ray_dir(const float* xy, size_t n, float w, float h, float tanTheta, float* out) {
  for (size_t i = 0; i < n; ++i) st3(out + 3 * i, pixelToRayDir(xy[2 * i], xy[2 * i + 1], w, h, tanTheta));
}
void orc_round_to_half_not_smaller(const float* x, size_t n, uint16_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = roundToHalfNotSmaller(x[i]);
}
void orc_camera_sample(uint64_t rngSeed, uint32_t w, uint32_t h, float fov, float aa, const uint32_t* rcs, size_t n,
                       float* out) {
  const float tanTheta = fovTan(fov);
  const uint64_t key = splitmix64(rngSeed);
  for (size_t i = 0; i < n; ++i) {
    Rng rng;
    seedStream(rng, key, rcs[3 * i] * w + rcs[3 * i + 1], rcs[3 * i + 2]);
    const uint64_t a = nextRng(rng), b = nextRng(rng);
    float g0, g1;
    gaussianPair(a, b, g0, g1);
    const float pu = (float)rcs[3 * i] + aa * g0, pv = (float)rcs[3 * i + 1] + aa * g1;
    st3(out + 3 * i, pixelToRayDir(pv, pu, (float)w, (float)h, tanTheta));
  }
}
int orc_nif_eval_partials(const b200rt_nif_desc* nif, const float* uv, size_t n, float* out, int threads, int halfChunk) {
  const NifHost host(*nif);
#pragma omp parallel for schedule(static) num_threads(threadsOr(threads))
  for (long long i = 0; i < (long long)n; ++i) nifForward(host, uv[2 * i], uv[2 * i + 1], out + 3 * i, halfChunk);
  return 0;
}

int orc_nif_eval(const b200rt_nif_desc* nif, const float* uv, size_t n, float* out, int threads) {
  const NifHost host(*nif);
#pragma omp parallel for schedule(static) num_threads(threadsOr(threads))
  for (long long i = 0; i < (long long)n; ++i) nifForward(host, uv[2 * i], uv[2 * i + 1], out + 3 * i);
  return 0;
}
void orc_dir_to_uv(const float* dirs, size_t n, float rotation, float* out) {
  for (size_t i = 0; i < n; ++i) dirToUv(ld3(dirs + 3 * i), rotation, out[2 * i], out[2 * i + 1]);
}

}  // extern "C"
