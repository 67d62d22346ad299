// TEST INFRASTRUCTURE: runs the traversal core of the CUDA kernels (ipu_ray_lib_b200/csrc/rt_prims.h — the pair table,
// the NaN-free fast slab test, near-first closest hit, any hit) on the CPU, compiled by g++ from the very same header,
// so that tests can compare it with the oracle in this GPU-less container. It is NOT a fallback: nothing in the product
// links or loads this library.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../ipu_ray_lib_b200/csrc/scene_tables.hpp"

namespace {
struct HostView {
  rt::SceneTables tables;
  rt::DevScene dev{};
};
std::string g_err;

bool make_view(const b200rt_scene_desc* d, HostView& v) {
  g_err = rt::build_scene_tables(*d, v.tables);
  if (!g_err.empty()) return false;
  rt::DevScene& s = v.dev;
  s.nodes = (const uint2*)d->bvh_nodes;
  s.geoms = v.tables.geoms.data();
  s.triVerts = (const float4*)v.tables.triVerts.data();
  s.triNormals = v.tables.triNormals.empty() ? nullptr : (const float4*)v.tables.triNormals.data();
  s.triFaceNormals = (const float4*)v.tables.triFaceNormals.data();
  s.spheres = (const float4*)d->spheres;
  s.discs = d->discs;
  s.matIDs = d->mat_ids;
  s.materials = (const float*)d->materials;
  s.numNodes = d->num_bvh_nodes;
  s.numMaterials = d->num_materials;
  s.pairs = (const uint4*)v.tables.pairs.words.data();
  s.leafOrig = v.tables.pairs.leafOrig.data();
  s.numPairs = v.tables.pairs.numPairs;
  s.rootRef = v.tables.pairs.rootRef;
  s.rootGeom = v.tables.pairs.rootGeom;
  s.boundsFinite = v.tables.pairs.boundsFinite ? 1u : 0u;
  s.leafInfo = (const uint4*)v.tables.pairs.leafInfo.data();
  s.numTris = d->num_tris;
  s.numSpheres = d->num_spheres;
  s.trisBounded = v.tables.trisBounded ? 1u : 0u;
  return true;
}
}  // namespace

extern "C" {

const char* hostpair_last_error() { return g_err.c_str(); }

// counters: [0] node visits, [1] primitive tests, [2] queries that took the NaN-free fast slab path
int hostpair_intersect(const b200rt_scene_desc* d, const void* raysIn, size_t n, b200rt_hit* out, uint64_t* counters) {
  HostView v;
  if (!make_view(d, v)) return -1;
  const float* rays = (const float*)raysIn;
  uint64_t nv = 0, np = 0, nf = 0;
  std::vector<uint2> stack(rt::kMaxStack);
  for (size_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    const rt::V3 o = rt::mk(r[0], r[1], r[2]), dir = rt::mk(r[4], r[5], r[6]);
    rt::PairHit h;
    uint32_t a = 0, b = 0;
    rt::pair_closest_hit<false, true>(v.dev, v.dev.pairs, o, dir, r[3], r[7], h, stack.data(), a, b);
    nv += a; np += b;
    nf += rt::fast_slab_ok(v.dev, o, rt::mk(1.f / dir.x, 1.f / dir.y, 1.f / dir.z), r[3], r[7]) ? 1 : 0;
    b200rt_hit q;
    uint32_t tri;
    q.t = h.t; q.geom_id = h.geomID;
    rt::hit_ids(v.dev, h, q.prim_id, tri);
    q.normal[0] = q.normal[1] = q.normal[2] = 0.f;
    if (h.geomID != rt::kInvalidGeom) {
      const rt::V3 nn = rt::prim_normal(v.dev, h.geomID, tri, h.b0, h.b1, h.b2, o + dir * h.t);
      q.normal[0] = nn.x; q.normal[1] = nn.y; q.normal[2] = nn.z;
    }
    out[i] = q;
  }
  if (counters) { counters[0] = nv; counters[1] = np; counters[2] = nf; }
  return 0;
}

// The per-lane state machine of wf_trace_kernel (stream_begin / stream_trav / stream_leaf / stream_pop), one query at a
// time, tMin = 0 and tMax = inf. counters: [0] node visits, [1] primitive tests, [2] queries on the NaN-free fast path.
int hostpair_stream_intersect(const b200rt_scene_desc* d, const void* raysIn, size_t n, b200rt_hit* out, uint64_t* counters) {
  HostView v;
  if (!make_view(d, v)) return -1;
  const float* rays = (const float*)raysIn;
  uint64_t nv = 0, np = 0, nf = 0;
  std::vector<uint2> stack(rt::kMaxStack + 1, make_uint2(0u, 0u));
  for (size_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    const rt::V3 o = rt::mk(r[0], r[1], r[2]), dir = rt::mk(r[4], r[5], r[6]);
    // as the kernels do it: constants prepared where the ray is made, the direction kept aside (lazyD)
    rt::StreamQuery q;
    rt::V3 inv;
    float sx, sy, sz;
    uint32_t flags;
    rt::stream_prepare(v.dev, o, dir, inv, sx, sy, sz, flags);
    rt::stream_begin_prepared(v.dev, q, o, inv, sx, sy, sz, flags);
    q.d = rt::mk(0.f, 0.f, 0.f);
    const float4 lazy = make_float4(dir.x, dir.y, dir.z, 0.f);
    nf += q.fast ? 1 : 0;
    uint32_t a = 1, b = 0;
    if (q.fast) rt::stream_run<true, true>(v.dev, v.dev.pairs, q, stack.data(), &lazy, a, b);
    else rt::stream_run<false, true>(v.dev, v.dev.pairs, q, stack.data(), &lazy, a, b);
    nv += a; np += b;
    b200rt_hit h;
    uint32_t tri;
    h.t = q.hitT;
    rt::stream_hit_ids(v.dev, q.hitRef, h.geom_id, h.prim_id, tri);
    h.normal[0] = h.normal[1] = h.normal[2] = 0.f;
    if (h.geom_id != rt::kInvalidGeom) {
      const rt::V3 nn = rt::prim_normal(v.dev, h.geom_id, tri, q.b0, q.b1, q.b2, o + dir * q.hitT);
      h.normal[0] = nn.x; h.normal[1] = nn.y; h.normal[2] = nn.z;
    }
    out[i] = h;
  }
  if (counters) { counters[0] = nv; counters[1] = np; counters[2] = nf; }
  return 0;
}

int hostpair_occluded(const b200rt_scene_desc* d, const void* raysIn, size_t n, uint8_t* out) {
  HostView v;
  if (!make_view(d, v)) return -1;
  const float* rays = (const float*)raysIn;
  std::vector<uint32_t> stack(rt::kMaxStack);
  for (size_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    uint32_t a = 0, b = 0;
    out[i] = rt::pair_any_hit<false, false>(v.dev, v.dev.pairs, rt::mk(r[0], r[1], r[2]), rt::mk(r[4], r[5], r[6]), r[3], r[7],
                                            stack.data(), a, b) ? 1 : 0;
  }
  return 0;
}

// Structural validation only (what b200rt_scene_create runs before anything is uploaded).
int hostpair_validate(const b200rt_scene_desc* d, uint32_t* numPairs, uint32_t* maxDepth, uint32_t* boundsFinite) {
  HostView v;
  if (!make_view(d, v)) return -1;
  if (numPairs) *numPairs = v.tables.pairs.numPairs;
  if (maxDepth) *maxDepth = v.tables.pairs.maxDepth;
  if (boundsFinite) *boundsFinite = v.dev.boundsFinite;
  return 0;
}

}  // extern "C"
