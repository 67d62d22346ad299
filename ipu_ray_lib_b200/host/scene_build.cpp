// Host-side scene construction for the trace path: built-in scenes, glTF-binary mesh import,
// flattening into the unified arrays of `SceneData`, and a binned-SAH BVH2 builder that emits the
// reference's 24-byte CompactBVH2Node array in pre-order.
//
// Reference behaviour reproduced (not its code): src/scene_utils.cpp:27-56 (quads), :102-149
// (importMesh placement), :319-597 (built-in scenes), src/app_utils.cpp:145-188 (one build
// primitive per triangle / sphere / disc), :291-364 (array flattening, geomID order = meshes,
// spheres, discs), src/CompactBvhBuild.cpp:5-56 (pre-order flatten, first child = index + 1,
// roundToHalfNotSmaller extents, depth of root = 1). Embree's rtcBuildBVH (un-vendored dependency)
// is replaced by the builder below; topology therefore differs from a real reference build, which
// only matters for equal-t ties.
#include "scene_build.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "../csrc/rt_math.h"
#include "mini_json.hpp"

namespace b200rt {

// ------------------------------------------------------------------------------------------------
// fp16 helpers (precision_utils.hpp:28-47)
static std::uint16_t halfBitsRne(float f) {
  _Float16 h = (_Float16)f;
  std::uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}
static float halfBitsToFloat(std::uint16_t b) {
  _Float16 h;
  std::memcpy(&h, &b, 2);
  return (float)h;
}
std::uint16_t roundToHalfNotSmaller(float f) {
  std::uint16_t b = halfBitsRne(f);
  if (halfBitsToFloat(b) < f) b = (std::uint16_t)(b + 1);
  return b;
}

// ------------------------------------------------------------------------------------------------
// BVH2 builder
namespace {

struct Box {
  float lo[3], hi[3];
  Box() {
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
  }
  void grow(const float* mn, const float* mx) {
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], mn[a]); hi[a] = std::max(hi[a], mx[a]); }
  }
  void grow(const Box& o) { grow(o.lo, o.hi); }
  float halfArea() const {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
  }
};

struct BuildPrim {
  Box box;
  float centroid[3];
  std::uint32_t geomID, primID;
};

constexpr int kBins = 32;

struct Builder {
  std::vector<BuildPrim>& prims;
  std::vector<BvhNode>& nodes;
  std::uint32_t maxDepth = 0;
  bool sweep = true;   // B200RT_BVH_SWEEP=0 falls back to the 32-bin estimate everywhere
  static constexpr std::size_t kSweepLimit = 1u << 16;

  static BvhNode pack(const Box& b) {
    BvhNode n;
    n.min_x = b.lo[0]; n.min_y = b.lo[1]; n.min_z = b.lo[2];
    const float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
    for (float e : d)
      if (e > 65504.f) throw std::runtime_error("Cannot compress BVH bounds into fp16 (half)");
    n.dx = roundToHalfNotSmaller(d[0]);
    n.dy = roundToHalfNotSmaller(d[1]);
    n.dz = roundToHalfNotSmaller(d[2]);
    n.geomID = BvhNode::InvalidGeomID;
    n.primOrSecondChild = 0;
    return n;
  }

  // Exact SAH: every split position along every axis of the centroid-sorted list (O(n log n) per node). Used for
  // ranges up to kSweepLimit primitives; larger ranges use the binned estimate below.
  std::size_t splitSweep(std::size_t b, std::size_t e) {
    const std::size_t n = e - b;
    float bestCost = INFINITY;
    int bestAxis = -1;
    std::size_t bestLeft = 0;
    std::vector<std::uint32_t> order(n), bestOrder;
    std::vector<float> rightArea(n + 1);
    for (int axis = 0; axis < 3; ++axis) {
      for (std::size_t i = 0; i < n; ++i) order[i] = (std::uint32_t)i;
      std::stable_sort(order.begin(), order.end(), [&](std::uint32_t x, std::uint32_t y) {
        return prims[b + x].centroid[axis] < prims[b + y].centroid[axis];
      });
      Box acc;
      for (std::size_t i = n; i-- > 1;) {
        acc.grow(prims[b + order[i]].box);
        rightArea[i] = acc.halfArea();
      }
      acc = Box();
      for (std::size_t i = 1; i < n; ++i) {  // left = first i primitives
        acc.grow(prims[b + order[i - 1]].box);
        const float cost = acc.halfArea() * (float)i + rightArea[i] * (float)(n - i);
        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestLeft = i; }
      }
      if (bestAxis == axis) bestOrder = order;
    }
    if (bestAxis < 0) return b + n / 2;
    std::vector<BuildPrim> tmp(n);
    for (std::size_t i = 0; i < n; ++i) tmp[i] = prims[b + bestOrder[i]];
    std::copy(tmp.begin(), tmp.end(), prims.begin() + (long)b);
    return b + bestLeft;
  }

  // Returns the split position in [b+1, e-1]; partitions prims[b,e) in place.
  std::size_t split(std::size_t b, std::size_t e) {
    if (sweep && e - b <= kSweepLimit) return splitSweep(b, e);
    Box cb;
    for (std::size_t i = b; i < e; ++i) cb.grow(prims[i].centroid, prims[i].centroid);
    float bestCost = INFINITY;
    int bestAxis = -1, bestBin = -1;
    for (int axis = 0; axis < 3; ++axis) {
      const float ext = cb.hi[axis] - cb.lo[axis];
      if (!(ext > 0.f)) continue;
      Box binBox[kBins];
      std::uint32_t binCount[kBins] = {0};
      const float scale = (float)kBins / ext;
      auto binOf = [&](const BuildPrim& p) {
        int k = (int)((p.centroid[axis] - cb.lo[axis]) * scale);
        return k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
      };
      for (std::size_t i = b; i < e; ++i) {
        const int k = binOf(prims[i]);
        binBox[k].grow(prims[i].box);
        binCount[k] += 1;
      }
      float rightArea[kBins];
      std::uint32_t rightCount[kBins];
      Box acc;
      std::uint32_t cnt = 0;
      for (int k = kBins - 1; k > 0; --k) {
        acc.grow(binBox[k]);
        cnt += binCount[k];
        rightArea[k] = acc.halfArea();
        rightCount[k] = cnt;
      }
      acc = Box();
      cnt = 0;
      for (int k = 0; k < kBins - 1; ++k) {
        acc.grow(binBox[k]);
        cnt += binCount[k];
        if (cnt == 0 || rightCount[k + 1] == 0) continue;
        const float cost = acc.halfArea() * (float)cnt + rightArea[k + 1] * (float)rightCount[k + 1];
        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = k; }
      }
    }
    if (bestAxis >= 0) {
      const int axis = bestAxis;
      const float ext = cb.hi[axis] - cb.lo[axis];
      const float scale = (float)kBins / ext;
      auto mid = std::stable_partition(prims.begin() + b, prims.begin() + e, [&](const BuildPrim& p) {
        int k = (int)((p.centroid[axis] - cb.lo[axis]) * scale);
        k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
        return k <= bestBin;
      });
      const std::size_t m = (std::size_t)(mid - prims.begin());
      if (m > b && m < e) return m;
    }
    return b + (e - b) / 2;  // coincident centroids: split the list in half
  }

  Box build(std::size_t b, std::size_t e, std::uint32_t depth, std::uint32_t& myIndex) {
    myIndex = (std::uint32_t)nodes.size();
    nodes.emplace_back();
    if (depth > maxDepth) maxDepth = depth;
    if (e - b == 1) {
      BvhNode n = pack(prims[b].box);
      n.geomID = (std::uint16_t)prims[b].geomID;
      n.primOrSecondChild = prims[b].primID;
      nodes[myIndex] = n;
      return prims[b].box;
    }
    const std::size_t m = split(b, e);
    std::uint32_t first, second;
    Box box = build(b, m, depth + 1, first);
    box.grow(build(m, e, depth + 1, second));
    BvhNode n = pack(box);
    n.primOrSecondChild = second;
    nodes[myIndex] = n;
    return box;
  }
};

}  // namespace

std::uint32_t buildCompactBvh(const float* primBounds, const std::uint32_t* ids, std::uint32_t n,
                              std::vector<BvhNode>& nodes) {
  if (n == 0) throw std::runtime_error("BVH build needs at least one primitive");
  std::vector<BuildPrim> prims(n);
  for (std::uint32_t i = 0; i < n; ++i) {
    BuildPrim& p = prims[i];
    p.box.grow(primBounds + 6 * i, primBounds + 6 * i + 3);
    for (int a = 0; a < 3; ++a) p.centroid[a] = (p.box.hi[a] + p.box.lo[a]) * .5f;
    p.geomID = ids[2 * i];
    p.primID = ids[2 * i + 1];
  }
  nodes.clear();
  nodes.reserve(2 * (std::size_t)n - 1);
  Builder builder{prims, nodes};
  if (const char* e = std::getenv("B200RT_BVH_SWEEP")) builder.sweep = e[0] != '0';
  std::uint32_t root;
  builder.build(0, n, 1, root);
  return builder.maxDepth;
}

// ------------------------------------------------------------------------------------------------
// Flattening of a scene description into the unified arrays + BVH
void finaliseScene(const SceneParts& parts, HostScene& out) {
  out = HostScene();
  out.horizontalFov = parts.horizontalFov;
  out.materials = parts.materials;
  out.matIDs = parts.matIDs;
  out.spheres = parts.spheres;
  out.discs = parts.discs;

  std::vector<float> bounds;
  std::vector<std::uint32_t> ids;
  auto addPrim = [&](const float* lo, const float* hi, std::uint32_t geomID, std::uint32_t primID) {
    bounds.insert(bounds.end(), lo, lo + 3);
    bounds.insert(bounds.end(), hi, hi + 3);
    ids.push_back(geomID);
    ids.push_back(primID);
  };

  for (const auto& m : parts.meshes) {
    if (m.vertices.size() > 65536) throw std::runtime_error("mesh has more than 65536 vertices (u16 indices)");
    out.meshInfo.push_back(MeshInfo{(std::uint32_t)out.meshTris.size(), (std::uint32_t)out.meshVerts.size(),
                                    (std::uint32_t)m.triangles.size(), (std::uint32_t)m.vertices.size()});
    out.meshTris.insert(out.meshTris.end(), m.triangles.begin(), m.triangles.end());
    out.meshVerts.insert(out.meshVerts.end(), m.vertices.begin(), m.vertices.end());
    out.meshNormals.insert(out.meshNormals.end(), m.normals.begin(), m.normals.end());
  }
  std::uint32_t geomID = 0;
  for (std::size_t i = 0; i < parts.meshes.size(); ++i, ++geomID) {
    out.geometry.push_back(GeomRef{(std::uint16_t)i, GEOM_MESH, 0});
    const auto& m = parts.meshes[i];
    for (std::uint32_t t = 0; t < m.triangles.size(); ++t) {
      const Triangle& tri = m.triangles[t];
      Box b;
      for (std::uint16_t vi : {tri.v0, tri.v1, tri.v2}) {
        const Vec3& v = m.vertices[vi];
        const float p[3] = {v.x, v.y, v.z};
        b.grow(p, p);
      }
      addPrim(b.lo, b.hi, geomID, t);
    }
  }
  for (std::size_t i = 0; i < parts.spheres.size(); ++i, ++geomID) {
    out.geometry.push_back(GeomRef{(std::uint16_t)i, GEOM_SPHERE, 0});
    const SphereData& s = parts.spheres[i];
    const float lo[3] = {s.x - s.radius, s.y - s.radius, s.z - s.radius};
    const float hi[3] = {s.x + s.radius, s.y + s.radius, s.z + s.radius};
    addPrim(lo, hi, geomID, 0);
  }
  for (std::size_t i = 0; i < parts.discs.size(); ++i, ++geomID) {
    out.geometry.push_back(GeomRef{(std::uint16_t)i, GEOM_DISC, 0});
    const DiscData& d = parts.discs[i];
    const float lo[3] = {d.cx - d.r, d.cy - d.r, d.cz - d.r};
    const float hi[3] = {d.cx + d.r, d.cy + d.r, d.cz + d.r};
    addPrim(lo, hi, geomID, 0);
  }
  if (out.matIDs.size() < geomID) throw std::logic_error("All primitives must be assigned a material.");
  out.bvhMaxDepth = buildCompactBvh(bounds.data(), ids.data(), (std::uint32_t)ids.size() / 2, out.bvhNodes);
}

// ------------------------------------------------------------------------------------------------
// Built-in scenes
namespace {

void addQuad(MeshParts& mesh, const float (&q)[4][3]) {
  const std::uint32_t base = (std::uint32_t)mesh.vertices.size();
  for (auto& p : q) mesh.vertices.push_back(Vec3{p[0], p[1], p[2]});
  mesh.triangles.push_back(Triangle{(std::uint16_t)(base + 0), (std::uint16_t)(base + 1), (std::uint16_t)(base + 2)});
  mesh.triangles.push_back(Triangle{(std::uint16_t)(base + 2), (std::uint16_t)(base + 3), (std::uint16_t)(base + 0)});
}

// Cornell box measurements (the published Cornell data, as used by src/scene_utils.cpp:319-456).
const float kLight[4][3] = {{343, 548.7998f, 227}, {343, 548.7998f, 332}, {213, 548.7998f, 332}, {213, 548.7998f, 227}};
const float kFloor[4][3] = {{552.8f, 0, 0}, {0, 0, 0}, {0, 0, 559.2f}, {549.6f, 0, 559.2f}};
const float kCeiling[4][3] = {{556, 548.8f, 0}, {556, 548.8f, 559.2f}, {0, 548.8f, 559.2f}, {0, 548.8f, 0}};
const float kBack[4][3] = {{549.6f, 0, 559.2f}, {0, 0, 559.2f}, {0, 548.8f, 559.2f}, {556, 548.8f, 559.2f}};
const float kRightWall[4][3] = {{0, 0, 559.2f}, {0, 0, 0}, {0, 548.8f, 0}, {0, 548.8f, 559.2f}};
const float kLeftWall[4][3] = {{552.8f, 0, 0}, {549.6f, 0, 559.2f}, {556, 548.8f, 559.2f}, {556, 548.8f, 0}};
const float kShort[5][4][3] = {
    {{130, 165, 65}, {82, 165, 225}, {240, 165, 272}, {290, 165, 114}},
    {{290, 0, 114}, {290, 165, 114}, {240, 165, 272}, {240, 0, 272}},
    {{130, 0, 65}, {130, 165, 65}, {290, 165, 114}, {290, 0, 114}},
    {{82, 0, 225}, {82, 165, 225}, {130, 165, 65}, {130, 0, 65}},
    {{240, 0, 272}, {240, 165, 272}, {82, 165, 225}, {82, 0, 225}}};
const float kTall[5][4][3] = {
    {{423, 330, 247}, {265, 330, 296}, {314, 330, 456}, {472, 330, 406}},
    {{423, 0, 247}, {423, 330, 247}, {472, 330, 406}, {472, 0, 406}},
    {{472, 0, 406}, {472, 330, 406}, {314, 330, 456}, {314, 0, 456}},
    {{314, 0, 456}, {314, 330, 456}, {265, 330, 296}, {265, 0, 296}},
    {{265, 0, 296}, {265, 330, 296}, {423, 330, 247}, {423, 0, 247}}};

Material makeMaterial(Vec3 albedo, Vec3 emission, MaterialType type) {
  Material m;
  std::memset(&m, 0, sizeof(m));
  m.albedo = albedo;
  m.ior = 1.52f;
  m.emission = emission;
  m.type = type;
  m.emissive = (emission.x != 0.f || emission.y != 0.f || emission.z != 0.f) ? 1 : 0;
  return m;
}

}  // namespace

void importMeshForBox(const std::string& file, std::vector<MeshParts>& meshes);  // gltf_import.cpp

SceneParts makeCornellBoxScene(const std::string& meshFile, bool boxOnly) {
  SceneParts scene;
  MeshParts light, white, red, green, shortBlock, tallBlock;
  addQuad(light, kLight);
  addQuad(white, kFloor);
  addQuad(white, kCeiling);
  addQuad(white, kBack);
  addQuad(green, kRightWall);
  addQuad(red, kLeftWall);
  for (auto& q : kShort) addQuad(shortBlock, q);
  for (auto& q : kTall) addQuad(tallBlock, q);
  scene.meshes = {light, white, red, green, shortBlock, tallBlock};

  if (!boxOnly) {
    scene.spheres.push_back(SphereData{450.f, 37.f, 90.f, 37.f});
    scene.spheres.push_back(SphereData{350.f, 37.f, 90.f, 37.f});
    scene.discs.push_back(DiscData{1.f, 0.f, 0.f, 60.f, 0.0002f, 300.f, 250.f});
    importMeshForBox(meshFile, scene.meshes);
  }

  // Camera to the origin and handedness flip (src/scene_utils.cpp:474-508).
  const Vec3 cam{278.f, 273.f, -800.f};
  for (auto& m : scene.meshes)
    for (auto& v : m.vertices) {
      v.x -= cam.x; v.y -= cam.y; v.z -= cam.z;
      v.x = -v.x;
      v.z = -v.z;
    }
  for (auto& s : scene.spheres) {
    s.x -= cam.x; s.y -= cam.y; s.z -= cam.z;
    s.x = -s.x;
    s.z = -s.z;
  }
  for (auto& d : scene.discs) {
    d.cx -= cam.x; d.cy -= cam.y; d.cz -= cam.z;
    d.cx = -d.cx;
    d.cz = -d.cz;
    d.nx = -d.nx;
    d.nz = -d.nz;
  }

  const Vec3 black{0.f, 0.f, 0.f}, red_{.66f, 0.f, 0.f}, green_{0.f, .48f, 0.f}, blue{0.4f, 0.4f, .85f};
  const Vec3 blueLight{0.4f * 2.f, 0.7f * 2.f, .92f * 2.f};
  const Vec3 white_{.75f, .75f, .75f}, grey{.4f, .4f, .4f}, lightR{0.78f, 0.78f, 0.78f};
  const Vec3 lightE{(100.f * 15.6f + 100.f * 18.4f) / 255.f, (100.f * 8.f + 74.5f * 15.6f) / 255.f,
                    (57.3f * 8.f) / 255.f};
  scene.materials = {makeMaterial(white_, black, MAT_DIFFUSE),  makeMaterial(red_, black, MAT_DIFFUSE),
                     makeMaterial(green_, black, MAT_DIFFUSE),  makeMaterial(blue, black, MAT_REFRACTIVE),
                     makeMaterial(lightR, lightE, MAT_DIFFUSE), makeMaterial(grey, black, MAT_SPECULAR),
                     makeMaterial(blue, blueLight, MAT_DIFFUSE), makeMaterial(blue, black, MAT_DIFFUSE)};
  // light, white parts, left wall, right wall, short box, tall box | loaded meshes | sphere, sphere, disc
  scene.matIDs = {4, 0, 1, 2, 0, 5, 0, 0, 3, 7, 6};
  scene.horizontalFov = 0.78539816339744830962f;  // Piby4
  return scene;
}

SceneParts makePrimitiveScene() {
  SceneParts scene;
  scene.horizontalFov = 1.57079632679489661923f;  // Piby2
  scene.spheres = {SphereData{-1.8575f, -0.98714f, -3.6f, 0.6f},  SphereData{0.74795f, -0.55f, -4.3816f, 1.05f},
                   SphereData{1.9929f, -1.08666f, (float)-3.23, 0.5f}, SphereData{(float)-0.19931, -1.183f, -2.75f, 0.4f},
                   SphereData{(float)-0.19931, -1.183f, -2.75f, 0.4010f}};
  scene.discs = {DiscData{0.f, 1.f, 0.f, 3.5f, 0.f, -1.6f, -5.22f}};
  const Vec3 zero{0.f, 0.f, 0.f}, one{1.f, 1.f, 1.f};
  scene.materials = {makeMaterial(Vec3{1.f, .89f, .55f}, zero, MAT_DIFFUSE),
                     makeMaterial(one, zero, MAT_SPECULAR),
                     makeMaterial(Vec3{0.75f, 0.75f, 0.75f}, zero, MAT_REFRACTIVE),
                     makeMaterial(Vec3{.8f, .06f, .391f}, zero, MAT_DIFFUSE),
                     makeMaterial(one, zero, MAT_REFRACTIVE),
                     makeMaterial(Vec3{.98f, .76f, .66f}, zero, MAT_DIFFUSE)};
  scene.matIDs = {0, 1, 2, 3, 4, 5};
  return scene;
}

// ------------------------------------------------------------------------------------------------
// Ray stream helpers (src/app_utils.cpp:19-59)
void initPerspectiveRayStream(TraceResult* rays, int imgW, int imgH, CropWindow win, float fovRadians) {
  float s, c;
  rt::sincos_tbl(fovRadians / 2.f, s, c);
  const float tanTheta = s / c;
  std::size_t i = 0;
  for (std::uint32_t r = (std::uint32_t)win.r; r < (std::uint32_t)(win.r + win.h); ++r) {
    for (std::uint32_t col = (std::uint32_t)win.c; col < (std::uint32_t)(win.c + win.w); ++col) {
      const rt::V3 d = rt::pixel_to_ray_dir((float)col, (float)r, (float)imgW, (float)imgH, tanTheta);
      TraceResult& t = rays[i++];
      std::memset(&t, 0, sizeof(t));
      t.p = PixelCoord{(float)r, (float)col};
      t.h.r.origin = Vec3{0.f, 0.f, 0.f};
      t.h.r.tMin = 0.f;
      t.h.r.direction = Vec3{d.x, d.y, d.z};
      t.h.r.tMax = INFINITY;
      t.h.primID = HitRecord::InvalidPrimID;
      t.h.normal = Vec3{0.f, 0.f, 1.f};
      t.h.geomID = HitRecord::InvalidGeomID;
      t.h.flags = 0;
    }
  }
}

long visualiseHits(const TraceResult* rays, std::size_t n, const b200rt_scene_desc& scene, int mode, float* image,
                   int imgW, int imgH) {
  const auto* matIDs = scene.mat_ids;
  const auto* materials = (const Material*)scene.materials;
  long hits = 0;
  for (std::size_t i = 0; i < n; ++i) {
    const TraceResult& t = rays[i];
    const HitRecord& h = t.h;
    const bool valid = h.geomID != HitRecord::InvalidGeomID;
    float b = 0.f, g = 0.f, r = 0.f;  // stored B,G,R like cv::Vec3f
    switch (mode) {
      case 0: b = t.rgb.z; g = t.rgb.y; r = t.rgb.x; break;
      case 1: if (valid) { b = (float)(h.geomID + 1); g = (float)(h.primID + 1); r = (float)(matIDs[h.geomID] + 1); } break;
      case 2: if (valid) { b = h.normal.z; g = h.normal.y; r = h.normal.x; } break;
      case 3: b = g = r = h.r.tMax; break;
      case 4: if (valid) { const Vec3& c = materials[matIDs[h.geomID]].albedo; b = c.z; g = c.y; r = c.x; } break;
      case 5: if (valid) { b = h.r.origin.z; g = h.r.origin.y; r = h.r.origin.x; } break;
      default: throw std::runtime_error("bad visualise mode");
    }
    const long row = (long)t.p.u, col = (long)t.p.v;
    if (row >= 0 && row < imgH && col >= 0 && col < imgW) {
      float* px = image + 3 * ((std::size_t)row * imgW + col);
      px[0] = b; px[1] = g; px[2] = r;
    }
    if (valid) hits += 1;
  }
  return hits;
}

}  // namespace b200rt
