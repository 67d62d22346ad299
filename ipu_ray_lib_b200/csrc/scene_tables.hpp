// Host-side derived tables of a scene (built once at b200rt_scene_create, then uploaded): the
// geomID -> (type, first) table that collapses GeomRef + MeshInfo (include/Scene.hpp:27-32,
// include/Mesh.hpp:15-20), triangle vertices / vertex normals gathered per triangle, and the pair
// table of the BVH (pair_build.hpp). Pure host code, shared with the CPU check of the traversal core
// (tests/host_pair_check.cpp).
#pragma once
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "pair_build.hpp"

namespace rt {

struct SceneTables {
  std::vector<GeomEntry> geoms;
  std::vector<float> triVerts;    // [3][num_tris][3][4]: as given, then rotated for kz = 0 (y,z,x) and kz = 1 (z,x,y)
  bool trisBounded = true;        // every vertex finite and below 2^20 in magnitude (tri_test_fast)
  std::vector<float> triFaceNormals;  // [num_tris][4]: Primitive::normal of a mesh triangle without vertex normals (Mesh.hpp:106-121)
  std::vector<float> triNormals;  // same shape, empty when the scene has no normals
  PairTable pairs;
};

// Returns an empty string on success, else the reason the description is invalid.
inline std::string build_scene_tables(const b200rt_scene_desc& d, SceneTables& out) {
  struct GeomRefH { uint16_t index; uint8_t type; uint8_t pad; };
  struct MeshInfoH { uint32_t firstIndex, firstVertex, numTriangles, numVertices; };
  const auto* geom = (const GeomRefH*)d.geometry;
  const auto* info = (const MeshInfoH*)d.mesh_info;
  const auto* tris = (const uint16_t*)d.mesh_tris;
  const auto* verts = (const float*)d.mesh_verts;
  const auto* normals = (const float*)d.mesh_normals;
  out.geoms.assign(d.num_geometry, GeomEntry{0u, 0u});
  std::vector<uint32_t> primCount(d.num_geometry, 0xFFFFFFFFu);
  for (uint32_t g = 0; g < d.num_geometry; ++g) {
    out.geoms[g].type = geom[g].type;
    if (geom[g].type == 0) {
      if (geom[g].index >= d.num_meshes) return "GeomRef mesh index out of range";
      out.geoms[g].first = info[geom[g].index].firstIndex;
      primCount[g] = info[geom[g].index].numTriangles;
    } else if (geom[g].type == 1) {
      if (geom[g].index >= d.num_spheres) return "GeomRef sphere index out of range";
      out.geoms[g].first = geom[g].index;
    } else if (geom[g].type == 2) {
      if (geom[g].index >= d.num_discs) return "GeomRef disc index out of range";
      out.geoms[g].first = geom[g].index;
    } else {
      return "unknown GeomType";
    }
  }
  out.triVerts.assign((size_t)d.num_tris * 12 * 3, 0.f);
  out.trisBounded = true;
  out.triNormals.clear();
  if (d.num_normals) out.triNormals.assign((size_t)d.num_tris * 12, 0.f);
  for (uint32_t m = 0; m < d.num_meshes; ++m) {
    const MeshInfoH& mi = info[m];
    if ((uint64_t)mi.firstIndex + mi.numTriangles > d.num_tris || (uint64_t)mi.firstVertex + mi.numVertices > d.num_verts)
      return "MeshInfo range exceeds the unified arrays";
    for (uint32_t t = 0; t < mi.numTriangles; ++t) {
      const size_t gt = (size_t)mi.firstIndex + t;
      for (int k = 0; k < 3; ++k) {
        const uint32_t vi = tris[3 * gt + k];
        if (vi >= mi.numVertices) return "triangle index exceeds its mesh's vertex window";
        const size_t gv = (size_t)mi.firstVertex + vi;
        const float* v = verts + 3 * gv;
        std::memcpy(&out.triVerts[12 * gt + 4 * k], v, 12);
        float* r0 = &out.triVerts[12 * ((size_t)d.num_tris + gt) + 4 * k];      // permute(v, 0) = (y, z, x)
        float* r1 = &out.triVerts[12 * (2 * (size_t)d.num_tris + gt) + 4 * k];  // permute(v, 1) = (z, x, y)
        r0[0] = v[1]; r0[1] = v[2]; r0[2] = v[0];
        r1[0] = v[2]; r1[1] = v[0]; r1[2] = v[1];
        for (int c = 0; c < 3; ++c)
          if (!(std::fabs(v[c]) < kFastBound)) out.trisBounded = false;
        if (d.num_normals) std::memcpy(&out.triNormals[12 * gt + 4 * k], normals + 3 * gv, 12);
      }
    }
  }
  out.triFaceNormals.assign((size_t)d.num_tris * 4, 0.f);
  for (uint32_t t = 0; t < d.num_tris; ++t) {
    const float* v = &out.triVerts[12 * (size_t)t];
    const V3 p0 = mk(v[0], v[1], v[2]), p1 = mk(v[4], v[5], v[6]), p2 = mk(v[8], v[9], v[10]);
    const V3 n = normalized(cross(p1 - p0, p2 - p0));  // same statement as prim_normal
    float* o = &out.triFaceNormals[4 * (size_t)t];
    o[0] = n.x; o[1] = n.y; o[2] = n.z;
  }
  return build_pair_table(d.bvh_nodes, d.num_bvh_nodes, out.geoms.data(), d.num_geometry, primCount.data(), d.num_tris,
                          d.num_spheres, d.num_discs, out.pairs);
}

}  // namespace rt
