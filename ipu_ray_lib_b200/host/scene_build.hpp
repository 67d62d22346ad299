// Host-side scene construction API (C++). See scene_build.cpp for the reference mapping.
#pragma once
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "rt_types.hpp"

namespace b200rt {

// One mesh before flattening (HostTriangleMesh, include/Mesh.hpp:141).
struct MeshParts {
  std::vector<Triangle> triangles;
  std::vector<Vec3> vertices;
  std::vector<Vec3> normals;  // empty, or one per vertex
};

// `SceneDescription` (include/scene_utils.hpp:30-45) without the sampler.
struct SceneParts {
  std::vector<MeshParts> meshes;
  std::vector<SphereData> spheres;
  std::vector<DiscData> discs;
  std::vector<Material> materials;
  std::vector<std::uint32_t> matIDs;
  float horizontalFov = 0.78539816339744830962f;
};

std::uint16_t roundToHalfNotSmaller(float f);
// Returns max depth (root = 1). nodes.size() == 2n-1 on return.
std::uint32_t buildCompactBvh(const float* primBounds, const std::uint32_t* ids, std::uint32_t n,
                              std::vector<BvhNode>& nodes);
void finaliseScene(const SceneParts& parts, HostScene& out);

SceneParts makeCornellBoxScene(const std::string& meshFile, bool boxOnly);
SceneParts makePrimitiveScene();
SceneParts importScene(const std::string& file, bool loadNormals);  // scene_import.cpp

void initPerspectiveRayStream(TraceResult* rays, int imgW, int imgH, CropWindow win, float fovRadians);
long visualiseHits(const TraceResult* rays, std::size_t n, const b200rt_scene_desc& scene, int mode, float* image,
                   int imgW, int imgH);

void writeExr(const std::string& path, const float* bgr, int w, int h);  // image_io.cpp
void writePfm(const std::string& path, const float* bgr, int w, int h);

}  // namespace b200rt
