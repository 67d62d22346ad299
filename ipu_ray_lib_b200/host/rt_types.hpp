// POD wire types of the trace path. Layouts are fixed by the reference;
// every struct is checked with static_assert so that arrays produced by the reference's
// own scene builder / CompactBvhBuild can be passed through the C ABI unchanged.
//
//   Vec3fa / Ray / HitRecord / TraceResult : include/embree_utils/geometry.hpp:27-260
//   CompactBVH2Node                        : include/CompactBVH2Node.hpp:52-85
//   Triangle                               : include/Primitives.hpp:21-25
//   MeshInfo                               : include/Mesh.hpp:15-20
//   GeomRef / GeomType / CropWindow        : include/Scene.hpp:12-32
//   Material                               : include/Material.hpp:8-33
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace b200rt {

struct Vec3 {
  float x, y, z;
};
static_assert(sizeof(Vec3) == 12 && alignof(Vec3) == 4, "Vec3fa layout");

struct PixelCoord {
  float u;  // row
  float v;  // col
};

struct Ray {
  Vec3 origin;
  float tMin;
  Vec3 direction;
  float tMax;
};
static_assert(sizeof(Ray) == 32, "Ray layout");

struct HitRecord {
  static constexpr std::uint16_t InvalidGeomID = 0xFFFFu;
  static constexpr std::uint32_t InvalidPrimID = 0xFFFFFFFFu;
  static constexpr std::uint16_t ERROR = 1;
  static constexpr std::uint16_t ESCAPED = 2;
  Ray r;
  std::uint32_t primID;
  Vec3 normal;
  Vec3 throughput;
  std::uint16_t geomID;
  std::uint16_t flags;
};
static_assert(sizeof(HitRecord) == 64, "HitRecord layout");
static_assert(offsetof(HitRecord, primID) == 32 && offsetof(HitRecord, normal) == 36 &&
              offsetof(HitRecord, throughput) == 48 && offsetof(HitRecord, geomID) == 60 &&
              offsetof(HitRecord, flags) == 62, "HitRecord offsets");

struct TraceResult {
  Vec3 rgb;
  PixelCoord p;
  HitRecord h;
};
static_assert(sizeof(TraceResult) == 84 && alignof(TraceResult) == 4, "TraceResult layout");
static_assert(offsetof(TraceResult, p) == 12 && offsetof(TraceResult, h) == 20, "TraceResult offsets");

struct alignas(8) BvhNode {
  static constexpr std::uint16_t InvalidGeomID = 0xFFFFu;
  float min_x, min_y, min_z;
  std::uint32_t primOrSecondChild;  // leaf: primID; inner: index of second child (first child = index + 1)
  std::uint16_t dx, dy, dz;         // IEEE binary16 bit patterns of the box extents
  std::uint16_t geomID;             // 0xFFFF => inner node
};
static_assert(sizeof(BvhNode) == 24 && alignof(BvhNode) == 8, "CompactBVH2Node layout");
static_assert(offsetof(BvhNode, primOrSecondChild) == 12 && offsetof(BvhNode, dx) == 16 &&
              offsetof(BvhNode, geomID) == 22, "CompactBVH2Node offsets");

struct Triangle {
  std::uint16_t v0, v1, v2;
};
static_assert(sizeof(Triangle) == 6 && alignof(Triangle) == 2, "Triangle layout");

struct MeshInfo {
  std::uint32_t firstIndex, firstVertex, numTriangles, numVertices;
};
static_assert(sizeof(MeshInfo) == 16, "MeshInfo layout");

enum GeomType : std::uint8_t { GEOM_MESH = 0, GEOM_SPHERE = 1, GEOM_DISC = 2 };
struct GeomRef {
  std::uint16_t index;
  std::uint8_t type;
  std::uint8_t pad;
};
static_assert(sizeof(GeomRef) == 4, "GeomRef layout");

enum MaterialType : std::int32_t { MAT_DIFFUSE = 0, MAT_SPECULAR = 1, MAT_REFRACTIVE = 2 };
struct Material {
  Vec3 albedo;
  float ior;
  Vec3 emission;
  std::int32_t type;
  std::uint8_t emissive;
  std::uint8_t pad[3];
};
static_assert(sizeof(Material) == 36 && offsetof(Material, type) == 28 && offsetof(Material, emissive) == 32,
              "Material layout");

struct SphereData { float x, y, z, radius; };                 // C-ABI sphere record
struct DiscData   { float nx, ny, nz, r, cx, cy, cz; };       // C-ABI disc record

struct CropWindow { std::int32_t w, h, c, r; };

// Owning container of everything a trace needs: `SceneData` (include/Scene.hpp:36-46)
// plus the primitive arrays and camera/render scalars of `SceneDescription`
// (include/scene_utils.hpp:30-45).
struct HostScene {
  std::vector<GeomRef> geometry;
  std::vector<MeshInfo> meshInfo;
  std::vector<Triangle> meshTris;
  std::vector<Vec3> meshVerts;
  std::vector<Vec3> meshNormals;
  std::vector<std::uint32_t> matIDs;
  std::vector<Material> materials;
  std::vector<BvhNode> bvhNodes;
  std::uint32_t bvhMaxDepth = 0;
  std::vector<SphereData> spheres;
  std::vector<DiscData> discs;
  float horizontalFov = 0.78539816339744830962f;
};

}  // namespace b200rt
