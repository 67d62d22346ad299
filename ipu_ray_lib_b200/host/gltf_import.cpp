// Minimal glTF-binary (.glb) mesh reader: JSON chunk + one BIN chunk, POSITION/NORMAL + u16/u32
// indices, node TRS/matrix transforms baked into the vertices. It stands in for the assimp pipeline
// the reference calls (aiProcess_PreTransformVertices | Triangulate | JoinIdenticalVertices ...,
// src/scene_utils.cpp:102-112): output meshes are grouped by material index like
// PreTransformVertices does. assimp is an un-vendored dependency, so vertex order/merging is
// "parity unpinned" (SURVEY.md §8c); parity is defined on the arrays this loader produces.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <stdexcept>

#include "mini_json.hpp"
#include "scene_build.hpp"

namespace b200rt {

namespace {

struct Mat4 {
  float m[4][4];
  static Mat4 identity() {
    Mat4 r{};
    for (int i = 0; i < 4; ++i) r.m[i][i] = 1.f;
    return r;
  }
  Mat4 operator*(const Mat4& o) const {
    Mat4 r{};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += m[i][k] * o.m[k][j];
        r.m[i][j] = s;
      }
    return r;
  }
  Vec3 point(const Vec3& p) const {
    return Vec3{m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3],
                m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
                m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3]};
  }
  Vec3 dir(const Vec3& p) const {
    return Vec3{m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z, m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z,
                m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z};
  }
};

Mat4 nodeLocal(const mini_json::Value& node) {
  Mat4 M = Mat4::identity();
  if (const auto* mat = node.find("matrix")) {  // column-major in glTF
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) M.m[r][c] = (float)mat->at((size_t)(c * 4 + r)).num;
    return M;
  }
  float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
  if (const auto* v = node.find("translation")) for (int i = 0; i < 3; ++i) t[i] = (float)v->at((size_t)i).num;
  if (const auto* v = node.find("rotation")) for (int i = 0; i < 4; ++i) q[i] = (float)v->at((size_t)i).num;
  if (const auto* v = node.find("scale")) for (int i = 0; i < 3; ++i) s[i] = (float)v->at((size_t)i).num;
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - z * w), 2.f * (x * z + y * w)},
                         {2.f * (x * y + z * w), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - x * w)},
                         {2.f * (x * z - y * w), 2.f * (y * z + x * w), 1.f - 2.f * (x * x + y * y)}};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M.m[r][c] = R[r][c] * s[c];
    M.m[r][3] = t[r];
  }
  return M;
}

struct Glb {
  mini_json::Value json;
  std::vector<unsigned char> bin;
};

Glb readGlb(const std::string& file) {
  std::ifstream f(file, std::ios::binary);
  if (!f) throw std::runtime_error("Could not open mesh file '" + file + "'");
  std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  auto u32 = [&](size_t o) {
    if (o + 4 > d.size()) throw std::runtime_error("glb: truncated file");
    std::uint32_t v;
    std::memcpy(&v, d.data() + o, 4);
    return v;
  };
  if (d.size() < 20 || u32(0) != 0x46546C67u) throw std::runtime_error("glb: bad magic in '" + file + "'");
  Glb g;
  size_t off = 12;
  while (off + 8 <= d.size()) {
    const std::uint32_t len = u32(off), type = u32(off + 4);
    if (off + 8 + len > d.size()) throw std::runtime_error("glb: chunk overruns file");
    if (type == 0x4E4F534Au) {
      g.json = mini_json::parse(std::string((const char*)d.data() + off + 8, len));
    } else if (type == 0x004E4942u && g.bin.empty()) {
      g.bin.assign(d.begin() + (long)(off + 8), d.begin() + (long)(off + 8 + len));
    }
    off += 8 + ((len + 3u) & ~3u);
  }
  if (g.json.kind != mini_json::Value::Object) throw std::runtime_error("glb: no JSON chunk");
  return g;
}

struct AccessorView {
  const unsigned char* base;
  size_t count, stride;
  int componentType, numComponents;
};

AccessorView accessor(const Glb& g, size_t index) {
  const auto& a = g.json.at("accessors").at(index);
  const auto& bv = g.json.at("bufferViews").at((size_t)a.at("bufferView").num);
  const std::string& type = a.at("type").str;
  AccessorView v;
  v.componentType = (int)a.at("componentType").num;
  v.numComponents = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
  if (!v.numComponents) throw std::runtime_error("glb: unsupported accessor type " + type);
  const size_t compSize = v.componentType == 5126 || v.componentType == 5125 ? 4
                          : v.componentType == 5123 || v.componentType == 5122 ? 2 : 1;
  v.count = (size_t)a.at("count").num;
  const size_t offset = (size_t)(bv.find("byteOffset") ? bv.at("byteOffset").num : 0.0) +
                        (size_t)(a.find("byteOffset") ? a.at("byteOffset").num : 0.0);
  v.stride = bv.find("byteStride") ? (size_t)bv.at("byteStride").num : compSize * (size_t)v.numComponents;
  if (offset + (v.count ? (v.count - 1) * v.stride + compSize * (size_t)v.numComponents : 0) > g.bin.size())
    throw std::runtime_error("glb: accessor overruns buffer");
  v.base = g.bin.data() + offset;
  return v;
}

void walk(const Glb& g, size_t nodeIndex, const Mat4& parent, bool loadNormals,
          std::map<long, MeshParts>& byMaterial) {
  const auto& node = g.json.at("nodes").at(nodeIndex);
  const Mat4 M = parent * nodeLocal(node);
  if (const auto* meshRef = node.find("mesh")) {
    const auto& mesh = g.json.at("meshes").at((size_t)meshRef->num);
    for (const auto& prim : mesh.at("primitives").arr) {
      if (prim.find("mode") && (int)prim.at("mode").num != 4) throw std::runtime_error("Only triangle meshes are supported.");
      const long material = prim.find("material") ? (long)prim.at("material").num : -1;
      MeshParts& out = byMaterial[material];
      const size_t base = out.vertices.size();
      const AccessorView pos = accessor(g, (size_t)prim.at("attributes").at("POSITION").num);
      if (pos.componentType != 5126 || pos.numComponents != 3) throw std::runtime_error("glb: POSITION must be float VEC3");
      for (size_t i = 0; i < pos.count; ++i) {
        Vec3 p;
        std::memcpy(&p, pos.base + i * pos.stride, 12);
        out.vertices.push_back(M.point(p));
      }
      if (loadNormals && prim.at("attributes").find("NORMAL")) {
        const AccessorView nrm = accessor(g, (size_t)prim.at("attributes").at("NORMAL").num);
        for (size_t i = 0; i < nrm.count; ++i) {
          Vec3 p;
          std::memcpy(&p, nrm.base + i * nrm.stride, 12);
          out.normals.push_back(M.dir(p));
        }
      }
      std::vector<std::uint32_t> idx;
      if (prim.find("indices")) {
        const AccessorView iv = accessor(g, (size_t)prim.at("indices").num);
        idx.resize(iv.count);
        for (size_t i = 0; i < iv.count; ++i) {
          const unsigned char* p = iv.base + i * iv.stride;
          if (iv.componentType == 5123) { std::uint16_t v; std::memcpy(&v, p, 2); idx[i] = v; }
          else if (iv.componentType == 5125) { std::uint32_t v; std::memcpy(&v, p, 4); idx[i] = v; }
          else idx[i] = *p;
        }
      } else {
        idx.resize(pos.count);
        for (size_t i = 0; i < pos.count; ++i) idx[i] = (std::uint32_t)i;
      }
      if (base + pos.count > 65536) throw std::runtime_error("glb: mesh exceeds 65536 vertices (u16 indices)");
      for (size_t i = 0; i + 2 < idx.size(); i += 3)
        out.triangles.push_back(Triangle{(std::uint16_t)(base + idx[i]), (std::uint16_t)(base + idx[i + 1]),
                                         (std::uint16_t)(base + idx[i + 2])});
    }
  }
  if (const auto* kids = node.find("children"))
    for (const auto& k : kids->arr) walk(g, (size_t)k.num, M, loadNormals, byMaterial);
}

}  // namespace

// Reads every mesh of the default scene, one output mesh per material (ascending material index).
std::vector<MeshParts> readGlbMeshes(const std::string& file, bool loadNormals) {
  const Glb g = readGlb(file);
  const size_t sceneIndex = g.json.find("scene") ? (size_t)g.json.at("scene").num : 0;
  std::map<long, MeshParts> byMaterial;
  for (const auto& n : g.json.at("scenes").at(sceneIndex).at("nodes").arr)
    walk(g, (size_t)n.num, Mat4::identity(), loadNormals, byMaterial);
  std::vector<MeshParts> out;
  for (auto& kv : byMaterial) out.push_back(std::move(kv.second));
  return out;
}

// importMesh (src/scene_utils.cpp:102-149): scale each mesh so its bounding-box diagonal is 175,
// turn it to face the camera and put it on the short block. Normals are not loaded (:119).
void importMeshForBox(const std::string& file, std::vector<MeshParts>& meshes) {
  if (file.empty()) return;
  for (auto& mesh : readGlbMeshes(file, false)) {
    Vec3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY};
    for (const auto& v : mesh.vertices) {
      lo.x = std::fmin(lo.x, v.x); lo.y = std::fmin(lo.y, v.y); lo.z = std::fmin(lo.z, v.z);
      hi.x = std::fmax(hi.x, v.x); hi.y = std::fmax(hi.y, v.y); hi.z = std::fmax(hi.z, v.z);
    }
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    const float scale = 175.f / std::sqrt(dx * dx + dy * dy + dz * dz);
    for (auto& v : mesh.vertices) {
      v.x = -v.x;
      v.z = -v.z;
      v.x *= scale; v.y *= scale; v.z *= scale;
      v.x += 210.f; v.y += 165.f; v.z += 160.f;
    }
    meshes.push_back(std::move(mesh));
  }
}

}  // namespace b200rt
