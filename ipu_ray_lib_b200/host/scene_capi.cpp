// C API of the host-side scene utilities (include/b200rt_scene.h).
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../../include/b200rt_scene.h"
#include "../csrc/rt_math.h"
#include "mini_json.hpp"
#include "scene_build.hpp"

using namespace b200rt;

struct b200rt_host_scene {
  HostScene scene;
};

static thread_local std::string g_err;

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return B200RT_OK;
  } catch (const std::bad_alloc&) {
    g_err = "out of host memory";
    return B200RT_ERR_OOM;
  } catch (const std::exception& e) {
    g_err = e.what();
    return B200RT_ERR_INVALID_ARG;
  }
}

// ---- the reference's serialised SceneRef ----------------------------------------------------------------------
namespace {
// Every object sits at its own alignment relative to a 16-byte aligned base (Serialiser::calculatePadding).
struct BlobCursor {
  const unsigned char* base;
  size_t size, at = 0;
  void align(size_t a) {
    const size_t rem = (16 + at) % a;
    if (rem) at += a - rem;
  }
  template <class T> T scalar() {
    align(alignof(T));
    if (at + sizeof(T) > size) throw std::runtime_error("serialised scene is truncated");
    T v;
    std::memcpy(&v, base + at, sizeof(T));
    at += sizeof(T);
    return v;
  }
  const void* array(uint32_t& count, size_t elemSize, size_t elemAlign) {
    count = scalar<uint32_t>();
    align(elemAlign);
    if (at + (size_t)count * elemSize > size) throw std::runtime_error("serialised scene is truncated");
    const void* p = base + at;
    at += (size_t)count * elemSize;
    return p;
  }
};
struct BlobSink {
  unsigned char* out;
  size_t cap, at = 0;
  void put(const void* src, size_t n) {
    if (out && at + n <= cap) std::memcpy(out + at, src, n);
    at += n;
  }
  void align(size_t a) {
    static const unsigned char zeros[16] = {0};
    const size_t rem = (16 + at) % a;
    if (rem) put(zeros, a - rem);
  }
  template <class T> void scalar(const T& v) { align(alignof(T)); put(&v, sizeof(T)); }
  void array(const void* data, uint32_t count, size_t elemSize, size_t elemAlign) {
    scalar(count);
    align(elemAlign);
    put(data, (size_t)count * elemSize);
  }
};
}  // namespace

extern "C" {

const char* b200rt_scene_last_error(void) { return g_err.c_str(); }

int b200rt_host_scene_builtin(const char* name, const char* meshFile, b200rt_host_scene** out) {
  return guarded([&] {
    if (!name || !out) throw std::runtime_error("null argument");
    const std::string n(name);
    SceneParts parts;
    if (n == "box" || n == "box-simple") {
      parts = makeCornellBoxScene(meshFile ? meshFile : "", n == "box-simple");
    } else if (n == "spheres") {
      parts = makePrimitiveScene();
    } else {
      throw std::runtime_error("Invalid scene selection: '" + n + "'");
    }
    auto* h = new b200rt_host_scene();
    try { finaliseScene(parts, h->scene); } catch (...) { delete h; throw; }
    *out = h;
  });
}

int b200rt_host_scene_import(const char* file, int loadNormals, b200rt_host_scene** out) {
  return guarded([&] {
    if (!file || !out) throw std::runtime_error("null argument");
    SceneParts parts = importScene(file, loadNormals != 0);
    auto* h = new b200rt_host_scene();
    try { finaliseScene(parts, h->scene); } catch (...) { delete h; throw; }
    *out = h;
  });
}

void b200rt_host_scene_free(b200rt_host_scene* s) { delete s; }

int b200rt_host_scene_desc(const b200rt_host_scene* s, b200rt_scene_desc* d) {
  return guarded([&] {
    if (!s || !d) throw std::runtime_error("null argument");
    const HostScene& h = s->scene;
    *d = b200rt_scene_desc{};
    d->geometry = h.geometry.data();      d->num_geometry = (uint32_t)h.geometry.size();
    d->mesh_info = h.meshInfo.data();     d->num_meshes = (uint32_t)h.meshInfo.size();
    d->mesh_tris = h.meshTris.data();     d->num_tris = (uint32_t)h.meshTris.size();
    d->mesh_verts = h.meshVerts.data();   d->num_verts = (uint32_t)h.meshVerts.size();
    d->mesh_normals = h.meshNormals.data(); d->num_normals = (uint32_t)h.meshNormals.size();
    d->mat_ids = h.matIDs.data();         d->num_mat_ids = (uint32_t)h.matIDs.size();
    d->materials = h.materials.data();    d->num_materials = (uint32_t)h.materials.size();
    d->bvh_nodes = h.bvhNodes.data();     d->num_bvh_nodes = (uint32_t)h.bvhNodes.size();
    d->max_leaf_depth = h.bvhMaxDepth;
    d->spheres = (const float*)h.spheres.data(); d->num_spheres = (uint32_t)h.spheres.size();
    d->discs = (const float*)h.discs.data();     d->num_discs = (uint32_t)h.discs.size();
    d->fov_radians = h.horizontalFov;
    d->anti_alias_scale = .25f;
    d->max_path_length = 10;
    d->roulette_start_depth = 3;
    d->samples_per_pixel = 256;
    d->rng_seed = 1442;
    d->path_trace = 1;
    d->device = -1;
  });
}

int b200rt_build_bvh(const float* primBounds, const uint32_t* ids, uint32_t n, void* nodesOut, uint32_t* maxDepthOut) {
  int count = -1;
  const int rc = guarded([&] {
    if (!primBounds || !ids || !nodesOut) throw std::runtime_error("null argument");
    std::vector<BvhNode> nodes;
    const uint32_t depth = buildCompactBvh(primBounds, ids, n, nodes);
    std::memcpy(nodesOut, nodes.data(), nodes.size() * sizeof(BvhNode));
    if (maxDepthOut) *maxDepthOut = depth;
    count = (int)nodes.size();
  });
  return rc == B200RT_OK ? count : rc;
}

int b200rt_init_ray_stream(void* rays, int imgW, int imgH, int winW, int winH, int winC, int winR, float fov) {
  return guarded([&] {
    if (!rays || imgW <= 0 || imgH <= 0 || winW <= 0 || winH <= 0 || winC < 0 || winR < 0)
      throw std::runtime_error("bad ray stream window");
    initPerspectiveRayStream((TraceResult*)rays, imgW, imgH, CropWindow{winW, winH, winC, winR}, fov);
  });
}

void b200rt_scale_rgb(void* rays, size_t n, float scale) {
  auto* r = (TraceResult*)rays;
  for (size_t i = 0; i < n; ++i) { r[i].rgb.x *= scale; r[i].rgb.y *= scale; r[i].rgb.z *= scale; }
}

long b200rt_visualise_hits(const void* rays, size_t n, const b200rt_scene_desc* scene, int mode, float* image,
                           int imgW, int imgH) {
  long hits = -1;
  guarded([&] {
    if (!rays || !scene || !image) throw std::runtime_error("null argument");
    hits = visualiseHits((const TraceResult*)rays, n, *scene, mode, image, imgW, imgH);
  });
  return hits;
}

int b200rt_write_exr(const char* path, const float* image, int w, int h) {
  return guarded([&] { writeExr(path, image, w, h); });
}
int b200rt_write_pfm(const char* path, const float* image, int w, int h) {
  return guarded([&] { writePfm(path, image, w, h); });
}

int b200rt_read_nif_metadata(const char* path, b200rt_nif_metadata* out) {
  return guarded([&] {
    if (!path || !out) throw std::runtime_error("null argument");
    std::ifstream f(path);
    if (!f) throw std::runtime_error(std::string("Could not open '") + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    const auto j = mini_json::parse(ss.str());
    *out = b200rt_nif_metadata{};
    out->embedding_dimension = (uint32_t)j.at("embedding_dimension").num;
    const auto& shape = j.at("original_image_shape");
    for (size_t i = 0; i < 3 && i < shape.size(); ++i) out->image_shape[i] = (uint32_t)shape.at(i).num;
    const auto& enc = j.at("encode_params");
    out->eps = (float)enc.at("eps").num;
    out->log_tone_map = enc.at("log_tone_map").b ? 1 : 0;
    out->max = (float)enc.at("max").num;
    for (size_t i = 0; i < 3; ++i) out->mean[i] = (float)enc.at("mean").at(i).num;
    if (out->log_tone_map) for (float& m : out->mean) m -= out->eps;  // NifMetaData.cpp:48-53
    bool next = false;
    for (const auto& tok : j.at("train_command").arr) {
      if (next) { out->hidden_size = (uint32_t)std::atoi(tok.str.c_str()); next = false; }
      if (tok.str == "--layer-size") next = true;
    }
  });
}

int b200rt_scene_desc_from_blob(const void* blob, size_t bytes, b200rt_scene_desc* out) {
  return guarded([&] {
    if (!blob || !out) throw std::runtime_error("null argument");
    if (reinterpret_cast<uintptr_t>(blob) % 16) throw std::runtime_error("serialised scene must be 16-byte aligned");
    BlobCursor c{static_cast<const unsigned char*>(blob), bytes};
    b200rt_scene_desc d{};
    d.geometry = c.array(d.num_geometry, sizeof(GeomRef), alignof(GeomRef));
    d.mesh_info = c.array(d.num_meshes, sizeof(MeshInfo), alignof(MeshInfo));
    d.mesh_tris = c.array(d.num_tris, sizeof(Triangle), alignof(Triangle));
    d.mesh_verts = c.array(d.num_verts, sizeof(Vec3), alignof(Vec3));
    d.mesh_normals = c.array(d.num_normals, sizeof(Vec3), alignof(Vec3));
    d.mat_ids = static_cast<const uint32_t*>(c.array(d.num_mat_ids, 4, 4));
    d.materials = c.array(d.num_materials, sizeof(Material), alignof(Material));
    d.bvh_nodes = c.array(d.num_bvh_nodes, sizeof(BvhNode), 4);  // CompactBVH2Node: 4-byte aligned in the stream
    d.max_leaf_depth = c.scalar<uint32_t>();
    d.image_width = c.scalar<float>();
    d.image_height = c.scalar<float>();
    d.fov_radians = c.scalar<float>();
    d.anti_alias_scale = c.scalar<float>();
    d.max_path_length = c.scalar<uint32_t>();
    d.roulette_start_depth = c.scalar<uint32_t>();
    d.samples_per_pixel = c.scalar<uint32_t>();
    if (c.at != bytes) throw std::runtime_error("serialised scene has trailing bytes");
    d.device = -1;
    *out = d;
  });
}

size_t b200rt_scene_blob_write(const b200rt_scene_desc* d, void* out, size_t cap) {
  if (!d) return 0;
  BlobSink w{static_cast<unsigned char*>(out), cap};
  w.array(d->geometry, d->num_geometry, sizeof(GeomRef), alignof(GeomRef));
  w.array(d->mesh_info, d->num_meshes, sizeof(MeshInfo), alignof(MeshInfo));
  w.array(d->mesh_tris, d->num_tris, sizeof(Triangle), alignof(Triangle));
  w.array(d->mesh_verts, d->num_verts, sizeof(Vec3), alignof(Vec3));
  w.array(d->mesh_normals, d->num_normals, sizeof(Vec3), alignof(Vec3));
  w.array(d->mat_ids, d->num_mat_ids, 4, 4);
  w.array(d->materials, d->num_materials, sizeof(Material), alignof(Material));
  w.array(d->bvh_nodes, d->num_bvh_nodes, sizeof(BvhNode), 4);
  w.scalar(d->max_leaf_depth);
  w.scalar(d->image_width);
  w.scalar(d->image_height);
  w.scalar(d->fov_radians);
  w.scalar(d->anti_alias_scale);
  w.scalar(d->max_path_length);
  w.scalar(d->roulette_start_depth);
  w.scalar(d->samples_per_pixel);
  return w.at;
}

void b200rt_sincos(float x, float* s, float* c) { rt::sincos_tbl(x, *s, *c); }

}  // extern "C"
