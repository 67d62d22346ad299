// ref_driver.cpp — thin driver around the REFERENCE'S OWN kernel sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h). This file is compiled together with
//   /root/reference/src/Mesh.cpp  src/Primitives.cpp  src/CompactBVH2Node.cpp  ext/math/sincos.cpp
// and includes the reference headers from where they lie (never copied into this repo), so
// bvh.intersect/occluded, the primitive tests, updateHit/offsetRay/traceShadowRay, the BxDFs,
// sincos and xoshiro below are the reference's code, not a restatement. Only what cannot be
// compiled here (trace.cpp drags in Poplar/OpenCV/boost/Embree) is restated: the pathTrace
// bounce loop (trace.cpp:115-188) and the per-sample camera ray (codelets/TraceCodelets.cpp:142-164),
// both with the build-defined per-(pixel,sample) RNG streams.
#define ORC_PREFIX ref_
#include "oracle_api.h"

#include <Arrays.hpp>
#include <Primitives.hpp>
#include <Mesh.hpp>
#include <Scene.hpp>
#include <CompactBvh.hpp>
#include <Material.hpp>
#include <Render.hpp>
#include <BxDF.hpp>
#include <xoshiro.hpp>
#include <math/sincos.hpp>

#include <omp.h>
#include <random>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include <serialisation/serialisation.hpp>

using embree_utils::HitRecord;
using embree_utils::Ray;
using embree_utils::TraceResult;
using embree_utils::Vec3fa;

static_assert(sizeof(TraceResult) == 84, "TraceResult");
static_assert(sizeof(CompactBVH2Node) == 24, "CompactBVH2Node");
static_assert(sizeof(Material) == 36, "Material");
static_assert(sizeof(GeomRef) == 4 && sizeof(MeshInfo) == 16 && sizeof(Triangle) == 6, "scene records");

namespace {

// The reference objects a render needs, built from the flat C-ABI arrays the way
// renderCPU builds them from SceneRef (trace.cpp:197-230).
struct RefScene {
  std::vector<CompiledTriangleMesh> meshes;
  std::vector<Sphere> spheres;
  std::vector<Disc> discs;
  ArrayRef<GeomRef> geometry;
  ArrayRef<std::uint32_t> matIDs;
  ArrayRef<Material> materials;
  ArrayRef<CompactBVH2Node> nodes;
  std::uint32_t maxDepth;

  explicit RefScene(const b200rt_scene_desc& d)
      : geometry((GeomRef*)d.geometry, d.num_geometry),
        matIDs((std::uint32_t*)d.mat_ids, d.num_mat_ids),
        materials((Material*)d.materials, d.num_materials),
        nodes((CompactBVH2Node*)d.bvh_nodes, d.num_bvh_nodes),
        maxDepth(d.max_leaf_depth) {
    auto* info = (const MeshInfo*)d.mesh_info;
    auto* tris = (Triangle*)d.mesh_tris;
    auto* verts = (Vec3fa*)d.mesh_verts;
    auto* normals = (Vec3fa*)d.mesh_normals;
    meshes.reserve(d.num_meshes);
    for (std::uint32_t m = 0; m < d.num_meshes; ++m) {
      std::uint32_t firstNormal = 0, numNormals = 0;
      if (d.num_normals) { firstNormal = info[m].firstVertex; numNormals = info[m].numVertices; }
      meshes.emplace_back(embree_utils::Bounds3d(),
                          ArrayRef<Triangle>(tris + info[m].firstIndex, info[m].numTriangles),
                          ArrayRef<Vec3fa>(verts + info[m].firstVertex, info[m].numVertices),
                          ArrayRef<Vec3fa>(normals + firstNormal, numNormals));
    }
    spheres.reserve(d.num_spheres);
    for (std::uint32_t i = 0; i < d.num_spheres; ++i) {
      const float* s = d.spheres + 4 * i;
      spheres.emplace_back(Vec3fa(s[0], s[1], s[2]), s[3]);
    }
    discs.reserve(d.num_discs);
    for (std::uint32_t i = 0; i < d.num_discs; ++i) {
      const float* p = d.discs + 7 * i;
      discs.emplace_back(Vec3fa(p[0], p[1], p[2]), Vec3fa(p[4], p[5], p[6]), p[3]);
    }
  }

  const Primitive* prim(std::uint16_t geomID) const {
    const GeomRef& g = geometry[geomID];
    switch (g.type) {
      case GeomType::Mesh: return &meshes[g.index];
      case GeomType::Sphere: return &spheres[g.index];
      case GeomType::Disc: return &discs[g.index];
      default: return nullptr;
    }
  }
};

struct Counters {
  std::uint64_t closest = 0, occl = 0, prims = 0, samples = 0, escaped = 0;
  void flush(std::uint64_t* out) const {
    if (!out) return;
#pragma omp atomic
    out[0] += closest;
#pragma omp atomic
    out[1] += occl;
#pragma omp atomic
    out[3] += prims;
#pragma omp atomic
    out[4] += samples;
#pragma omp atomic
    out[5] += escaped;
  }
};

int pickThreads(int threads) { return threads > 0 ? threads : omp_get_max_threads(); }

// --- build-defined RNG stream and Gaussian; mirrors ipu_ray_lib_b200/csrc/rt_math.h ---
inline void seedStream(xoshiro::State& st, std::uint64_t key, std::uint32_t pixelIndex, std::uint32_t sample) {
  const std::uint64_t id = ((std::uint64_t)pixelIndex << 32) | (std::uint64_t)sample;
  xoshiro::seed(st, xoshiro::splitmix64(id ^ key));
}

inline float detLog(float x) {
  std::uint32_t u;
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

inline void gaussianPair(std::uint64_t a, std::uint64_t b, float& g0, float& g1) {
  const float u1 = (float)((std::uint32_t)(a >> 40) + 1u) * 5.9604644775390625e-08f;
  const float u2 = (float)((std::uint32_t)(b >> 40)) * 5.9604644775390625e-08f;
  const float rad = std::sqrt(-2.f * detLog(u1));
  float s, c;
  sincos(6.28318530717958647692f * u2, s, c);  // the reference's own sincos
  g0 = rad * c;
  g1 = rad * s;
}

inline Vec3fa cameraSample(xoshiro::State& st, float row, float col, float w, float h, float tanTheta, float aa) {
  const std::uint64_t a = xoshiro::next128ss(st);
  const std::uint64_t b = xoshiro::next128ss(st);
  float g0, g1;
  gaussianPair(a, b, g0, g1);
  const float pu = row + aa * g0;
  const float pv = col + aa * g1;
  return pixelToRayDir(pv, pu, w, h, tanTheta);  // reference Render.hpp:74-85
}

inline float fovTan(float fovRadians) {
  float s, c;
  sincos(fovRadians / 2.f, s, c);
  return s / c;
}

}  // namespace

extern "C" {

size_t ref_serialise_scene(const b200rt_scene_desc* d, uint8_t* out, size_t cap) {
  // the reference's own writer, fed a SceneRef over the caller's arrays (src/IpuScene.cpp:31, :52)
  SceneRef ref;
  ref.geometry = ArrayRef<GeomRef>((GeomRef*)d->geometry, d->num_geometry);
  ref.meshInfo = ArrayRef<MeshInfo>((MeshInfo*)d->mesh_info, d->num_meshes);
  ref.meshTris = ArrayRef<Triangle>((Triangle*)d->mesh_tris, d->num_tris);
  ref.meshVerts = ArrayRef<Vec3fa>((Vec3fa*)d->mesh_verts, d->num_verts);
  ref.meshNormals = ArrayRef<Vec3fa>((Vec3fa*)d->mesh_normals, d->num_normals);
  ref.matIDs = ArrayRef<std::uint32_t>((std::uint32_t*)d->mat_ids, d->num_mat_ids);
  ref.materials = ArrayRef<Material>((Material*)d->materials, d->num_materials);
  ref.bvhNodes = ArrayRef<CompactBVH2Node>((CompactBVH2Node*)d->bvh_nodes, d->num_bvh_nodes);
  ref.maxLeafDepth = d->max_leaf_depth;
  ref.imageWidth = d->image_width; ref.imageHeight = d->image_height;
  ref.fovRadians = d->fov_radians; ref.antiAliasScale = d->anti_alias_scale;
  ref.maxPathLength = d->max_path_length; ref.rouletteStartDepth = d->roulette_start_depth;
  ref.samplesPerPixel = d->samples_per_pixel;
  Serialiser<16> ser(600 * 1024);
  ser << ref;
  if (out && cap >= ser.bytes.size()) std::memcpy(out, ser.bytes.data(), ser.bytes.size());
  return ser.bytes.size();
}

const char* ref_kind(void) { return "reference"; }

int ref_shadow_trace(const b200rt_scene_desc* d, void* raysV, size_t n, const float lp[3], float ambient,
                     int threads, uint64_t* counters) {
  RefScene sc(*d);
  CompactBvh bvh(sc.nodes, sc.maxDepth);
  auto* rays = (TraceResult*)raysV;
  const Vec3fa light(lp[0], lp[1], lp[2]);
#pragma omp parallel num_threads(pickThreads(threads))
  {
    Counters cnt;
    auto lookup = [&](std::uint16_t geomID, std::uint32_t) { cnt.prims += 1; return sc.prim(geomID); };
#pragma omp for schedule(dynamic, 512)
    for (long long i = 0; i < (long long)n; ++i) {
      cnt.closest += 1;
      const bool wasHit = rays[i].h.geomID != HitRecord::InvalidGeomID;
      traceShadowRay(bvh, sc.matIDs, sc.materials, ambient, rays[i], lookup, light);
      if (!wasHit && rays[i].h.geomID != HitRecord::InvalidGeomID) cnt.occl += 1;
    }
    cnt.flush(counters);
  }
  return 0;
}

int ref_path_trace(const b200rt_scene_desc* d, void* raysV, size_t n, uint32_t firstSample, uint32_t numSamples,
                   const b200rt_nif_desc* /*nif: the reference CPU path has none (trace.cpp:171-174)*/,
                   float /*hdri*/, int threads, uint64_t* counters) {
  RefScene sc(*d);
  CompactBvh bvh(sc.nodes, sc.maxDepth);
  auto* rays = (TraceResult*)raysV;
  const float tanTheta = fovTan(d->fov_radians);
  const std::uint64_t key = xoshiro::splitmix64(d->rng_seed);
  const auto imgW = (std::uint32_t)d->image_width;
#pragma omp parallel num_threads(pickThreads(threads))
  {
    Counters cnt;
    auto lookup = [&](std::uint16_t geomID, std::uint32_t) { cnt.prims += 1; return sc.prim(geomID); };
#pragma omp for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)n; ++i) {
      TraceResult& result = rays[i];
      const auto row = (std::uint32_t)result.p.u, col = (std::uint32_t)result.p.v;
      for (std::uint32_t s = firstSample; s < firstSample + numSamples; ++s) {
        xoshiro::State st;
        seedStream(st, key, row * imgW + col, s);
        const Vec3fa dir = cameraSample(st, result.p.u, result.p.v, d->image_width, d->image_height, tanTheta,
                                        d->anti_alias_scale);
        result.h = HitRecord(Vec3fa(0.f, 0.f, 0.f), dir);
        cnt.samples += 1;

        // --- the bounce loop of trace.cpp:115-188, RNG draws taken from the per-sample stream ---
        auto& hit = result.h;
        hit.throughput = Vec3fa(1.f, 1.f, 1.f);
        Vec3fa color(0.f, 0.f, 0.f);
        for (std::uint32_t i2 = 0; i2 < d->max_path_length; ++i2) {
          offsetRay(hit.r, hit.normal);
          hit.r.tMin = 0.f;
          hit.r.tMax = std::numeric_limits<float>::infinity();
          cnt.closest += 1;
          auto isect = bvh.intersect(hit.r, lookup);
          if (isect) {
            updateHit(isect, hit);
            const Material& material = sc.materials[sc.matIDs[hit.geomID]];
            if (material.emissive) { color += hit.throughput * material.emission; }
            if (material.type == Material::Type::Diffuse) {
              const float u1 = xoshiro::uniform_0_1(st);
              const float u2 = xoshiro::uniform_0_1(st);
              hit.r.direction = sampleDiffuse(hit.normal, u1, u2);
              hit.throughput *= material.albedo;
            } else if (material.type == Material::Type::Specular) {
              hit.r.direction = reflect(hit.r.direction, hit.normal);
              hit.throughput *= material.albedo;
            } else if (material.type == Material::Type::Refractive) {
              const float u1 = xoshiro::uniform_0_1(st);
              const auto [ndir, refracted] = dielectric(hit.r, hit.normal, material.ior, u1);
              hit.r.direction = ndir;
              if (refracted) { hit.throughput *= material.albedo; }
            } else {
              result.rgb *= std::numeric_limits<float>::quiet_NaN();
              hit.flags |= HitRecord::ERROR;
            }
          } else {
            hit.flags |= HitRecord::ESCAPED;
            cnt.escaped += 1;
            break;
          }
          if (i2 > d->roulette_start_depth) {
            const float u1 = xoshiro::uniform_0_1(st);
            if (evaluateRoulette(u1, hit.throughput)) { break; }
          }
        }
        result.rgb += color;
      }
    }
    cnt.flush(counters);
  }
  return 0;
}

int ref_intersect(const b200rt_scene_desc* d, const void* raysV, size_t n, b200rt_hit* out, int threads,
                  uint64_t* counters) {
  RefScene sc(*d);
  CompactBvh bvh(sc.nodes, sc.maxDepth);
  auto* rays = (const Ray*)raysV;
#pragma omp parallel num_threads(pickThreads(threads))
  {
    Counters cnt;
    auto lookup = [&](std::uint16_t geomID, std::uint32_t) { cnt.prims += 1; return sc.prim(geomID); };
#pragma omp for schedule(dynamic, 512)
    for (long long i = 0; i < (long long)n; ++i) {
      cnt.closest += 1;
      auto isect = bvh.intersect(rays[i], lookup);
      b200rt_hit h;
      if (isect) {
        // what updateHit (Render.hpp:15-23) would record
        const Vec3fa p = rays[i].origin + rays[i].direction * isect.t;
        const Vec3fa nrm = isect.prim->normal(isect, p);
        h.t = isect.t;
        h.geom_id = isect.geomID;
        h.prim_id = isect.primID;
        h.normal[0] = nrm.x; h.normal[1] = nrm.y; h.normal[2] = nrm.z;
      } else {
        h.t = rays[i].tMax;
        h.geom_id = 0xFFFFu;
        h.prim_id = 0xFFFFFFFFu;
        h.normal[0] = h.normal[1] = h.normal[2] = 0.f;
      }
      out[i] = h;
    }
    cnt.flush(counters);
  }
  return 0;
}

int ref_occluded(const b200rt_scene_desc* d, const void* raysV, size_t n, uint8_t* out, int threads) {
  RefScene sc(*d);
  CompactBvh bvh(sc.nodes, sc.maxDepth);
  auto* rays = (const Ray*)raysV;
#pragma omp parallel num_threads(pickThreads(threads))
  {
    auto lookup = [&](std::uint16_t geomID, std::uint32_t) { return sc.prim(geomID); };
#pragma omp for schedule(dynamic, 512)
    for (long long i = 0; i < (long long)n; ++i) { out[i] = bvh.occluded(rays[i], lookup) ? 1 : 0; }
  }
  return 0;
}

void ref_sincos(const float* x, size_t n, float* s, float* c) {
  for (size_t i = 0; i < n; ++i) sincos(x[i], s[i], c[i]);
}

void ref_uniform_stream(uint64_t seed, size_t n, float* out) {
  xoshiro::Generator g(seed);
  for (size_t i = 0; i < n; ++i) out[i] = g.uniform_0_1();
}

void ref_raw_stream(uint64_t seed, size_t n, uint64_t* out) {
  xoshiro::Generator g(seed);
  for (size_t i = 0; i < n; ++i) out[i] = g();
}

void ref_sample_diffuse(const float* nrm, const float* u12, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    const Vec3fa d = sampleDiffuse(Vec3fa(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), u12[2 * i], u12[2 * i + 1]);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}

void ref_dielectric(const float* dirs, const float* nrm, const float* iorU1, size_t n, float* out, uint8_t* refr) {
  for (size_t i = 0; i < n; ++i) {
    Ray r(Vec3fa(0.f, 0.f, 0.f), Vec3fa(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
    const auto [d, refracted] =
        dielectric(r, Vec3fa(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), iorU1[2 * i], iorU1[2 * i + 1]);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    refr[i] = refracted ? 1 : 0;
  }
}

void ref_reflect(const float* dirs, const float* nrm, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    const Vec3fa d = reflect(Vec3fa(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]),
                             Vec3fa(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}

void ref_offset_ray(const float* org, const float* dirs, const float* nrm, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    Ray r(Vec3fa(org[3 * i], org[3 * i + 1], org[3 * i + 2]), Vec3fa(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
    offsetRay(r, Vec3fa(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    out[3 * i] = r.origin.x; out[3 * i + 1] = r.origin.y; out[3 * i + 2] = r.origin.z;
  }
}

void ref_pixel_to_ray_dir(const float* xy, size_t n, float w, float h, float tanTheta, float* out) {
  for (size_t i = 0; i < n; ++i) {
    const Vec3fa d = pixelToRayDir(xy[2 * i], xy[2 * i + 1], w, h, tanTheta);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}

void ref_round_to_half_not_smaller(const float* x, size_t n, uint16_t* out) {
  for (size_t i = 0; i < n; ++i) {
    half h = roundToHalfNotSmaller(x[i]);
    std::memcpy(out + i, &h, 2);
  }
}

void ref_camera_sample(uint64_t rngSeed, uint32_t w, uint32_t h, float fov, float aa, const uint32_t* rcs, size_t n,
                       float* out) {
  const float tanTheta = fovTan(fov);
  const std::uint64_t key = xoshiro::splitmix64(rngSeed);
  for (size_t i = 0; i < n; ++i) {
    xoshiro::State st;
    seedStream(st, key, rcs[3 * i] * w + rcs[3 * i + 1], rcs[3 * i + 2]);
    const Vec3fa d = cameraSample(st, (float)rcs[3 * i], (float)rcs[3 * i + 1], (float)w, (float)h, tanTheta, aa);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}

// renderCPU's path-trace loop AS SHIPPED (trace.cpp:236-245 around pathTrace :115-188): ONE global xoshiro::Generator,
// camera rays regenerated serially before every sample with std::normal_distribution jitter (initPerspectiveRayStream,
// src/app_utils.cpp:19-47), every draw of the bounce loop inside `omp critical(sample)` (trace.cpp:144-148, :161-165,
// :176-181), `omp parallel for schedule(auto)` over the rays. The order of the draws depends on the OpenMP schedule, so
// the image is not reproducible: this entry exists to TIME the reference's CPU path the way it ships (SURVEY.md 8d
// "(A)"), next to ref_path_trace, which is the same kernels with one RNG stream per (pixel, sample) and no lock.
int ref_path_trace_as_shipped(const b200rt_scene_desc* d, void* raysV, size_t n, uint32_t numSamples, int threads,
                              uint64_t* counters) {
  RefScene sc(*d);
  CompactBvh bvh(sc.nodes, sc.maxDepth);
  auto* rays = (TraceResult*)raysV;
  const float tanTheta = fovTan(d->fov_radians);
  xoshiro::Generator sampler(d->rng_seed);  // makePathTraceSettings, src/app_utils.cpp:237-239
  std::normal_distribution<float> jitter{0.f, d->anti_alias_scale};
  Counters total;
  for (std::uint32_t s = 0; s < numSamples; ++s) {
    for (size_t i = 0; i < n; ++i) {  // serial, like the reference
      const float pu = rays[i].p.u + jitter(sampler), pv = rays[i].p.v + jitter(sampler);
      rays[i].h = HitRecord(Vec3fa(0.f, 0.f, 0.f), pixelToRayDir(pv, pu, d->image_width, d->image_height, tanTheta));
    }
#pragma omp parallel num_threads(pickThreads(threads))
    {
      Counters cnt;
      auto lookup = [&](std::uint16_t geomID, std::uint32_t) { cnt.prims += 1; return sc.prim(geomID); };
#pragma omp for schedule(auto)
      for (long long i = 0; i < (long long)n; ++i) {
        TraceResult& result = rays[i];
        auto& hit = result.h;
        cnt.samples += 1;
        hit.throughput = Vec3fa(1.f, 1.f, 1.f);
        Vec3fa color(0.f, 0.f, 0.f);
        for (std::uint32_t i2 = 0; i2 < d->max_path_length; ++i2) {
          offsetRay(hit.r, hit.normal);
          hit.r.tMin = 0.f;
          hit.r.tMax = std::numeric_limits<float>::infinity();
          cnt.closest += 1;
          auto isect = bvh.intersect(hit.r, lookup);
          if (isect) {
            updateHit(isect, hit);
            const Material& material = sc.materials[sc.matIDs[hit.geomID]];
            if (material.emissive) { color += hit.throughput * material.emission; }
            if (material.type == Material::Type::Diffuse) {
              float u1, u2;
#pragma omp critical(sample)
              {
                u1 = sampler.uniform_0_1();
                u2 = sampler.uniform_0_1();
              }
              hit.r.direction = sampleDiffuse(hit.normal, u1, u2);
              hit.throughput *= material.albedo;
            } else if (material.type == Material::Type::Specular) {
              hit.r.direction = reflect(hit.r.direction, hit.normal);
              hit.throughput *= material.albedo;
            } else if (material.type == Material::Type::Refractive) {
              float u1;
#pragma omp critical(sample)
              { u1 = sampler.uniform_0_1(); }
              const auto [ndir, refracted] = dielectric(hit.r, hit.normal, material.ior, u1);
              hit.r.direction = ndir;
              if (refracted) { hit.throughput *= material.albedo; }
            } else {
              result.rgb *= std::numeric_limits<float>::quiet_NaN();
              hit.flags |= HitRecord::ERROR;
            }
          } else {
            hit.flags |= HitRecord::ESCAPED;
            cnt.escaped += 1;
            break;
          }
          if (i2 > d->roulette_start_depth) {
            float u1;
#pragma omp critical(sample)
            { u1 = sampler.uniform_0_1(); }
            if (evaluateRoulette(u1, hit.throughput)) { break; }
          }
        }
        result.rgb += color;
      }
      cnt.flush(counters);
    }
  }
  return 0;
}

int ref_nif_eval(const b200rt_nif_desc*, const float*, size_t, float*, int) { return -4; /* no CPU NIF in the reference */ }
void ref_dir_to_uv(const float*, size_t, float, float*) {}

}  // extern "C"
