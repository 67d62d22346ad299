// B200Scene — C++ mirror of the reference's `IpuScene` (include/IpuScene.hpp:22-112) over the C ABI.
//
// Same public surface, same call order as renderIPU (trace.cpp:270-336):
//     B200Scene scene(spheres, discs, sceneRef, rayStream, raysPerWorker, callbackPtr);
//     scene.setRuntimeConfig({numGpus});            // ipus -> GPUs, one replica per device
//     scene.loadNifModel(path); scene.setHdriRotation(deg); scene.setMaxNifBatchSize(n);
//     int rc = scene.run();                         // GraphManager().run(scene, opts)
//     scene.getTraceTimeSecs();
// Batching follows src/IpuScene.cpp:78-172: the stream is cut into batches of
// raysPerWorker * 6 * 1440 rays, batch i belongs to replica i % R, and a registered callback sees
// (batchIndex, batch) with the reference's numbering. Replicas run concurrently, one host thread per
// GPU; results are written back in place into the caller's ray stream.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200rt.h"
#include "../../include/b200rt_scene.h"
#include "rt_types.hpp"

namespace b200rt {

// `SceneRef` (include/Scene.hpp:50-74): non-owning views + render parameters.
struct SceneRef {
  const HostScene* data = nullptr;
  float imageWidth = 0, imageHeight = 0, fovRadians = 0, antiAliasScale = .25f;
  std::uint32_t maxPathLength = 10, rouletteStartDepth = 3, samplesPerPixel = 256;
  std::uint64_t rngSeed = 1442;
  CropWindow window{0, 0, 0, 0};
  bool pathTrace = true;
};

struct RuntimeConfig {  // ipu_utils::RuntimeConfig (include/ipu_utils.hpp:174-183), the fields that still mean something
  std::uint32_t numGpus = 1;
  std::uint32_t numReplicas = 1;
};

// NIF weights in the library's own container (the reference's Keras HDF5 needs libhdf5, absent here):
//   "B2NF" u32 version=1 u32 embedding u32 numLayers f32 max f32 mean[3] u32 logToneMap
//   per layer: u32 in, u32 out, u32 relu, u32 hasBias, fp16 kernel[in*out], fp16 bias[out]
struct NifWeightsFile {
  std::uint32_t embedding = 12;
  float max = 0.f, mean[3] = {0, 0, 0};
  std::uint32_t logToneMap = 1;
  struct Layer { std::uint32_t in, out, relu; std::vector<std::uint16_t> kernel, bias; };
  std::vector<Layer> layers;

  // The reference's own container: the Keras h5 model `converted.hdf5` (src/IpuScene.cpp:177, src/keras/Hdf5Model.cpp),
  // read by the self-contained reader in host/keras_hdf5.cpp.
  static NifWeightsFile load(const std::string& path) {
    b200rt_keras_model* m = nullptr;
    if (b200rt_keras_hdf5_open(path.c_str(), &m) != 0) throw std::runtime_error(b200rt_keras_last_error());
    struct Close { b200rt_keras_model* m; ~Close() { b200rt_keras_hdf5_close(m); } } close{m};
    NifWeightsFile w;
    const std::uint32_t n = b200rt_keras_hdf5_num_layers(m);
    if (n == 0 || n > 16) throw std::runtime_error("implausible number of Dense layers in '" + path + "'");
    w.layers.resize(n);
    for (std::uint32_t i = 0; i < n; ++i) {
      b200rt_keras_layer kl;
      if (b200rt_keras_hdf5_layer(m, i, &kl) != 0) throw std::runtime_error(b200rt_keras_last_error());
      Layer& L = w.layers[i];
      L.in = kl.layer.in_features; L.out = kl.layer.out_features; L.relu = (std::uint32_t)kl.layer.relu;
      L.kernel.assign(kl.layer.kernel_f16, kl.layer.kernel_f16 + (size_t)L.in * L.out);
      if (kl.layer.bias_f16) L.bias.assign(kl.layer.bias_f16, kl.layer.bias_f16 + L.out);
    }
    return w;
  }
};

class B200Scene {
 public:
  using RayCallbackFn = std::function<void(std::size_t, const std::vector<TraceResult>&)>;

  B200Scene(const std::vector<SphereData>& spheres, const std::vector<DiscData>& discs, SceneRef& sceneRef,
            std::vector<TraceResult>& results, std::size_t raysPerWorker, RayCallbackFn* fn = nullptr)
      : spheres_(spheres), discs_(discs), data_(sceneRef), rayStream_(results), rayFunc_(fn),
        maxRaysPerWorker_(raysPerWorker ? raysPerWorker : 1) {}

  void setRuntimeConfig(const RuntimeConfig& c) { config_ = c; }
  const RuntimeConfig& getRuntimeConfig() const { return config_; }

  // IpuScene::loadNifModel (src/IpuScene.cpp:174-187): metadata + weights from the assets.extra directory.
  bool loadNifModel(const std::string& assetPath) {
    try {
      b200rt_nif_metadata md{};
      if (b200rt_read_nif_metadata((assetPath + "/nif_metadata.txt").c_str(), &md) != 0)
        throw std::runtime_error(b200rt_scene_last_error());
      nif_ = NifWeightsFile::load(assetPath + "/converted.hdf5");
      nif_.max = md.max;
      std::copy(md.mean, md.mean + 3, nif_.mean);
      nif_.logToneMap = (std::uint32_t)md.log_tone_map;
      nif_.embedding = md.embedding_dimension;
      haveNif_ = true;
      std::fprintf(stderr, "[info] Loaded NIF model from '%s'\n", assetPath.c_str());
      return true;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "[error] Could not load NIF model from '%s'. Exception: %s\n", assetPath.c_str(), e.what());
    }
    return false;
  }
  void setNifWeights(const NifWeightsFile& w) { nif_ = w; haveNif_ = true; }
  void setHdriRotation(float degrees) { hdriRotationDegrees_ = degrees; }
  void setAvailableMemoryProportion(float) {}  // IPU matmul planner knob; accepted and ignored
  void setMaxNifBatchSize(std::size_t n) { nifMaxRaysPerBatch_ = n; }

  double getTraceTimeSecs() const { return traceTimeSecs_; }
  RayCallbackFn* getRayCallback() { return rayFunc_; }
  std::vector<std::vector<TraceResult>>& getRayBatches() { return rayBatches_; }
  std::size_t getRayStreamSize() const { return raysPerBatch() * sizeof(TraceResult); }
  const b200rt_trace_stats& getStats() const { return stats_; }

  // ipu_utils::GraphManager::run (include/ipu_utils.hpp:531-596): everything in one call, exceptions -> EXIT_FAILURE.
  int run() {
    try {
      execute();
      return EXIT_SUCCESS;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "[error] Exception: %s\n", e.what());
      return EXIT_FAILURE;
    }
  }

 private:
  std::size_t raysPerBatch() const { return maxRaysPerWorker_ * 6 * 1440; }  // workers x compute tiles of one Mk2 IPU

  b200rt_scene_desc makeDesc(int device) const {
    const HostScene& h = *data_.data;
    b200rt_scene_desc d{};
    d.geometry = h.geometry.data(); d.num_geometry = (std::uint32_t)h.geometry.size();
    d.mesh_info = h.meshInfo.data(); d.num_meshes = (std::uint32_t)h.meshInfo.size();
    d.mesh_tris = h.meshTris.data(); d.num_tris = (std::uint32_t)h.meshTris.size();
    d.mesh_verts = h.meshVerts.data(); d.num_verts = (std::uint32_t)h.meshVerts.size();
    d.mesh_normals = h.meshNormals.data(); d.num_normals = (std::uint32_t)h.meshNormals.size();
    d.mat_ids = h.matIDs.data(); d.num_mat_ids = (std::uint32_t)h.matIDs.size();
    d.materials = h.materials.data(); d.num_materials = (std::uint32_t)h.materials.size();
    d.bvh_nodes = h.bvhNodes.data(); d.num_bvh_nodes = (std::uint32_t)h.bvhNodes.size();
    d.max_leaf_depth = h.bvhMaxDepth;
    d.spheres = (const float*)spheres_.data(); d.num_spheres = (std::uint32_t)spheres_.size();
    d.discs = (const float*)discs_.data(); d.num_discs = (std::uint32_t)discs_.size();
    d.image_width = data_.imageWidth; d.image_height = data_.imageHeight;
    d.fov_radians = data_.fovRadians; d.anti_alias_scale = data_.antiAliasScale;
    d.max_path_length = data_.maxPathLength; d.roulette_start_depth = data_.rouletteStartDepth;
    d.samples_per_pixel = data_.samplesPerPixel; d.rng_seed = data_.rngSeed;
    d.path_trace = data_.pathTrace ? 1 : 0;
    d.device = device;
    return d;
  }

  // batchIndex is the batch's index in the whole stream = k * numReplicas + replica, the numbering of
  // RayCallback::fetch (src/RayCallback.cpp:8-24). Runs on a CUDA-owned host thread of that replica; replicas may call
  // concurrently (as the reference's per-replica stream callbacks do), each on its own rayBatches_ entry.
  static void trampoline(std::size_t batchIndex, const void* rays, std::size_t n, void* user) {
    auto* self = (B200Scene*)user;
    auto& batch = self->rayBatches_[batchIndex];
    batch.resize(n);
    std::memcpy(batch.data(), rays, n * sizeof(TraceResult));
    (*self->rayFunc_)(batchIndex, batch);
  }

  void execute() {
    const int available = b200rt_device_count();
    if (available < 1) throw std::runtime_error("no B200 device available (there is no CPU fallback for this path)");
    const std::size_t R = std::max<std::uint32_t>(1, std::min<std::uint32_t>(config_.numGpus, (std::uint32_t)available));
    const std::size_t per = raysPerBatch();
    const std::size_t numBatches = (rayStream_.size() + per - 1) / per;
    // createRayBatches (src/IpuScene.cpp:110-172) without the dud-ray padding the IPU graph needed. The per-batch
    // vectors are only materialised for the callback, which is handed one (RayCallback::fetch).
    rayBatches_.assign(numBatches, {});
    if (rayFunc_)
      for (std::size_t b = 0; b < numBatches; ++b) rayBatches_[b].resize(std::min(rayStream_.size(), (b + 1) * per) - b * per);
    // Every replica renders ITS batches (i % R == replica, src/IpuScene.cpp:676-684) of the caller's stream in place:
    // b200rt_trace takes the stride, so the stream is neither regrouped nor copied on the host.
    // Phase 1 (untimed, like the reference's compile/load/prepareEngine): device scenes, NIF weights, page-locking.
    std::vector<std::string> errors(R);
    std::vector<b200rt_scene*> scenes(R, nullptr);
    auto forEachReplica = [&](auto&& body) {
      std::vector<std::thread> threads;
      for (std::size_t r = 0; r < R; ++r) threads.emplace_back([&, r] { body(r); });
      for (auto& t : threads) t.join();
    };
    forEachReplica([&](std::size_t r) {
      const b200rt_scene_desc d = makeDesc(b200rt_device_ordinal((int)r));  // r-th usable (sm_100) device, not raw ordinal r
      if (b200rt_scene_create(&d, &scenes[r]) != 0) { errors[r] = b200rt_last_error(); return; }
      if (haveNif_ && data_.pathTrace) {
        std::vector<b200rt_nif_layer> layers(nif_.layers.size());
        for (std::size_t i = 0; i < layers.size(); ++i) {
          layers[i].in_features = nif_.layers[i].in; layers[i].out_features = nif_.layers[i].out;
          layers[i].kernel_f16 = nif_.layers[i].kernel.data();
          layers[i].bias_f16 = nif_.layers[i].bias.empty() ? nullptr : nif_.layers[i].bias.data();
          layers[i].relu = (std::int32_t)nif_.layers[i].relu;
        }
        b200rt_nif_desc nd{};
        nd.embedding_dimension = nif_.embedding; nd.num_layers = (std::uint32_t)layers.size(); nd.layers = layers.data();
        nd.max = nif_.max; std::copy(nif_.mean, nif_.mean + 3, nd.mean); nd.log_tone_map = (std::int32_t)nif_.logToneMap;
        if (b200rt_scene_load_nif(scenes[r], &nd) != 0) { errors[r] = b200rt_last_error(); return; }
        b200rt_scene_set_hdri_rotation(scenes[r], hdriRotationDegrees_);
        b200rt_scene_set_max_nif_batch_size(scenes[r], nifMaxRaysPerBatch_);
      }
    });
    // page-lock the shared stream once (portable: every device DMA's from it)
    const bool pinnedStream = !rayStream_.empty() &&
                              b200rt_host_register(rayStream_.data(), rayStream_.size() * sizeof(TraceResult)) == 0;

    // Phase 2 (timed; the span of IpuScene::getTraceTimeSecs, src/IpuScene.cpp:672-696): stream in, trace, stream out.
    std::vector<b200rt_trace_stats> stats(R);
    const auto t0 = std::chrono::steady_clock::now();
    forEachReplica([&](std::size_t r) {
      if (!errors[r].empty() || rayStream_.empty() || r >= numBatches) return;
      b200rt_trace_params p{};
      p.rays_per_batch = (std::uint32_t)per;
      p.batch_stride = (std::uint32_t)R;
      p.first_batch = (std::uint32_t)r;
      if (b200rt_trace(scenes[r], &p, rayStream_.data(), rayStream_.size(), rayFunc_ ? &B200Scene::trampoline : nullptr, this) != 0)
        errors[r] = b200rt_last_error();
      b200rt_get_trace_stats(scenes[r], &stats[r]);
    });
    traceTimeSecs_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    if (pinnedStream) b200rt_host_unregister(rayStream_.data());
    for (std::size_t r = 0; r < R; ++r)
      if (scenes[r]) b200rt_scene_destroy(scenes[r]);
    for (auto& e : errors)
      if (!e.empty()) throw std::runtime_error(e);

    stats_ = b200rt_trace_stats{};
    for (auto& s : stats) {
      stats_.closest_hit_queries += s.closest_hit_queries; stats_.occlusion_queries += s.occlusion_queries;
      stats_.samples += s.samples; stats_.escaped_samples += s.escaped_samples;
      stats_.kernel_launches += s.kernel_launches;
      stats_.kernel_ms = std::max(stats_.kernel_ms, s.kernel_ms);
      stats_.h2d_ms = std::max(stats_.h2d_ms, s.h2d_ms); stats_.d2h_ms = std::max(stats_.d2h_ms, s.d2h_ms);
    }
  }

  const std::vector<SphereData>& spheres_;
  const std::vector<DiscData>& discs_;
  SceneRef data_;
  std::vector<TraceResult>& rayStream_;
  RayCallbackFn* rayFunc_;
  std::size_t maxRaysPerWorker_;
  RuntimeConfig config_;
  NifWeightsFile nif_;
  bool haveNif_ = false;
  float hdriRotationDegrees_ = 0.f;
  std::size_t nifMaxRaysPerBatch_ = 0;
  double traceTimeSecs_ = 0.0;
  b200rt_trace_stats stats_{};
  std::vector<std::vector<TraceResult>> rayBatches_;
};

}  // namespace b200rt
