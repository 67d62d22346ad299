// b200rt.cu — the C ABI of include/b200rt.h over the sm_100a kernels in trace_kernels.cuh.
//
// One b200rt_scene == one replica of the reference's IpuScene (src/IpuScene.cpp): the scene arrays
// are uploaded once (the reference broadcasts them to every tile, :477-483), the TraceResult stream
// is copied to HBM, rendered in place and copied back. There is no CPU fallback anywhere in this
// file: without an sm_100 device every compute entry point returns B200RT_ERR_CUDA.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "nif.cuh"
#include "scene_tables.hpp"
#include "trace_kernels.cuh"
#include "wavefront.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(e__ == cudaErrorMemoryAllocation ? B200RT_ERR_OOM : B200RT_ERR_CUDA,            \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                           \
  } while (0)

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t upload(const void* src, size_t n, size_t padTo = 16) {
    bytes = ((n + padTo - 1) / padTo) * padTo;
    if (bytes == 0) bytes = padTo;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return e;
    if (n && src) e = cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
    // A pageable-source cudaMemcpy may return once the data is staged, before the DMA has landed; the kernels that
    // read this buffer run on the scene's own non-blocking stream, which is not ordered after the default stream.
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    return e;
  }
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

// Streams and events of the host-streaming pipeline (b200rt_trace), created on first use and kept with the scene: a
// single-pass render of a 1440^2 stream lasts a few milliseconds end to end, stream / event creation must not be in it.
struct StreamPipe {
  static constexpr int kRing = 3;
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t evIn[kRing]{}, evRender[kRing]{}, evOut[kRing]{};
  std::vector<cudaEvent_t> timing;  // pool; per tile: h2d begin, h2d end, d2h begin, d2h end
  size_t timingUsed = 0;
  cudaError_t init() {
    if (in) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking);
    for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&evIn[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evRender[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evOut[i], cudaEventDisableTiming);
    }
    return e;
  }
  cudaError_t record_timing(cudaStream_t st) {
    if (timingUsed == timing.size()) {
      cudaEvent_t e;
      const cudaError_t rc = cudaEventCreate(&e);
      if (rc != cudaSuccess) return rc;
      timing.push_back(e);
    }
    return cudaEventRecord(timing[timingUsed++], st);
  }
  void release() {
    for (int i = 0; i < kRing; ++i)
      for (cudaEvent_t e : {evIn[i], evRender[i], evOut[i]})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : timing) cudaEventDestroy(e);
    timing.clear();
    if (in) cudaStreamDestroy(in);
    if (out) cudaStreamDestroy(out);
    in = out = nullptr;
  }
};

struct b200rt_scene {
  int device = 0;
  int numSMs = 0;
  int maxSmemOptin = 0;
  bool cooperativeLaunch = false;  // wf_tail_kernel needs grid-wide barriers
  cudaStream_t stream = nullptr;
  cudaEvent_t evStart = nullptr, evStop = nullptr;
  b200rt_scene_desc desc{};  // scalars only are used after creation
  rt::DevScene dev{};
  uint32_t nodeBytes = 0, pairBytes = 0;
  DeviceBuffer nodes, pairs, leafOrig, leafInfo, geoms, triVerts, triNormals, triFaceNormals, spheres, discs, matIDs, materials;
  DeviceBuffer workCounter, counters, primA, primB;
  DeviceBuffer rays;                                   // device copy of the stream (host-buffer entry point)
  StreamPipe pipe;
  std::vector<cudaEvent_t> timerEvents;                // pool of the per-kernel timing events (KernelTimer), kept across calls
  DeviceBuffer slotColor, slotEscape, slotEnv, escapeQueue, escapeCount;  // NIF wavefront
  // Chunk overlap (render_tile): the NIF + accumulate of chunk c run on nifStream while sc.stream traces chunk c + 1
  // into the second set of per-sample records.
  DeviceBuffer slotColorB, slotEscapeB, escapeQueueB, escapeCountB;
  cudaStream_t nifStream = nullptr;
  cudaEvent_t evTraceDone[2]{}, evNifDone[2]{};
  DeviceBuffer wfState[2][7], wfHitA, wfHitB, wfCounts;  // wavefront path state (two slot-indexed arrays of records)
  b200rt_trace_stats stats{};
  float hdriRotationDegrees = 0.f;
  float rootBox[6] = {0.f, 0.f, 0.f, 1.f, 1.f, 1.f};  // min, extent of the root node
  size_t maxNifBatch = 0;
  rt::NifModel* nif = nullptr;

  ~b200rt_scene() {
    cudaSetDevice(device);
    if (nif) rt::nif_destroy(nif);
    for (DeviceBuffer* b : {&nodes, &pairs, &leafOrig, &leafInfo, &geoms, &triVerts, &triNormals, &triFaceNormals, &spheres, &discs, &matIDs, &materials,
                            &workCounter, &counters, &primA, &primB, &rays, &slotColor, &slotEscape, &slotEnv, &escapeQueue,
                            &escapeCount, &wfHitA, &wfHitB, &wfCounts, &slotColorB, &slotEscapeB, &escapeQueueB, &escapeCountB})
      b->release();
    for (cudaEvent_t e : {evTraceDone[0], evTraceDone[1], evNifDone[0], evNifDone[1]})
      if (e) cudaEventDestroy(e);
    if (nifStream) cudaStreamDestroy(nifStream);
    for (auto& set : wfState)
      for (DeviceBuffer& b : set) b.release();
    pipe.release();
    for (cudaEvent_t e : timerEvents) cudaEventDestroy(e);
    if (evStart) cudaEventDestroy(evStart);
    if (evStop) cudaEventDestroy(evStop);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

int validate_desc(const b200rt_scene_desc& d) {
  if (d.num_bvh_nodes == 0 || !d.bvh_nodes) return fail(B200RT_ERR_INVALID_ARG, "scene has no BVH nodes");
  if (d.num_geometry == 0 || !d.geometry) return fail(B200RT_ERR_INVALID_ARG, "scene has no geometry");
  if (d.num_mat_ids < d.num_geometry || !d.mat_ids)
    return fail(B200RT_ERR_INVALID_ARG, "All primitives must be assigned a material.");
  if (d.num_materials == 0 || !d.materials) return fail(B200RT_ERR_INVALID_ARG, "scene has no materials");
  if (d.num_normals != 0 && d.num_normals != d.num_verts)
    return fail(B200RT_ERR_INVALID_ARG, "mesh_normals must be empty or one per vertex");
  if (d.path_trace && d.max_path_length == 0)
    return fail(B200RT_ERR_INVALID_ARG, "max_path_length must be at least 1 for path tracing");
  const auto* mat = (const uint32_t*)d.mat_ids;
  for (uint32_t i = 0; i < d.num_geometry; ++i)
    if (mat[i] >= d.num_materials) return fail(B200RT_ERR_INVALID_ARG, "material index out of range");
  return B200RT_OK;
}

template <bool kShared, bool kOrdered, bool kCount>
cudaError_t launch_shadow(const rt::TraceArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
  auto k = rt::shadow_trace_kernel<kShared, kOrdered, kCount>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}

template <bool kShared, bool kOrdered, bool kCount, bool kNif>
cudaError_t launch_path(const rt::TraceArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
  auto k = rt::path_trace_kernel<kShared, kOrdered, kCount, kNif>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}

template <bool kShared, bool kOrdered>
cudaError_t dispatch_shadow(bool count, const rt::TraceArgs& a, int g, int b, size_t s, cudaStream_t st) {
  return count ? launch_shadow<kShared, kOrdered, true>(a, g, b, s, st)
               : launch_shadow<kShared, kOrdered, false>(a, g, b, s, st);
}
template <bool kShared, bool kOrdered>
cudaError_t dispatch_path(bool count, bool nif, const rt::TraceArgs& a, int g, int b, size_t s, cudaStream_t st) {
  if (count) return nif ? launch_path<kShared, kOrdered, true, true>(a, g, b, s, st)
                        : launch_path<kShared, kOrdered, true, false>(a, g, b, s, st);
  return nif ? launch_path<kShared, kOrdered, false, true>(a, g, b, s, st)
             : launch_path<kShared, kOrdered, false, false>(a, g, b, s, st);
}

template <bool kShared, bool kOrdered, bool kCount>
cudaError_t launch_primary(const rt::TraceArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
  auto k = rt::primary_hit_kernel<kShared, kOrdered, kCount>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}
template <bool kShared, bool kOrdered>
cudaError_t dispatch_primary(bool count, const rt::TraceArgs& a, int g, int b, size_t s, cudaStream_t st) {
  return count ? launch_primary<kShared, kOrdered, true>(a, g, b, s, st)
               : launch_primary<kShared, kOrdered, false>(a, g, b, s, st);
}

struct LaunchPlan {
  bool shared, ordered, count;
  int grid, block;
  size_t smem;
};

LaunchPlan plan_launch(const b200rt_scene& sc, const b200rt_trace_params& p, bool wavefront) {
  LaunchPlan L;
  L.ordered = p.traversal != 1;  // auto = near-first
  // what the kernel keeps in shared memory: the pair table (streaming kernels: shadow trace near-first, wf_trace) or
  // the caller's own node array (reference-order walk, megakernel)
  const bool pairTable = L.ordered && (!sc.desc.path_trace || wavefront);
  const size_t need = (((size_t)(pairTable ? sc.pairBytes : sc.nodeBytes) + 15) / 16 * 16) + 16;
  const bool fits = need + 1024 <= (size_t)sc.maxSmemOptin;
  L.shared = p.scene_residency == 1 ? fits : (p.scene_residency == 2 ? false : fits);
  L.count = p.count_visits != 0;
  static const int envThreads = [] { const char* e = std::getenv("B200RT_THREADS_PER_SM"); return e ? std::atoi(e) : 0; }();
  // 768 threads/SM (24 warps at ~80 registers): measured 20 % faster than 512 on the path tracer
  // the megakernels and the shadow kernel are compiled with __launch_bounds__(768)
  const int threadsPerSM = envThreads > 0 ? std::min(std::max(envThreads, 128) / 128 * 128, 768) : 768;
  if (L.shared) {
    L.block = threadsPerSM;  // one CTA per SM owns the staged BVH
    L.grid = sc.numSMs;
    L.smem = need;
  } else {
    L.block = 128;
    L.grid = sc.numSMs * (threadsPerSM / 128);
    L.smem = 0;
  }
  return L;
}

cudaError_t run_shadow(b200rt_scene& sc, const LaunchPlan& L, const rt::TraceArgs& a) {
  if (L.shared) return L.ordered ? dispatch_shadow<true, true>(L.count, a, L.grid, L.block, L.smem, sc.stream)
                                 : dispatch_shadow<true, false>(L.count, a, L.grid, L.block, L.smem, sc.stream);
  return L.ordered ? dispatch_shadow<false, true>(L.count, a, L.grid, L.block, L.smem, sc.stream)
                   : dispatch_shadow<false, false>(L.count, a, L.grid, L.block, L.smem, sc.stream);
}
cudaError_t run_primary(b200rt_scene& sc, const LaunchPlan& L, const rt::TraceArgs& a) {
  if (L.shared) return L.ordered ? dispatch_primary<true, true>(L.count, a, L.grid, L.block, L.smem, sc.stream)
                                 : dispatch_primary<true, false>(L.count, a, L.grid, L.block, L.smem, sc.stream);
  return L.ordered ? dispatch_primary<false, true>(L.count, a, L.grid, L.block, L.smem, sc.stream)
                   : dispatch_primary<false, false>(L.count, a, L.grid, L.block, L.smem, sc.stream);
}
cudaError_t run_path(b200rt_scene& sc, const LaunchPlan& L, bool nif, const rt::TraceArgs& a) {
  if (L.shared) return L.ordered ? dispatch_path<true, true>(L.count, nif, a, L.grid, L.block, L.smem, sc.stream)
                                 : dispatch_path<true, false>(L.count, nif, a, L.grid, L.block, L.smem, sc.stream);
  return L.ordered ? dispatch_path<false, true>(L.count, nif, a, L.grid, L.block, L.smem, sc.stream)
                   : dispatch_path<false, false>(L.count, nif, a, L.grid, L.block, L.smem, sc.stream);
}

// Default of b200rt_trace_params::chunk_overlap = 0 (see render_tile); build knob for A/B runs.
#ifndef B200RT_CHUNK_OVERLAP_DEFAULT
#define B200RT_CHUNK_OVERLAP_DEFAULT 1
#endif
constexpr bool kChunkOverlapDefault = B200RT_CHUNK_OVERLAP_DEFAULT != 0;

// Per-kernel device timing with CUDA event pairs recorded on the launching stream.
struct KernelTimer {
  enum Kind { TRACE = 0, NIF = 1, ACCUM = 2, SHADE = 3 };
  struct Span { cudaEvent_t a, b; Kind kind; int launches; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t>& pool;  // the scene's: creating events is host time inside short renders
  size_t used = 0;
  explicit KernelTimer(std::vector<cudaEvent_t>& p) : pool(p) {}
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
  void begin(Kind k, cudaStream_t st) { Span s{get(), get(), k, 1}; cudaEventRecord(s.a, st); spans.push_back(s); }
  void end(cudaStream_t st, int launches = 1) { spans.back().launches = launches; cudaEventRecord(spans.back().b, st); }
  void collect(b200rt_trace_stats& out) {
    for (const Span& s : spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) continue;
      if (s.kind == TRACE) { out.trace_kernel_ms += ms; out.trace_kernel_launches += 1; }
      else if (s.kind == NIF) { out.nif_kernel_ms += ms; out.nif_kernel_launches += (uint64_t)s.launches; }
      else if (s.kind == SHADE) { out.shade_kernel_ms += ms; out.shade_kernel_launches += 1; }
      else out.accumulate_kernel_ms += ms;
    }
    spans.clear();
    used = 0;
  }
};

float host_tan_half_fov(float fov) {
  float s, c;
  rt::sincos_tbl(fov / 2.f, s, c);
  return s / c;
}

// One b200rt_trace / b200rt_trace_device call: kernels are enqueued tile by tile on sc.stream without any host
// synchronisation; the statistics (per-kernel CUDA-event spans, device-side counters) are collected once at the end.
struct RenderRun {
  KernelTimer timer;
  uint64_t launches = 0;
  explicit RenderRun(b200rt_scene& sc) : timer(sc.timerEvents) {}
};

int render_begin(b200rt_scene& sc, RenderRun&) {
  CU_TRY(cudaMemsetAsync(sc.counters.p, 0, sizeof(rt::DeviceCounters), sc.stream));
  CU_TRY(cudaEventRecord(sc.evStart, sc.stream));
  return B200RT_OK;
}

// Enqueues the render of d_rays[0..n) (in place) on sc.stream.
int render_tile(b200rt_scene& sc, const b200rt_trace_params& p, float* d_rays, size_t n, RenderRun& run) {
  if (n == 0) return B200RT_OK;
  if (n > 0xFFFFFFFFull) return fail(B200RT_ERR_INVALID_ARG, "ray stream longer than 2^32 rays");
  KernelTimer& timer = run.timer;
  uint64_t& launches = run.launches;

  // auto = wavefront (measured 30 % faster than the megakernel); its packed record holds the bounce in 8 bits.
  // traversal 3 (the state-machine megakernel of round 1, measured slower) is gone: it selects the wavefront tracer too.
  const bool wavefront = sc.desc.path_trace && (p.traversal == 4 || p.traversal == 3 ||
                                               (p.traversal == 0 && sc.desc.max_path_length <= 255));
  LaunchPlan L = plan_launch(sc, p, wavefront);
  rt::TraceArgs a{};
  a.scene = sc.dev;
  a.rays = d_rays;
  a.numRays = (uint32_t)n;
  a.workCounter = (uint32_t*)sc.workCounter.p;
  a.counters = (rt::DeviceCounters*)sc.counters.p;
  a.nodeBytes = sc.nodeBytes;

  if (!sc.desc.path_trace) {
    const bool dflt = p.light_pos[0] == 0.f && p.light_pos[1] == 0.f && p.light_pos[2] == 0.f && p.ambient == 0.f;
    a.lightX = dflt ? 18.f : p.light_pos[0];
    a.lightY = dflt ? 257.f : p.light_pos[1];
    a.lightZ = dflt ? -1060.f : p.light_pos[2];
    a.ambient = dflt ? .05f : p.ambient;
    CU_TRY(cudaMemsetAsync(sc.workCounter.p, 0, 4, sc.stream));
    // near-first renders of a 16-byte aligned stream: tiles moved by the TMA engine (shadow_stream_kernel)
    static const bool envTma = [] { const char* e = std::getenv("B200RT_SHADOW_TMA"); return !e || e[0] != '0'; }();
    const bool tma = envTma && L.ordered && (reinterpret_cast<uintptr_t>(d_rays) % 16u) == 0u;
    timer.begin(KernelTimer::TRACE, sc.stream);
    if (tma) {
      const size_t ring = rt::kStreamRingBytes;
      const bool shared = p.scene_residency != 2 && (size_t)sc.pairBytes + ring <= (size_t)sc.maxSmemOptin;
      const size_t smem = (shared ? (size_t)sc.pairBytes : 0) + ring;
      auto k = shared ? (L.count ? rt::shadow_stream_kernel<true, true> : rt::shadow_stream_kernel<true, false>)
                      : (L.count ? rt::shadow_stream_kernel<false, true> : rt::shadow_stream_kernel<false, false>);
      CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<sc.numSMs, rt::kStreamThreads, smem, sc.stream>>>(a);
      CU_TRY(cudaGetLastError());
    } else {
      CU_TRY(run_shadow(sc, L, a));
    }
    timer.end(sc.stream);
    launches += 1;
  } else {
    if (!(sc.desc.image_width >= 1.f) || !(sc.desc.image_height >= 1.f))
      return fail(B200RT_ERR_INVALID_ARG, "image_width/image_height must be set for path tracing");
    a.imageWidth = sc.desc.image_width;
    a.imageHeight = sc.desc.image_height;
    a.tanTheta = host_tan_half_fov(sc.desc.fov_radians);
    a.antiAlias = sc.desc.anti_alias_scale;
    a.maxPathLength = sc.desc.max_path_length;
    a.rouletteStartDepth = sc.desc.roulette_start_depth;
    a.rngKey = rt::splitmix64(sc.desc.rng_seed);
    const uint32_t first = p.first_sample;
    const uint32_t count = p.num_samples ? p.num_samples : sc.desc.samples_per_pixel;
    static const int envPrimary = [] { const char* e = std::getenv("B200RT_PRIMARY_PASS"); return e ? std::atoi(e) : 0; }();
    const uint32_t primarySel = p.primary_pass ? p.primary_pass : (uint32_t)envPrimary;  // 0 = auto (off: no measured gain)
    const bool primaryPass = !wavefront && primarySel == 1;
    if (!sc.nif && !wavefront && !primaryPass) {
      a.firstSample = first;
      a.endSample = first + count;
      CU_TRY(cudaMemsetAsync(sc.workCounter.p, 0, 4, sc.stream));
      timer.begin(KernelTimer::TRACE, sc.stream);
      CU_TRY(run_path(sc, L, false, a));
      timer.end(sc.stream);
      launches += 1;
    } else {
      // Chunks of samples: trace -> (compacted escaped rays) NIF -> ordered accumulate.
      // Default chunk: enough samples for ~64 M paths per chunk (32 samples of a 1440 x 1440 frame), so that a GPU that
      // owns a small shard of the stream (multi-GPU, crops) launches as few, as large kernels as one that owns it all.
      uint32_t chunk = p.samples_per_chunk;
      if (!chunk) {
        chunk = 32u;
        while (chunk < 1024u && (size_t)chunk * 2 * n <= ((size_t)64 << 20)) chunk *= 2;
      }
      // bound the per-path arrays (slots; plus 2 x 112 B of path state + 24 B of hit in wavefront mode) to ~8 / ~24 GiB
      const bool slots = sc.nif || wavefront;  // per-sample colour / escape records handed to NIF + accumulate
      const bool primB = primaryPass && sc.dev.triNormals != nullptr;
      const size_t perSlot = (slots ? (3 + 5 + 3 + 1) * sizeof(float) : 0) + (wavefront ? 248 : 0) +
                             (primaryPass ? (primB ? 32 : 16) : 0);
      static const size_t envBudget = [] { const char* e = std::getenv("B200RT_CHUNK_GIB"); return e ? (size_t)std::atoi(e) : (size_t)0; }();
      const size_t budget = envBudget ? envBudget << 30 : (wavefront ? (size_t)24 << 30 : (size_t)8 << 30);
      while (chunk > 1 && (size_t)chunk * n * perSlot > budget) chunk /= 2;
      if (chunk > count) chunk = count ? count : 1;
      // Chunk overlap (below) needs at least two chunks: a short render that would fit one chunk is cut into up to four,
      // as long as a chunk keeps >= 4 M paths (kernels of that size still run at full rate: 8-spp chunks of the 1440^2
      // frame cost 2.5 % over 32-spp ones, profiles/r02_overlap_ab.txt).
      static const int envOverlap = [] { const char* e = std::getenv("B200RT_OVERLAP"); return e ? std::atoi(e) : -1; }();
      const bool overlapWanted = (envOverlap >= 0 ? envOverlap != 0 : (p.chunk_overlap == 0u ? kChunkOverlapDefault : p.chunk_overlap == 2u)) &&
                                 sc.nif && wavefront;
      if (overlapWanted && !p.samples_per_chunk && chunk >= count && count >= 2u) {
        uint32_t parts = 4u;
        while (parts > 1u && (size_t)((count + parts - 1u) / parts) * n < ((size_t)4 << 20)) --parts;
        chunk = (count + parts - 1u) / parts;
      }
      if ((size_t)chunk * n > 0xFFFFFFF0ull) return fail(B200RT_ERR_UNSUPPORTED, "ray stream too long for one chunk");
      const size_t P = (size_t)chunk * n;
      CU_TRY(sc.escapeCount.reserve(16));
      if (slots) {
        CU_TRY(sc.slotColor.reserve(P * 3 * sizeof(float)));
        CU_TRY(sc.slotEscape.reserve(P * 5 * sizeof(float)));
        CU_TRY(sc.slotEnv.reserve(P * 3 * sizeof(float)));
        CU_TRY(sc.escapeQueue.reserve(P * sizeof(uint32_t)));
        a.slotColor = (float*)sc.slotColor.p;
        a.slotEscape = (float*)sc.slotEscape.p;
        a.escapeQueue = (uint32_t*)sc.escapeQueue.p;
        a.escapeCount = (uint32_t*)sc.escapeCount.p;
      }
      if (primaryPass) {
        CU_TRY(sc.primA.reserve(P * 16));
        a.primA = (uint4*)sc.primA.p;
        if (primB) { CU_TRY(sc.primB.reserve(P * 16)); a.primB = (float4*)sc.primB.p; }
      }
      a.hdriRotation = (sc.hdriRotationDegrees / 360.f) * (float)(2.0 * M_PI);  // src/IpuScene.cpp:641
      rt::WfArgs w{};
      if (wavefront) {
        if (sc.desc.max_path_length > 255) return fail(B200RT_ERR_UNSUPPORTED, "wavefront path tracer: max_path_length > 255");
        const bool needBary = sc.dev.triNormals != nullptr;  // barycentrics only feed interpolated normals
        for (int k = 0; k < 2; ++k) {
          for (DeviceBuffer& b : sc.wfState[k]) CU_TRY(b.reserve(P * 16));
          rt::WfState& st = w.b.st[k];
          st.rayO = (float4*)sc.wfState[k][0].p; st.rayD = (float4*)sc.wfState[k][1].p; st.rayI = (float4*)sc.wfState[k][2].p;
          st.rayS = (float4*)sc.wfState[k][3].p; st.thr = (float4*)sc.wfState[k][4].p; st.rng = (uint4*)sc.wfState[k][5].p;
          st.nrm = (float4*)sc.wfState[k][6].p;
        }
        CU_TRY(sc.wfHitA.reserve(P * 8));
        if (needBary) CU_TRY(sc.wfHitB.reserve(P * 16));
        CU_TRY(sc.wfCounts.reserve(16));
        w.b.hitA = (float2*)sc.wfHitA.p; w.b.hitB = needBary ? (float4*)sc.wfHitB.p : nullptr;
        w.b.counts = (uint32_t*)sc.wfCounts.p;
        w.lastSample = first + count - 1;
      }
      // Chunk overlap: the NIF MLP (tensor pipe, weights streamed from L2) and the accumulate of chunk c run on a second
      // stream while sc.stream traces and shades chunk c + 1 (issue- and HBM-bound SIMT work) into the other set of
      // per-sample records; the accumulates stay in chunk order on that one stream, so rgb is bit-identical.
      // params.chunk_overlap: 0 = auto, 1 = off, 2 = on; B200RT_OVERLAP=0/1 overrides it for A/B runs.
      const bool overlap = overlapWanted && count > chunk;
      if (overlap) {
        if (!sc.nifStream) {
          // higher priority than sc.stream: when an SM frees resources the NIF's one CTA per SM is placed before the
          // next trace / shade blocks, which then fill what is left beside it
          static const bool envPrio = [] { const char* e = std::getenv("B200RT_NIF_PRIORITY"); return !e || e[0] != '0'; }();
          int prioLow = 0, prioHigh = 0;
          CU_TRY(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
          CU_TRY(cudaStreamCreateWithPriority(&sc.nifStream, cudaStreamNonBlocking, envPrio ? prioHigh : prioLow));
          for (int k = 0; k < 2; ++k) {
            CU_TRY(cudaEventCreateWithFlags(&sc.evTraceDone[k], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&sc.evNifDone[k], cudaEventDisableTiming));
          }
        }
        CU_TRY(sc.slotColorB.reserve(P * 3 * sizeof(float)));
        CU_TRY(sc.slotEscapeB.reserve(P * 5 * sizeof(float)));
        CU_TRY(sc.escapeQueueB.reserve(P * sizeof(uint32_t)));
        CU_TRY(sc.escapeCountB.reserve(16));
      }
      if (overlap && L.shared && p.scene_residency == 0u) {
        // The NIF CTA fills its SM's shared memory, so a wf_trace CTA that stages the pair table could only alternate with
        // it. The L2-resident form (128-thread CTAs, table read through L1; 2 % slower alone) runs beside it.
        b200rt_trace_params pl2 = p;
        pl2.scene_residency = 2u;
        L = plan_launch(sc, pl2, wavefront);
      }
      cudaStream_t const nifStream = overlap ? sc.nifStream : sc.stream;
      uint32_t chunkIndex = 0;
      for (uint32_t s0 = first; s0 < first + count; s0 += chunk, ++chunkIndex) {
        const uint32_t c = std::min(chunk, first + count - s0);
        a.firstSample = s0;
        a.endSample = s0 + c;
        const int set = overlap ? (int)(chunkIndex & 1u) : 0;
        if (set) {
          a.slotColor = (float*)sc.slotColorB.p; a.slotEscape = (float*)sc.slotEscapeB.p;
          a.escapeQueue = (uint32_t*)sc.escapeQueueB.p; a.escapeCount = (uint32_t*)sc.escapeCountB.p;
        } else if (slots) {
          a.slotColor = (float*)sc.slotColor.p; a.slotEscape = (float*)sc.slotEscape.p;
          a.escapeQueue = (uint32_t*)sc.escapeQueue.p; a.escapeCount = (uint32_t*)sc.escapeCount.p;
        }
        // this set's records are free again once the NIF + accumulate of two chunks ago are done
        if (overlap && chunkIndex >= 2u) CU_TRY(cudaStreamWaitEvent(sc.stream, sc.evNifDone[set], 0));
        CU_TRY(cudaMemsetAsync(set ? sc.escapeCountB.p : sc.escapeCount.p, 0, 4, sc.stream));
        if (!wavefront) {
          CU_TRY(cudaMemsetAsync(sc.workCounter.p, 0, 8, sc.stream));
          timer.begin(KernelTimer::TRACE, sc.stream);  // one span: pre-pass + path tracer = the trace step of a chunk
          if (primaryPass) { CU_TRY(run_primary(sc, L, a)); launches += 1; }
          CU_TRY(run_path(sc, L, sc.nif != nullptr, a));
          timer.end(sc.stream);
          launches += 1;
        } else {
          w.t = a;
          w.chunk = c;
          w.chunkShift = (c & (c - 1)) == 0 ? __builtin_ctz(c) : -1;
          w.numPaths = (uint32_t)((size_t)c * n);
          const int gridSmall = sc.numSMs * rt::kShadeBlocksPerSM;
          static const int envWfThreads = [] { const char* e = std::getenv("B200RT_WF_THREADS"); return e ? std::atoi(e) : 0; }();
          // wf_trace needs <= 64 registers: 32 warps per SM (measured 5 % faster than 24)
          const int wfBlock = envWfThreads > 0 ? std::min(envWfThreads, 1024) : (L.shared ? 1024 : L.block);
          const int wfGrid = L.shared ? L.grid : sc.numSMs * (1024 / L.block);
          CU_TRY(cudaMemsetAsync(sc.wfCounts.p, 0, 16, sc.stream));
          static const bool envPhase = std::getenv("B200RT_WF_PHASE_STATS") != nullptr;
          unsigned long long* dPhase = nullptr;
          if (envPhase && L.count) { cudaMalloc(&dPhase, 16 * 6 * 8); cudaMemsetAsync(dPhase, 0, 16 * 6 * 8, sc.stream); }
          // Bounces from tailStart on run in one cooperative launch (wf_tail_kernel): with Russian roulette the paths alive
          // two bounces after it starts are a fraction of a per cent of the chunk. B200RT_WF_TAIL overrides params.tail_bounce (1 = off, N = from bounce N) for A/B runs.
          static const int envTail = [] { const char* e = std::getenv("B200RT_WF_TAIL"); return e ? std::atoi(e) : -1; }();
          const uint32_t tailWanted = envTail >= 0 ? (uint32_t)envTail : p.tail_bounce;  // 0 = auto, 1 = never, N = from bounce N
          uint32_t tailStart = tailWanted == 0u ? sc.desc.roulette_start_depth + 2u : tailWanted;
          if (tailStart < 2u || tailStart + 2u > sc.desc.max_path_length || dPhase || !sc.cooperativeLaunch) tailStart = 0xFFFFFFFFu;
          for (uint32_t b = 0; b < sc.desc.max_path_length; ++b) {
            if (b == tailStart) {
              w.qIn = (int)(b & 1u);
              w.phaseStats = nullptr;
              uint32_t bounceBegin = b, bounceEnd = sc.desc.max_path_length;
              void* kargs[] = {&w, &bounceBegin, &bounceEnd};
              const void* fn;
              const bool nifOn = sc.nif != nullptr;
              if (L.shared) fn = L.count ? (nifOn ? (const void*)rt::wf_tail_kernel<true, true, true> : (const void*)rt::wf_tail_kernel<true, true, false>)
                                         : (nifOn ? (const void*)rt::wf_tail_kernel<true, false, true> : (const void*)rt::wf_tail_kernel<true, false, false>);
              else fn = L.count ? (nifOn ? (const void*)rt::wf_tail_kernel<false, true, true> : (const void*)rt::wf_tail_kernel<false, true, false>)
                                : (nifOn ? (const void*)rt::wf_tail_kernel<false, false, true> : (const void*)rt::wf_tail_kernel<false, false, false>);
              const size_t tailSmem = L.shared ? L.smem : 0;
              if (tailSmem) CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tailSmem));
              timer.begin(KernelTimer::TRACE, sc.stream);
              CU_TRY(cudaLaunchCooperativeKernel(fn, dim3((unsigned)sc.numSMs), dim3(1024), kargs, tailSmem, sc.stream));
              timer.end(sc.stream);
              launches += 1;
              break;
            }
            w.qIn = (int)(b & 1u);
            w.phaseStats = dPhase ? dPhase + 6 * std::min<uint32_t>(b, 15u) : nullptr;
            // no memsets between the kernels: wf_trace empties the counter its wf_shade appends to, wf_shade resets the
            // fetch cursor of the next wf_trace
            const bool first = b == 0;  // bounce 0: camera rays are generated in the kernels, slot = path
            timer.begin(KernelTimer::TRACE, sc.stream);
            {
              cudaError_t e;
              if (L.shared) {
                auto k = L.count ? (first ? rt::wf_trace_kernel<true, true, true> : rt::wf_trace_kernel<true, true, false>)
                                 : (first ? rt::wf_trace_kernel<true, false, true> : rt::wf_trace_kernel<true, false, false>);
                e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
                if (e == cudaSuccess) { k<<<wfGrid, wfBlock, L.smem, sc.stream>>>(w); e = cudaGetLastError(); }
              } else {
                auto k = L.count ? (first ? rt::wf_trace_kernel<false, true, true> : rt::wf_trace_kernel<false, true, false>)
                                 : (first ? rt::wf_trace_kernel<false, false, true> : rt::wf_trace_kernel<false, false, false>);
                k<<<wfGrid, L.block, 0, sc.stream>>>(w);
                e = cudaGetLastError();
              }
              CU_TRY(e);
            }
            timer.end(sc.stream);
            timer.begin(KernelTimer::SHADE, sc.stream);
            if (sc.nif) {
              if (first) rt::wf_shade_kernel<true, true><<<gridSmall, rt::kShadeThreads, 0, sc.stream>>>(w);
              else rt::wf_shade_kernel<true, false><<<gridSmall, rt::kShadeThreads, 0, sc.stream>>>(w);
            } else {
              if (first) rt::wf_shade_kernel<false, true><<<gridSmall, rt::kShadeThreads, 0, sc.stream>>>(w);
              else rt::wf_shade_kernel<false, false><<<gridSmall, rt::kShadeThreads, 0, sc.stream>>>(w);
            }
            timer.end(sc.stream);
            CU_TRY(cudaGetLastError());
            launches += 2;
          }
          if (dPhase) {  // debugging aid (B200RT_WF_PHASE_STATS + count_visits): scheduler statistics per bounce
            unsigned long long h[16 * 6];
            cudaStreamSynchronize(sc.stream);
            cudaMemcpy(h, dPhase, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(dPhase);
            for (int b = 0; b < 6; ++b) {
              const unsigned long long* r = h + 6 * b;
              std::fprintf(stderr, "[wf phase stats] bounce %d: trav %llu iters x %.1f lanes, leaf %llu x %.1f, fetch %llu x %.1f\n", b,
                           r[0], r[0] ? (double)r[1] / r[0] : 0.0, r[2], r[2] ? (double)r[3] / r[2] : 0.0, r[4],
                           r[4] ? (double)r[5] / r[4] : 0.0);
            }
          }
        }
        if (overlap) {
          CU_TRY(cudaEventRecord(sc.evTraceDone[set], sc.stream));
          CU_TRY(cudaStreamWaitEvent(nifStream, sc.evTraceDone[set], 0));
        }
        if (sc.nif) {
          int nifLaunches = 0;
          timer.begin(KernelTimer::NIF, nifStream);
          const int rc = rt::nif_eval_queue(sc.nif, a.slotEscape, a.escapeQueue, a.escapeCount,
                                            (uint32_t)std::min<size_t>((size_t)c * n, 0xFFFFFFFFull),
                                            (uint32_t)std::min<size_t>(sc.maxNifBatch, 0xFFFFFFFFull), (float*)sc.slotEnv.p, nifStream,
                                            &nifLaunches);
          timer.end(nifStream, nifLaunches);
          if (rc != 0) return fail(B200RT_ERR_CUDA, std::string("NIF evaluation failed: ") + rt::nif_last_error());
          launches += (uint64_t)nifLaunches;
        }
        if (!slots) continue;  // path tracer without an environment light: rgb was accumulated in the kernel
        const uint32_t threads = 256, blocks = (uint32_t)((n + rt::kAccPixels - 1) / rt::kAccPixels);
        timer.begin(KernelTimer::ACCUM, nifStream);
        rt::wf_accumulate_kernel<<<blocks, threads, 0, nifStream>>>(d_rays, (uint32_t)n, c, a.slotColor, a.slotEscape,
                                                                    sc.nif ? (const float*)sc.slotEnv.p : nullptr);
        timer.end(nifStream);
        CU_TRY(cudaGetLastError());
        launches += 1;
        if (overlap) CU_TRY(cudaEventRecord(sc.evNifDone[set], nifStream));
      }
      // join: whatever follows on sc.stream (the next tile, the D2H copy, evStop) is ordered after the last accumulate
      if (overlap && chunkIndex > 0u) CU_TRY(cudaStreamWaitEvent(sc.stream, sc.evNifDone[(chunkIndex - 1u) & 1u], 0));
    }
  }
  return B200RT_OK;
}

// Path-traced streams much longer than a frame are rendered in ray tiles, so that the per-path state of a chunk
// (tile rays x samples per chunk) keeps its ~64 M-path size instead of the chunk shrinking to a couple of samples
// (8192^2 rays: 2 samples per chunk = ~10 000 tiny launches per 1000 spp). Per-(pixel, sample) RNG streams make the
// result independent of the tiling.
int render_enqueue(b200rt_scene& sc, const b200rt_trace_params& p, float* d_rays, size_t n, RenderRun& run) {
  constexpr size_t kTile = (size_t)1 << 21;
  if (!sc.desc.path_trace || n <= 2 * kTile) return render_tile(sc, p, d_rays, n, run);
  for (size_t off = 0; off < n; off += kTile)
    if (int rc = render_tile(sc, p, d_rays + off * rt::TR_WORDS, std::min(kTile, n - off), run)) return rc;
  return B200RT_OK;
}

int render_end(b200rt_scene& sc, RenderRun& run) {
  CU_TRY(cudaEventRecord(sc.evStop, sc.stream));
  CU_TRY(cudaStreamSynchronize(sc.stream));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, sc.evStart, sc.evStop));
  run.timer.collect(sc.stats);
  const uint64_t launches = run.launches;
  rt::DeviceCounters hc{};
  CU_TRY(cudaMemcpy(&hc, sc.counters.p, sizeof(hc), cudaMemcpyDeviceToHost));
  sc.stats.closest_hit_queries += hc.closest;
  sc.stats.occlusion_queries += hc.occlusion;
  sc.stats.node_visits += hc.nodeVisits;
  sc.stats.prim_tests += hc.primTests;
  sc.stats.samples += hc.samples;
  sc.stats.escaped_samples += hc.escaped;
  sc.stats.kernel_ms += ms;
  sc.stats.kernel_launches += launches;
  return B200RT_OK;
}

// Renders d_rays[0..n) in place on `stream` (nullptr = the scene's own stream) and waits for it.
int render_device(b200rt_scene& sc, const b200rt_trace_params& p, float* d_rays, size_t n, cudaStream_t stream) {
  if (n == 0) return B200RT_OK;
  cudaStream_t saved = sc.stream;
  if (stream) sc.stream = stream;
  struct Restore { b200rt_scene& s; cudaStream_t v; ~Restore() { s.stream = v; } } restore{sc, saved};
  RenderRun run(sc);
  if (int rc = render_begin(sc, run)) return rc;
  if (int rc = render_enqueue(sc, p, d_rays, n, run)) {
    // nothing of a failed render may still be running when the caller frees or reuses its buffers
    cudaStreamSynchronize(sc.stream);
    if (sc.nifStream) cudaStreamSynchronize(sc.nifStream);
    return rc;
  }
  return render_end(sc, run);
}

}  // namespace

extern "C" {

int b200rt_abi_version(void) { return B200RT_ABI_VERSION; }
const char* b200rt_last_error(void) { return g_error.c_str(); }

int b200rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int usable = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) usable++;
  }
  return usable;
}

int b200rt_device_ordinal(int index) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (int i = 0, usable = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10 && usable++ == index) return i;
  }
  return -1;
}

int b200rt_scene_create(const b200rt_scene_desc* d, b200rt_scene** out) {
  if (!d || !out) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (int rc = validate_desc(*d)) return rc;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(B200RT_ERR_CUDA, "no CUDA device: the trace path has no CPU fallback");
  }
  int device = d->device;
  if (device < 0) CU_TRY(cudaGetDevice(&device));
  if (device >= count) return fail(B200RT_ERR_INVALID_ARG, "device ordinal out of range");
  int major = 0;
  CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(B200RT_ERR_UNSUPPORTED, "kernels are built for sm_100a (B200) only");
  CU_TRY(cudaSetDevice(device));

  b200rt_scene* sc = new (std::nothrow) b200rt_scene();
  if (!sc) return fail(B200RT_ERR_OOM, "out of host memory");
  struct Guard { b200rt_scene* s; ~Guard() { delete s; } } guard{sc};
  sc->device = device;
  sc->desc = *d;
  CU_TRY(cudaDeviceGetAttribute(&sc->numSMs, cudaDevAttrMultiProcessorCount, device));
  CU_TRY(cudaDeviceGetAttribute(&sc->maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  { int coop = 0; CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device)); sc->cooperativeLaunch = coop != 0; }
  CU_TRY(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreate(&sc->evStart));
  CU_TRY(cudaEventCreate(&sc->evStop));

  // --- derived host tables (geometry lookup, gathered triangles, pair table) + structural validation of the BVH ---
  rt::SceneTables tables;
  {
    const std::string why = rt::build_scene_tables(*d, tables);
    if (!why.empty())
      return fail(why.find("deeper") != std::string::npos ? B200RT_ERR_UNSUPPORTED : B200RT_ERR_INVALID_ARG, why);
  }
  // --- node array: uploaded unchanged ---
  sc->nodeBytes = d->num_bvh_nodes * 24u;
  CU_TRY(sc->nodes.upload(d->bvh_nodes, sc->nodeBytes));
  sc->pairBytes = tables.pairs.numPairs * 48u;
  CU_TRY(sc->pairs.upload(tables.pairs.words.data(), sc->pairBytes));
  CU_TRY(sc->leafOrig.upload(tables.pairs.leafOrig.data(), tables.pairs.leafOrig.size() * 4));
  CU_TRY(sc->leafInfo.upload(tables.pairs.leafInfo.data(), tables.pairs.leafInfo.size() * 4));
  CU_TRY(sc->geoms.upload(tables.geoms.data(), tables.geoms.size() * sizeof(rt::GeomEntry)));
  CU_TRY(sc->triVerts.upload(tables.triVerts.data(), tables.triVerts.size() * sizeof(float)));
  if (d->num_normals) CU_TRY(sc->triNormals.upload(tables.triNormals.data(), tables.triNormals.size() * sizeof(float)));
  CU_TRY(sc->triFaceNormals.upload(tables.triFaceNormals.data(), tables.triFaceNormals.size() * sizeof(float)));
  CU_TRY(sc->spheres.upload(d->spheres, (size_t)d->num_spheres * 16));
  CU_TRY(sc->discs.upload(d->discs, (size_t)d->num_discs * 28));
  CU_TRY(sc->matIDs.upload(d->mat_ids, (size_t)d->num_mat_ids * 4));
  CU_TRY(sc->materials.upload(d->materials, (size_t)d->num_materials * 36));
  CU_TRY(sc->workCounter.upload(nullptr, 16));
  CU_TRY(sc->counters.upload(nullptr, sizeof(rt::DeviceCounters)));

  sc->dev.nodes = (const uint2*)sc->nodes.p;
  sc->dev.geoms = (const rt::GeomEntry*)sc->geoms.p;
  sc->dev.triVerts = (const float4*)sc->triVerts.p;
  sc->dev.triNormals = d->num_normals ? (const float4*)sc->triNormals.p : nullptr;
  sc->dev.triFaceNormals = d->num_tris ? (const float4*)sc->triFaceNormals.p : nullptr;
  sc->dev.spheres = (const float4*)sc->spheres.p;
  sc->dev.discs = (const float*)sc->discs.p;
  sc->dev.matIDs = (const uint32_t*)sc->matIDs.p;
  sc->dev.materials = (const float*)sc->materials.p;
  sc->dev.numNodes = d->num_bvh_nodes;
  sc->dev.numMaterials = d->num_materials;
  sc->dev.pairs = (const uint4*)sc->pairs.p;
  sc->dev.leafOrig = (const uint32_t*)sc->leafOrig.p;
  sc->dev.numPairs = tables.pairs.numPairs;
  sc->dev.rootRef = tables.pairs.rootRef;
  sc->dev.rootGeom = tables.pairs.rootGeom;
  sc->dev.boundsFinite = tables.pairs.boundsFinite ? 1u : 0u;
  sc->dev.leafInfo = (const uint4*)sc->leafInfo.p;
  sc->dev.numTris = d->num_tris;
  sc->dev.numSpheres = d->num_spheres;
  sc->dev.trisBounded = tables.trisBounded ? 1u : 0u;
  {
    struct Node { float mn[3]; uint32_t x; uint16_t d[3]; uint16_t g; };
    const Node& r = *static_cast<const Node*>(d->bvh_nodes);
    for (int k = 0; k < 3; ++k) {
      sc->rootBox[k] = r.mn[k];
      const float e = rt::half_bits_to_float(r.d[k]);
      sc->rootBox[3 + k] = e > 0.f ? e : 1.f;
    }
  }
  // the scalars stay; the host pointers must not be used after creation
  sc->desc.geometry = sc->desc.mesh_info = sc->desc.mesh_tris = sc->desc.mesh_verts = sc->desc.mesh_normals = nullptr;
  sc->desc.mat_ids = nullptr; sc->desc.materials = sc->desc.bvh_nodes = nullptr;
  sc->desc.spheres = sc->desc.discs = nullptr;
  guard.s = nullptr;
  *out = sc;
  return B200RT_OK;
}

void b200rt_scene_destroy(b200rt_scene* sc) { delete sc; }

int b200rt_scene_load_nif(b200rt_scene* sc, const b200rt_nif_desc* nif) {
  if (!sc || !nif) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  CU_TRY(cudaSetDevice(sc->device));
  if (sc->nif) { rt::nif_destroy(sc->nif); sc->nif = nullptr; }
  sc->nif = rt::nif_create(*nif, sc->device);
  if (!sc->nif) return fail(B200RT_ERR_INVALID_ARG, std::string("could not load NIF model: ") + rt::nif_last_error());
  return B200RT_OK;
}

int b200rt_scene_set_hdri_rotation(b200rt_scene* sc, float degrees) {
  if (!sc) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  sc->hdriRotationDegrees = degrees;
  return B200RT_OK;
}

int b200rt_scene_set_max_nif_batch_size(b200rt_scene* sc, size_t n) {
  if (!sc) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  sc->maxNifBatch = n;
  return B200RT_OK;
}

int b200rt_nif_eval(b200rt_scene* sc, const float* uv, size_t n, float* bgrOut) {
  if (!sc || !uv || !bgrOut) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  if (!sc->nif) return fail(B200RT_ERR_INVALID_ARG, "no NIF model loaded");
  CU_TRY(cudaSetDevice(sc->device));
  if (n == 0) return B200RT_OK;
  DeviceBuffer dUv, dOut;
  struct Free { DeviceBuffer& a; DeviceBuffer& b; ~Free() { a.release(); b.release(); } } fr{dUv, dOut};
  CU_TRY(dUv.upload(uv, n * 2 * sizeof(float)));
  CU_TRY(dOut.reserve(n * 3 * sizeof(float)));
  int launches = 0;
  if (rt::nif_eval_uv(sc->nif, (const float*)dUv.p, (uint32_t)n, (float*)dOut.p, sc->stream, &launches) != 0)
    return fail(B200RT_ERR_CUDA, std::string("NIF evaluation failed: ") + rt::nif_last_error());
  CU_TRY(cudaStreamSynchronize(sc->stream));
  CU_TRY(cudaMemcpy(bgrOut, dOut.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  sc->stats.kernel_launches += (uint64_t)launches;
  return B200RT_OK;
}

int b200rt_trace_device(b200rt_scene* sc, const b200rt_trace_params* params, void* dRays, size_t n, void* stream) {
  if (!sc || !dRays) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  CU_TRY(cudaSetDevice(sc->device));
  b200rt_trace_params p{};
  if (params) p = *params;
  sc->stats = b200rt_trace_stats{};
  const auto t0 = std::chrono::steady_clock::now();
  // NULL is the legacy default stream (what a torch "current stream" handle of 0 means), NOT the scene's private
  // non-blocking stream: the render is ordered after the caller's earlier work on that stream.
  const int rc = render_device(*sc, p, (float*)dRays, n, stream ? (cudaStream_t)stream : cudaStreamLegacy);
  sc->stats.trace_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

// Host-resident stream: the reference's load / compute / save pipeline (src/IpuScene.cpp:583-618: while batch k is
// traced, batch k+1 streams in and batch k-1 streams out; >= 2 batches per replica, :102-105). Here the stream is cut
// into tiles of whole ray batches; tile k+1's host->device copy (copy-in stream), tile k's kernels (the scene's stream)
// and tile k-1's device->host copy (copy-out stream) overlap, ordered by events, over a ring of three device buffers.
// A callback is invoked per finished ray batch from a CUDA-owned host thread (cudaLaunchHostFunc on the copy-out
// stream), like RayCallback::fetch on a Poplar runtime thread (src/RayCallback.cpp:8-24).
namespace {
struct CallbackJob {
  b200rt_ray_cb cb;
  void* user;
  size_t index;
  const void* rays;
  size_t n;
};
void CUDART_CB run_callback_job(void* p) {
  const CallbackJob* j = static_cast<const CallbackJob*>(p);
  j->cb(j->index, j->rays, j->n, j->user);
}
}  // namespace

int b200rt_trace(b200rt_scene* sc, const b200rt_trace_params* params, void* rays, size_t n, b200rt_ray_cb cb,
                 void* user) {
  if (!sc || (!rays && n)) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  CU_TRY(cudaSetDevice(sc->device));
  b200rt_trace_params p{};
  if (params) p = *params;
  sc->stats = b200rt_trace_stats{};
  if (n == 0) return B200RT_OK;
  if (n > 0xFFFFFFFFull) return fail(B200RT_ERR_INVALID_ARG, "ray stream longer than 2^32 rays");

  // The batches this call renders: all of them, or (multi-replica callers) every `batch_stride`-th one starting at
  // `first_batch` -- batch i -> replica i % R, src/IpuScene.cpp:676-684 -- read from and written back to their places in
  // the caller's stream, so that R replicas share one host stream without any host-side regrouping.
  // Tiles are made of whole batches (rays per batch = raysPerWorker * 6 * 1440 per replica, src/IpuScene.cpp:78-108).
  const size_t batch = p.rays_per_batch ? p.rays_per_batch : 8640;
  const size_t stride = p.batch_stride ? p.batch_stride : 1, firstBatch = p.first_batch;
  const size_t allBatches = (n + batch - 1) / batch;
  const size_t myBatches = firstBatch < allBatches ? (allBatches - firstBatch + stride - 1) / stride : 0;
  if (myBatches == 0) return B200RT_OK;
  static const size_t envTile = [] { const char* e = std::getenv("B200RT_TILE_RAYS"); return e ? (size_t)std::atoll(e) : (size_t)0; }();
  // single-pass renders are bound by the PCIe copies: tiles small enough that the two copy directions overlap for most
  // of the stream, large enough that the host can enqueue them faster than they drain; multi-sample renders are bound
  // by the kernels: large tiles keep the per-bounce launches large
  // (measured at 1440^2, 174 MB each way: 7.3 / 6.5 / 5.6 / 4.9 / 4.8 ms end to end with tiles of 32 K ... 512 K rays -- below
  // ~256 K rays the ~20 driver calls a tile costs the host are what the pipeline waits for)
  // 8192^2 (5.6 GB each way, PCIe floor 112 ms with both directions busy): 130 / 123 / 120 ms with 512 K / 1 M / 2 M rays.
  const size_t singlePass = std::min<size_t>(std::max<size_t>(n / 16, (size_t)1 << 19), (size_t)1 << 21);
  const size_t wantTile = envTile ? envTile : (sc->desc.path_trace ? (size_t)1 << 20 : singlePass);
  const size_t tileBatches = std::max<size_t>(1, wantTile / batch);
  const size_t numTiles = (myBatches + tileBatches - 1) / tileBatches;
  constexpr int kRing = StreamPipe::kRing;
  const int ring = (int)std::min<size_t>(kRing, numTiles);
  const size_t tileBytes = std::min(tileBatches, myBatches) * batch * 84;
  CU_TRY(sc->rays.reserve(tileBytes * ring));
  StreamPipe& pipe = sc->pipe;
  CU_TRY(pipe.init());
  pipe.timingUsed = 0;
  auto timing_event = [&](cudaStream_t st) -> cudaError_t { return pipe.record_timing(st); };
  std::vector<CallbackJob> jobs;
  if (cb) jobs.resize(myBatches);
  // rays of my j-th batch
  auto batch_lo = [&](size_t j) { return (firstBatch + j * stride) * batch; };
  auto batch_len = [&](size_t j) { return std::min(batch, n - batch_lo(j)); };

  const auto t0 = std::chrono::steady_clock::now();
  RenderRun run(*sc);
  // an early error return must not leave copies, kernels or queued callbacks (which point into `jobs`) in flight
  struct Drain {
    b200rt_scene& s; bool armed = true;
    ~Drain() { if (armed) { cudaStreamSynchronize(s.pipe.in); cudaStreamSynchronize(s.stream); if (s.nifStream) cudaStreamSynchronize(s.nifStream); cudaStreamSynchronize(s.pipe.out); } }
  } drain{*sc};
  if (int rc = render_begin(*sc, run)) return rc;
  for (size_t k = 0; k < numTiles; ++k) {
    const int slot = (int)(k % (size_t)ring);
    const size_t j0 = k * tileBatches, j1 = std::min(myBatches, j0 + tileBatches);
    size_t cnt = 0;
    for (size_t j = j0; j < j1; ++j) cnt += batch_len(j);
    char* dbuf = (char*)sc->rays.p + (size_t)slot * tileBytes;
    // host -> HBM (the reference's copyToRemoteBuffer, src/IpuScene.cpp:676-684), once the slot's last read-back is done
    if (k >= (size_t)ring) CU_TRY(cudaStreamWaitEvent(pipe.in, pipe.evOut[slot], 0));
    CU_TRY(timing_event(pipe.in));
    if (stride == 1) {
      CU_TRY(cudaMemcpyAsync(dbuf, (char*)rays + batch_lo(j0) * 84, cnt * 84, cudaMemcpyHostToDevice, pipe.in));
    } else {
      for (size_t j = j0; j < j1; ++j)
        CU_TRY(cudaMemcpyAsync(dbuf + (j - j0) * batch * 84, (char*)rays + batch_lo(j) * 84, batch_len(j) * 84,
                               cudaMemcpyHostToDevice, pipe.in));
    }
    CU_TRY(timing_event(pipe.in));
    CU_TRY(cudaEventRecord(pipe.evIn[slot], pipe.in));
    // kernels
    CU_TRY(cudaStreamWaitEvent(sc->stream, pipe.evIn[slot], 0));
    if (int rc = render_enqueue(*sc, p, (float*)dbuf, cnt, run)) return rc;
    CU_TRY(cudaEventRecord(pipe.evRender[slot], sc->stream));
    // HBM -> host, batch by batch when a callback wants the reference's batch granularity
    CU_TRY(cudaStreamWaitEvent(pipe.out, pipe.evRender[slot], 0));
    CU_TRY(timing_event(pipe.out));
    if (!cb && stride == 1) {
      CU_TRY(cudaMemcpyAsync((char*)rays + batch_lo(j0) * 84, dbuf, cnt * 84, cudaMemcpyDeviceToHost, pipe.out));
    } else {
      for (size_t j = j0; j < j1; ++j) {
        char* hb = (char*)rays + batch_lo(j) * 84;
        CU_TRY(cudaMemcpyAsync(hb, dbuf + (j - j0) * batch * 84, batch_len(j) * 84, cudaMemcpyDeviceToHost, pipe.out));
        if (cb) {
          jobs[j] = CallbackJob{cb, user, firstBatch + j * stride, hb, batch_len(j)};  // k * R + replica, RayCallback.cpp:8-24
          CU_TRY(cudaLaunchHostFunc(pipe.out, run_callback_job, &jobs[j]));
        }
      }
    }
    CU_TRY(timing_event(pipe.out));
    CU_TRY(cudaEventRecord(pipe.evOut[slot], pipe.out));
  }
  CU_TRY(cudaStreamSynchronize(pipe.out));
  if (int rc = render_end(*sc, run)) return rc;
  drain.armed = false;
  sc->stats.trace_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // copy time = sum over tiles (the copies of different tiles overlap the kernels, not each other)
  for (size_t i = 0; i + 4 <= pipe.timingUsed; i += 4) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pipe.timing[i], pipe.timing[i + 1]) == cudaSuccess) sc->stats.h2d_ms += ms;
    if (cudaEventElapsedTime(&ms, pipe.timing[i + 2], pipe.timing[i + 3]) == cudaSuccess) sc->stats.d2h_ms += ms;
  }
  return B200RT_OK;
}

int b200rt_get_trace_stats(const b200rt_scene* sc, b200rt_trace_stats* out) {
  if (!sc || !out) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  *out = sc->stats;
  return B200RT_OK;
}

double b200rt_get_trace_time_secs(const b200rt_scene* sc) { return sc ? sc->stats.trace_secs : 0.0; }

int b200rt_intersect(b200rt_scene* sc, const void* raysIn, size_t n, b200rt_hit* hitsOut, uint32_t traversal) {
  if (!sc || !raysIn || !hitsOut) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  CU_TRY(cudaSetDevice(sc->device));
  sc->stats = b200rt_trace_stats{};
  if (n == 0) return B200RT_OK;
  DeviceBuffer dR, dH;
  struct Free { DeviceBuffer& a; DeviceBuffer& b; ~Free() { a.release(); b.release(); } } fr{dR, dH};
  CU_TRY(dR.upload(raysIn, n * 32));
  CU_TRY(dH.reserve(n * sizeof(rt::QueryHit)));
  CU_TRY(cudaMemsetAsync(sc->counters.p, 0, sizeof(rt::DeviceCounters), sc->stream));
  const uint32_t threads = 128, blocks = (uint32_t)((n + threads - 1) / threads);
  // Bare queries may carry non-unit directions, for which only the reference's own visiting order is
  // guaranteed to reproduce its answers (see DESIGN.md "Traversal order"): auto = reference order here.
  if (traversal != 2)
    rt::intersect_kernel<false><<<blocks, threads, 0, sc->stream>>>(sc->dev, (const float*)dR.p, (uint32_t)n,
                                                                   (rt::QueryHit*)dH.p, (rt::DeviceCounters*)sc->counters.p);
  else
    rt::intersect_kernel<true><<<blocks, threads, 0, sc->stream>>>(sc->dev, (const float*)dR.p, (uint32_t)n,
                                                                  (rt::QueryHit*)dH.p, (rt::DeviceCounters*)sc->counters.p);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(sc->stream));
  static_assert(sizeof(rt::QueryHit) == sizeof(b200rt_hit), "hit record");
  CU_TRY(cudaMemcpy(hitsOut, dH.p, n * sizeof(b200rt_hit), cudaMemcpyDeviceToHost));
  rt::DeviceCounters hc{};
  CU_TRY(cudaMemcpy(&hc, sc->counters.p, sizeof(hc), cudaMemcpyDeviceToHost));
  sc->stats.closest_hit_queries = hc.closest;
  sc->stats.node_visits = hc.nodeVisits;
  sc->stats.prim_tests = hc.primTests;
  sc->stats.kernel_launches = 1;
  return B200RT_OK;
}

int b200rt_occluded(b200rt_scene* sc, const void* raysIn, size_t n, uint8_t* out) {
  if (!sc || !raysIn || !out) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  CU_TRY(cudaSetDevice(sc->device));
  if (n == 0) return B200RT_OK;
  DeviceBuffer dR, dO;
  struct Free { DeviceBuffer& a; DeviceBuffer& b; ~Free() { a.release(); b.release(); } } fr{dR, dO};
  CU_TRY(dR.upload(raysIn, n * 32));
  CU_TRY(dO.reserve(n));
  const uint32_t threads = 128, blocks = (uint32_t)((n + threads - 1) / threads);
  rt::occluded_kernel<<<blocks, threads, 0, sc->stream>>>(sc->dev, (const float*)dR.p, (uint32_t)n, (unsigned char*)dO.p);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(sc->stream));
  CU_TRY(cudaMemcpy(out, dO.p, n, cudaMemcpyDeviceToHost));
  return B200RT_OK;
}

int b200rt_host_register(void* rays, size_t bytes) {
  if (!rays || !bytes) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  // Portable: every device of a multi-GPU render DMA's from the same stream.
  const cudaError_t e = cudaHostRegister(rays, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return B200RT_OK; }
  CU_TRY(e);
  return B200RT_OK;
}

int b200rt_host_unregister(void* rays) {
  if (!rays) return fail(B200RT_ERR_INVALID_ARG, "null argument");
  const cudaError_t e = cudaHostUnregister(rays);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return B200RT_OK; }
  CU_TRY(e);
  return B200RT_OK;
}

}  // extern "C"
