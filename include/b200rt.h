/*
 * b200rt.h — C ABI of the B200-native ray-data-parallel trace path.
 *
 * This is the drop-in boundary: every entry point replaces one piece of the
 * reference's device orchestration object `IpuScene` (reference file:line is
 * cited per function). Plain pointers and sizes only; no C++/torch types.
 *
 * Array layouts are byte-for-byte the reference's:
 *   GeomRef            4 B  {u16 index; u8 type(0=Mesh,1=Sphere,2=Disc); u8 pad}     include/Scene.hpp:27-32
 *   MeshInfo          16 B  {u32 firstIndex, firstVertex, numTriangles, numVertices} include/Mesh.hpp:15-20
 *   Triangle           6 B  {u16 v0, v1, v2} (indices into the mesh's vertex window) include/Primitives.hpp:21-25
 *   Vec3fa            12 B  {f32 x, y, z}                                            include/embree_utils/geometry.hpp:27
 *   Material          36 B  {Vec3fa albedo; f32 ior; Vec3fa emission; i32 type; u8 emissive; pad[3]}
 *                                                                                   include/Material.hpp:8-33
 *   CompactBVH2Node   24 B  {f32 min[3]; u32 primID|secondChild; f16 d[3]; u16 geomID}
 *                                                                                   include/CompactBVH2Node.hpp:52-85
 *   TraceResult       84 B  {Vec3fa rgb; f32 row, col; HitRecord h(64 B)}            include/embree_utils/geometry.hpp:227-260
 * so the output of the reference's CompactBvhBuild loads unchanged.
 *
 * Error model: no function throws. Each returns 0 on success or a negative
 * b200rt_status; b200rt_last_error() returns a thread-local message
 * (the reference throws std::runtime_error / returns EXIT_FAILURE from
 * GraphManager::run, include/ipu_utils.hpp:590-593).
 *
 * Threading: a scene is bound to one CUDA device (one replica <-> one device,
 * as in src/IpuScene.cpp:676-684). Calls on one scene must be serialised by the
 * caller; distinct scenes are independent. There is NO CPU fallback: every
 * compute entry point fails with B200RT_ERR_CUDA when no sm_100 device exists.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_ABI_VERSION 1

typedef enum b200rt_status {
  B200RT_OK = 0,
  B200RT_ERR_INVALID_ARG = -1,
  B200RT_ERR_CUDA = -2,
  B200RT_ERR_OOM = -3,
  B200RT_ERR_UNSUPPORTED = -4,
  B200RT_ERR_IO = -5
} b200rt_status;

typedef struct b200rt_scene b200rt_scene;

/* Mirrors `SceneRef` (include/Scene.hpp:50-74) plus the sphere/disc arrays the
 * reference passes beside it (IpuScene ctor, src/IpuScene.cpp:24-29).
 * The caller keeps ownership of every array; the library copies to device. */
typedef struct b200rt_scene_desc {
  const void* geometry;      uint32_t num_geometry;   /* GeomRef[]  */
  const void* mesh_info;     uint32_t num_meshes;     /* MeshInfo[] */
  const void* mesh_tris;     uint32_t num_tris;       /* Triangle[] */
  const void* mesh_verts;    uint32_t num_verts;      /* Vec3fa[]   */
  const void* mesh_normals;  uint32_t num_normals;    /* Vec3fa[]; 0 or == num_verts */
  const uint32_t* mat_ids;   uint32_t num_mat_ids;    /* per geomID */
  const void* materials;     uint32_t num_materials;  /* Material[] */
  const void* bvh_nodes;     uint32_t num_bvh_nodes;  /* CompactBVH2Node[] */
  uint32_t max_leaf_depth;                            /* traversal stack bound */
  const float* spheres;      uint32_t num_spheres;    /* {x,y,z,radius} per sphere      (include/Primitives.hpp:36-57) */
  const float* discs;        uint32_t num_discs;      /* {nx,ny,nz,r,cx,cy,cz} per disc (include/Primitives.hpp:59-82) */
  /* render parameters (SceneRef scalars) */
  float image_width, image_height;   /* FULL image size, not the crop window */
  float fov_radians;
  float anti_alias_scale;
  uint32_t max_path_length;
  uint32_t roulette_start_depth;
  uint32_t samples_per_pixel;
  uint64_t rng_seed;
  int32_t path_trace;                /* 0 = shadow-trace, 1 = path-trace */
  int32_t device;                    /* CUDA device ordinal; -1 = current */
} b200rt_scene_desc;

/* Per-trace options. Zero-initialise for defaults. */
typedef struct b200rt_trace_params {
  float light_pos[3];          /* shadow-trace point light; reference hard-codes (18,257,-1060) trace.cpp:247 */
  float ambient;               /* shadow-trace ambient factor; reference 0.05 trace.cpp:252 */
  uint32_t first_sample;       /* path-trace: first sample index (for resuming); default 0 */
  uint32_t num_samples;        /* path-trace: samples to take; 0 = scene.samples_per_pixel */
  uint32_t rays_per_batch;     /* callback granularity; 0 = 8640 * rays_per_worker default (src/IpuScene.cpp:78-108) */
  uint32_t traversal;          /* 0 = auto (path-trace renders: 4; shadow-trace renders: 2; bare queries: 1), 1 = reference-order DFS (identical visiting
                                * order, identical answers for ANY ray), 2 = near-first ordered DFS (same answers for
                                * unit-length directions, which is all a render produces; ties go to the lowest leaf
                                * index like the reference's pre-order walk), 3 = alias of 4 (round 1's state-machine
                                * megakernel, measured slower, was removed),
                                * 4 = wavefront path tracer: per-bounce trace / shade kernels over path queues in HBM with
                                * near-first traversal over the derived two-children ("pair") node table (same arithmetic
                                * and results as 2) */
  uint32_t scene_residency;    /* 0 = auto, 1 = BVH staged in shared memory, 2 = global/L2-resident */
  uint32_t samples_per_chunk;  /* path-trace: samples per chunk of the wavefront / NIF pipeline; 0 = auto (~64 M paths per chunk) */
  uint32_t count_visits;       /* 1 = also count node visits / primitive tests (slower; parity tests) */
  uint32_t primary_pass;       /* path-trace, traversal 1/2: 0 = auto (off), 1 = on, 2 = off. On: the camera rays of a chunk
                                * are traced by a separate warp-coherent kernel (27.5 of 32 lanes active) and the path
                                * tracer starts from the parked hits: same rays, same hits, same results. Measured on
                                * the box scene: pre-pass 4.5 ms + path tracer 47.5 ms vs 52.0 ms without -- no net
                                * gain, because a warp's cost per bounce is its slowest bounce ray either way */
  uint32_t batch_stride;       /* b200rt_trace only. 0 / 1 = the whole stream. R > 1: this call renders the ray batches
                                * first_batch, first_batch + R, ... of the caller's stream in place (batch i -> replica
                                * i % R, src/IpuScene.cpp:676-684); R scenes on R devices can share one host stream from
                                * R threads without regrouping it. The callback's batch_index is the batch's index in
                                * the whole stream (= k * R + replica, src/RayCallback.cpp:8-24) */
  uint32_t first_batch;
  uint32_t tail_bounce;        /* wavefront path tracer: the bounces from this one on run in ONE cooperative launch
                                * (trace phase, grid barrier, shade phase, grid barrier, ... until no path is left) instead
                                * of two launches per bounce. 0 = auto (roulette_start_depth + 2: with Russian roulette the
                                * paths still alive by then are a fraction of a per cent of a chunk), 1 = never (one trace
                                * and one shade launch per bounce), N >= 2 = from bounce N. Same results either way */
  uint32_t chunk_overlap;      /* NIF-lit wavefront renders of more than one chunk: 0 = auto, 1 = off, 2 = on. On: the NIF MLP
                                * and the accumulate of chunk c run on a second CUDA stream while chunk c + 1 is traced
                                * and shaded (two sets of per-sample records); accumulates stay in chunk order, results
                                * are bit-identical */
} b200rt_trace_params;

/* Counters of the last trace (device-side counted, exact). */
typedef struct b200rt_trace_stats {
  uint64_t closest_hit_queries;   /* CompactBvh::intersect calls issued   */
  uint64_t occlusion_queries;     /* CompactBvh::occluded calls issued    */
  uint64_t node_visits;           /* BVH nodes popped and slab-tested     */
  uint64_t prim_tests;            /* leaf primitive intersection tests    */
  uint64_t samples;               /* camera paths started                 */
  uint64_t escaped_samples;       /* paths that ended with ESCAPED (NIF lookups) */
  double   kernel_ms;             /* device time of all kernels (CUDA events on the trace stream) */
  double   h2d_ms, d2h_ms;        /* host<->device copies of the ray stream (host-buffer entry points only) */
  double   trace_secs;            /* wall time; same span as IpuScene::getTraceTimeSecs (src/IpuScene.cpp:672-696) */
  uint64_t kernel_launches;       /* number of kernels this library launched */
  double   trace_kernel_ms;       /* sum over launches of shadow_trace / path_trace / wf_trace kernel time (CUDA events) */
  double   nif_kernel_ms;         /* sum over launches of the NIF MLP kernel time. With chunk_overlap on, the NIF of one
                                   * chunk shares the SMs with the trace / shade kernels of the next: the per-kernel sums
                                   * are event spans that include that sharing and add up to more than kernel_ms */
  double   accumulate_kernel_ms;  /* sum over launches of the ordered rgb accumulation kernel */
  uint64_t trace_kernel_launches, nif_kernel_launches;
  double   shade_kernel_ms;       /* wavefront path tracer: sum over launches of the shade kernels */
  uint64_t shade_kernel_launches;
} b200rt_trace_stats;

/* Called once per finished ray batch with (batch_index, rays, n, user), as soon as that batch's
 * device->host copy has landed and while later tiles are still being traced.
 * Mirrors IpuScene::RayCallbackFn (include/IpuScene.hpp:31) / RayCallback::fetch
 * (src/RayCallback.cpp:8-24). Invoked on a CUDA-owned host thread (cudaLaunchHostFunc), NOT on the
 * caller's thread: it must be thread-safe and must not call this library or any CUDA API.
 * All callbacks have returned when b200rt_trace returns. */
typedef void (*b200rt_ray_cb)(size_t batch_index, const void* rays, size_t n, void* user);

/* --- library --- */
int         b200rt_abi_version(void);
const char* b200rt_last_error(void);
/* Number of usable sm_100 devices (0 on a CPU-only host; never an error). */
int         b200rt_device_count(void);
/* CUDA ordinal of the index-th usable sm_100 device (for b200rt_scene_desc.device), or -1. */
int         b200rt_device_ordinal(int index);

/* --- scene lifetime: IpuScene ctor/dtor (src/IpuScene.cpp:24-64) --- */
int  b200rt_scene_create(const b200rt_scene_desc* desc, b200rt_scene** out);
void b200rt_scene_destroy(b200rt_scene* scene);

/* --- NIF environment light --- */
/* IpuScene::loadNifModel (src/IpuScene.cpp:174-187). `layers` describes Dense layers
 * in order, exactly as the reference's Hdf5Model yields them (src/keras/Hdf5Model.cpp:61-85):
 * kernel [in,out] row-major fp16, optional bias [out] fp16, relu or linear. The skip
 * concat is auto-detected where layer input width != previous width, as
 * src/neural_networks/NifModel.cpp:303-309. */
typedef struct b200rt_nif_layer {
  uint32_t in_features, out_features;
  const uint16_t* kernel_f16;   /* [in_features][out_features] */
  const uint16_t* bias_f16;     /* [out_features] or NULL */
  int32_t relu;                 /* 1 = relu, 0 = linear */
} b200rt_nif_layer;
typedef struct b200rt_nif_desc {
  uint32_t embedding_dimension;     /* 12 in the shipped metadata */
  uint32_t num_layers;
  const b200rt_nif_layer* layers;
  float max;                        /* encode_params.max  */
  float mean[3];                    /* encode_params.mean with eps already folded in (NifMetaData.cpp:48-53) */
  int32_t log_tone_map;
} b200rt_nif_desc;
int b200rt_scene_load_nif(b200rt_scene* scene, const b200rt_nif_desc* nif);
/* IpuScene::setHdriRotation (src/IpuScene.cpp:334-336), degrees. */
int b200rt_scene_set_hdri_rotation(b200rt_scene* scene, float degrees);
/* IpuScene::setMaxNifBatchSize (src/IpuScene.cpp:342-344, used at :265-327): the escaped rays of a chunk are looked up
 * in serial NIF launches of at most this many rays; 0 = auto (one launch per chunk). Results do not depend on it. */
int b200rt_scene_set_max_nif_batch_size(b200rt_scene* scene, size_t rays_per_batch);
/* Stand-alone NIF evaluation of n (u,v) pairs -> n bgr triples (fp32), for parity tests:
 * the encode + MLP + decode of src/neural_networks/NifModel.cpp:186-246,296-327.
 * Host pointers. */
int b200rt_nif_eval(b200rt_scene* scene, const float* uv, size_t n, float* bgr_out);

/* --- tracing: IpuScene::execute (src/IpuScene.cpp:628-733) --- */
/* `rays` is TraceResult[n] in HOST memory (pageable or pinned), pre-initialised by the
 * caller exactly as renderIPU does (initPerspectiveRayStream + zeroRgb, trace.cpp:276-278).
 * Results are written back in place; rgb is a running SUM over samples (caller scales by
 * 1/spp, trace.cpp:324-326). `cb` may be NULL (bulk read-back, src/IpuScene.cpp:699-711). */
int b200rt_trace(b200rt_scene* scene, const b200rt_trace_params* params,
                 void* rays, size_t n, b200rt_ray_cb cb, void* user);
/* Same, with TraceResult[n] already resident in DEVICE memory of the scene's device.
 * `stream` is a cudaStream_t; every kernel is enqueued on it, so the render is ordered after the
 * caller's earlier work on that stream (NULL = the legacy default stream, which is also what a
 * framework's "default stream" handle of 0 denotes). Returns after synchronising the stream. */
int b200rt_trace_device(b200rt_scene* scene, const b200rt_trace_params* params,
                        void* d_rays, size_t n, void* stream);

/* Page-lock / release a caller-owned ray stream so that b200rt_trace's copies run as DMA at PCIe speed (pageable
 * memory works too, at a fraction of the bandwidth). The reference pins nothing itself -- Poplar's stream callbacks
 * read the caller's vector in place (src/IpuScene.cpp:399-409, :703-707); this is the CUDA equivalent of that
 * contract. Returns B200RT_OK, or an error the caller may ignore (tracing still works unpinned). */
int b200rt_host_register(void* rays, size_t bytes);
int b200rt_host_unregister(void* rays);

int    b200rt_get_trace_stats(const b200rt_scene* scene, b200rt_trace_stats* out);
/* IpuScene::getTraceTimeSecs (include/IpuScene.hpp:55). */
double b200rt_get_trace_time_secs(const b200rt_scene* scene);

/* --- single-query entry points used by parity tests (CompactBvh::intersect / ::occluded,
 * include/CompactBvh.hpp:80-139 / :33-78). rays_in: Ray[n] (32 B each). Host pointers. --- */
typedef struct b200rt_hit {
  float t;              /* closest t, or the ray's tMax when nothing was hit */
  uint32_t geom_id;     /* 0xFFFF when nothing was hit */
  uint32_t prim_id;     /* 0xFFFFFFFF when nothing was hit */
  float normal[3];      /* primitive normal at the hit (Primitive::normal) */
} b200rt_hit;
int b200rt_intersect(b200rt_scene* scene, const void* rays_in, size_t n, b200rt_hit* hits_out, uint32_t traversal);
int b200rt_occluded(b200rt_scene* scene, const void* rays_in, size_t n, uint8_t* occluded_out);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
