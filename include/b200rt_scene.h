/*
 * b200rt_scene.h — host-side scene utilities (pure C++ behind a C API, no CUDA).
 *
 * These are the producers and consumers either side of the trace path (SURVEY.md §8f "next"
 * rows N1, N2, N4): they build the arrays b200rt_scene_desc points at and turn a finished ray
 * stream into images. They replace, for this path only:
 *   makeCornellBoxScene / makePrimitiveScene     src/scene_utils.cpp:458-597
 *   importMesh / importScene (assimp)            src/scene_utils.cpp:102-317
 *   buildSceneData + Embree rtcBuildBVH + flatten src/app_utils.cpp:291-364, include/embree_utils/bvh.hpp:46-69,
 *                                                 src/CompactBvhBuild.cpp:5-56
 *   initPerspectiveRayStream / zeroRgb / scaleRgb src/app_utils.cpp:19-59
 *   visualiseHits                                 src/app_utils.cpp:61-127
 *   cv::imwrite(.exr)                             trace.cpp:503-523
 */
#ifndef B200RT_SCENE_H
#define B200RT_SCENE_H
#include "b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200rt_host_scene b200rt_host_scene;

/* name: "box", "box-simple" or "spheres" (trace.cpp --scene). mesh_file: path of monkey_bust.glb,
 * needed by "box" only (src/app_utils.cpp:263-264). */
int  b200rt_host_scene_builtin(const char* name, const char* mesh_file, b200rt_host_scene** out);
/* --mesh-file: COLLADA 1.4.1 (.dae) or glTF binary (.glb) with a camera (importScene). */
int  b200rt_host_scene_import(const char* file, int load_normals, b200rt_host_scene** out);
void b200rt_host_scene_free(b200rt_host_scene* s);
/* Fills every array pointer/count, max_leaf_depth and fov_radians of `out` (arrays stay owned by
 * the handle). Render scalars are set to the reference CLI defaults (trace.cpp:338-378): anti-alias
 * 0.25, max path length 10, roulette start 3, 256 spp, seed 1442. */
int  b200rt_host_scene_desc(const b200rt_host_scene* s, b200rt_scene_desc* out);
const char* b200rt_scene_last_error(void);

/* Stand-alone BVH2 build (binned SAH, 1 primitive per leaf, pre-order flatten, fp16 extents rounded
 * up). prim_bounds: n x {minx,miny,minz,maxx,maxy,maxz}; ids: n x {geomID, primID}.
 * nodes_out must hold 2n-1 CompactBVH2Node (24 B each). Returns node count (<0 on error). */
int  b200rt_build_bvh(const float* prim_bounds, const uint32_t* ids, uint32_t n,
                      void* nodes_out, uint32_t* max_depth_out);

/* initPerspectiveRayStream (no jitter) + zeroRgb for the crop window (w,h,+c,+r) of a img_w x img_h image. */
int  b200rt_init_ray_stream(void* rays /*TraceResult[w*h]*/, int img_w, int img_h, int win_w, int win_h,
                            int win_c, int win_r, float fov_radians);
void b200rt_scale_rgb(void* rays, size_t n, float scale);

/* visualiseHits: mode 0 rgb, 1 id, 2 normal, 3 tfar, 4 color, 5 hitpoint (enum VisualiseMode,
 * include/app_utils.hpp:26-33). image_bgr: img_h x img_w x 3 fp32, BGR like cv::Mat. Returns hit count. */
long b200rt_visualise_hits(const void* rays, size_t n, const b200rt_scene_desc* scene, int mode,
                           float* image_bgr, int img_w, int img_h);
/* Image writers: uncompressed scanline OpenEXR (fp32 B,G,R channels) and PFM. */
int  b200rt_write_exr(const char* path, const float* image_bgr, int w, int h);
int  b200rt_write_pfm(const char* path, const float* image_bgr, int w, int h);

/* NIF metadata (nif_metadata.txt JSON, src/neural_networks/NifMetaData.cpp:11-71): embedding
 * dimension, hidden size, max, mean (eps folded in when log tone-mapped), log_tone_map. */
typedef struct b200rt_nif_metadata {
  uint32_t embedding_dimension, hidden_size;
  uint32_t image_shape[3];
  float max, eps, mean[3];
  int32_t log_tone_map;
} b200rt_nif_metadata;
int  b200rt_read_nif_metadata(const char* path, b200rt_nif_metadata* out);

/* NIF weights: the Keras "h5" model the reference loads (`<assets.extra>/converted.hdf5`, src/IpuScene.cpp:177) through
 * Hdf5Model (src/keras/Hdf5Model.cpp:8-133): root attributes keras_version / backend / model_config (a "Functional"
 * model; Dense layers kept in order, InputLayer / Concatenate ignored, anything else an error) and
 * /model_weights/<name>/<name>/{kernel:0, bias:0}. Read by a self-contained reader of the classic HDF5 layout
 * (host/keras_hdf5.cpp; libhdf5 is not a dependency); float32 datasets are rounded to fp16. The returned layers point
 * into the model handle and stay valid until b200rt_keras_hdf5_close. */
typedef struct b200rt_keras_model b200rt_keras_model;
typedef struct b200rt_keras_layer {
  char name[64];
  char activation[32];
  b200rt_nif_layer layer;   /* in/out features, fp16 kernel [in][out], fp16 bias or NULL, relu = (activation == "relu") */
} b200rt_keras_layer;
int         b200rt_keras_hdf5_open(const char* path, b200rt_keras_model** out);
void        b200rt_keras_hdf5_close(b200rt_keras_model* model);
uint32_t    b200rt_keras_hdf5_num_layers(const b200rt_keras_model* model);
int         b200rt_keras_hdf5_layer(const b200rt_keras_model* model, uint32_t index, b200rt_keras_layer* out);
const char* b200rt_keras_hdf5_version(const b200rt_keras_model* model);
const char* b200rt_keras_last_error(void);

/* The reference's serialised scene: `SceneRef` written by Serialiser<16> (include/serialisation/Serialiser.hpp:24-64,
 * serialisation.hpp:34-52) -- the byte stream IpuScene uploads (src/IpuScene.cpp:52, :665) and the device reads back
 * in place (deserialisation.hpp:29-59). b200rt_scene_desc_from_blob fills `out` with pointers INTO `blob` (zero copy:
 * the blob must outlive the desc and be 16-byte aligned, like the reference's buffer). The stream carries neither the
 * spheres / discs nor rng_seed, path_trace, device (src/IpuScene.cpp:208-215): those fields are left zero for the
 * caller. b200rt_scene_blob_write is the matching writer; it returns the size and writes when `cap` suffices. */
int    b200rt_scene_desc_from_blob(const void* blob, size_t bytes, b200rt_scene_desc* out);
size_t b200rt_scene_blob_write(const b200rt_scene_desc* scene, void* out, size_t cap);

/* Host sincos used for tan(fov/2) (ext/math/sincos.cpp:236); exported for tests. */
void b200rt_sincos(float x, float* s, float* c);

#ifdef __cplusplus
}
#endif
#endif
