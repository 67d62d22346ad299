/*
 * oracle_api.h — C entry points shared by the two CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (ipu_ray_lib_b200/, include/) may include,
 * link or call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it.
 *
 * Two shared objects export this API, distinguished by symbol prefix:
 *   orc_*  oracle/liboracle_port.so   — the restatement (oracle/oracle_port.cpp), "port"
 *   ref_*  oracle/_ref/liboracle_ref.so — the reference's OWN kernel sources compiled where
 *          they lie under /root/reference (src/Mesh.cpp, src/Primitives.cpp,
 *          src/CompactBVH2Node.cpp, ext/math/sincos.cpp + include/ *.hpp) behind a thin driver
 *          (oracle/ref_driver.cpp), "reference"
 * The scene is described by the product's public b200rt_scene_desc so that the very same
 * arrays feed the oracle and the GPU.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stddef.h>
#include <stdint.h>
#include "../include/b200rt.h"

#ifndef ORC_PREFIX
#define ORC_PREFIX orc_
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC(name) ORC_CAT(ORC_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* counters[0]=closest-hit queries, [1]=occlusion queries, [2]=node visits (port only, 0 in ref),
 * [3]=primitive tests, [4]=samples, [5]=escaped samples. May be NULL. */

/* traceShadowRay over a ray stream (include/Render.hpp:37-72 driven as trace.cpp:246-255). */
int ORC(shadow_trace)(const b200rt_scene_desc* scene, void* rays /*TraceResult[n]*/, size_t n,
                      const float light_pos[3], float ambient, int threads, uint64_t* counters);

/* pathTrace (trace.cpp:115-188 / codelets/TraceCodelets.cpp:184-263) with in-kernel camera
 * sampling (codelets/TraceCodelets.cpp:142-164) and the build-defined per-(pixel,sample) RNG
 * streams (ipu_ray_lib_b200/csrc/rt_math.h). Samples first_sample .. first_sample+num_samples-1.
 * `nif` may be NULL (escaped rays just stop, trace.cpp:171-174); the ref build ignores it. */
int ORC(path_trace)(const b200rt_scene_desc* scene, void* rays, size_t n, uint32_t first_sample,
                    uint32_t num_samples, const b200rt_nif_desc* nif, float hdri_rotation_degrees,
                    int threads, uint64_t* counters);

/* CompactBvh::intersect / ::occluded on bare rays (include/CompactBvh.hpp:80-139, :33-78). */
int ORC(intersect)(const b200rt_scene_desc* scene, const void* rays /*Ray[n]*/, size_t n,
                   b200rt_hit* hits_out, int threads, uint64_t* counters);
int ORC(occluded)(const b200rt_scene_desc* scene, const void* rays, size_t n, uint8_t* out, int threads);

/* Known-answer helpers for unit-level pinning of port vs reference. */
void ORC(sincos)(const float* x, size_t n, float* s, float* c);                 /* ext/math/sincos.cpp:236 */
void ORC(uniform_stream)(uint64_t seed, size_t n, float* out);                  /* xoshiro::Generator(seed).uniform_0_1() */
void ORC(raw_stream)(uint64_t seed, size_t n, uint64_t* out);                   /* next128ss */
void ORC(sample_diffuse)(const float* normals, const float* u12, size_t n, float* dirs_out); /* BxDF.hpp:11-30 */
void ORC(dielectric)(const float* dirs, const float* normals, const float* ior_u1, size_t n,
                     float* dirs_out, uint8_t* refracted_out);                  /* BxDF.hpp:57-75 */
void ORC(reflect)(const float* dirs, const float* normals, size_t n, float* dirs_out); /* BxDF.hpp:33-37 */
void ORC(offset_ray)(const float* origins, const float* dirs, const float* normals, size_t n, float* origins_out); /* Render.hpp:29-33 */
void ORC(pixel_to_ray_dir)(const float* xy, size_t n, float w, float h, float tan_theta, float* dirs_out); /* Render.hpp:74-85 */
void ORC(round_to_half_not_smaller)(const float* x, size_t n, uint16_t* out);  /* precision_utils.hpp:40-47 */
/* Camera sample of (row, col, sample): the jittered direction the path tracer starts from. */
void ORC(camera_sample)(uint64_t rng_seed, uint32_t image_width, uint32_t image_height, float fov_radians,
                        float anti_alias_scale, const uint32_t* row_col_sample /*[n][3]*/, size_t n, float* dirs_out);
/* NIF forward for n (u,v) pairs -> bgr (port only; the reference has no CPU NIF). */
int ORC(nif_eval)(const b200rt_nif_desc* nif, const float* uv, size_t n, float* bgr_out, int threads);
/* Same with the IPU's matmul option partialsType = half modelled (src/IpuScene.cpp:256-262): the running sum is rounded
 * to fp16 after every `half_chunk` products (0 = fp32 accumulation = nif_eval). Port only; a model, not a pinned
 * restatement (poplibs' accumulation order is unpublished). */
int ORC(nif_eval_partials)(const b200rt_nif_desc* nif, const float* uv, size_t n, float* bgr_out, int threads, int half_chunk);
/* Equirect (u,v) of a direction (codelets/TraceCodelets.cpp:337-348); port only. */
void ORC(dir_to_uv)(const float* dirs, size_t n, float rotation_radians, float* uv_out);

/* SceneRef serialised with Serialiser<16> (include/serialisation/serialisation.hpp:34-52): the byte stream the reference
 * uploads to the device (src/IpuScene.cpp:52, :665). Returns the size; writes it when cap is large enough. */
size_t ORC(serialise_scene)(const b200rt_scene_desc* scene, uint8_t* out, size_t cap);

const char* ORC(kind)(void);  /* "port" or "reference" */

#ifdef __cplusplus
}
#endif
#endif
