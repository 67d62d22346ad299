"""ctypes bindings of the two in-tree shared libraries.

* ``libb200rt.so``        — the CUDA trace path behind the C ABI of ``include/b200rt.h``
* ``libb200rt_scene.so``  — host-side scene utilities, ``include/b200rt_scene.h``

The structures below mirror the C headers field for field; the numpy dtypes mirror the
reference's wire layouts so arrays can be handed to the C ABI without copies.
There is no Python or CPU fallback for the trace path: if ``libb200rt.so`` is missing the
import of :func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent

# --------------------------------------------------------------------------------------------
# numpy wire types (all little-endian, packed exactly like the reference structs)
VEC3 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
RAY = np.dtype([("origin", "<f4", 3), ("tMin", "<f4"), ("direction", "<f4", 3), ("tMax", "<f4")])
HIT_RECORD = np.dtype(
    [("r", RAY), ("primID", "<u4"), ("normal", "<f4", 3), ("throughput", "<f4", 3), ("geomID", "<u2"), ("flags", "<u2")]
)
TRACE_RESULT = np.dtype([("rgb", "<f4", 3), ("p", "<f4", 2), ("h", HIT_RECORD)])
BVH_NODE = np.dtype(
    [("min", "<f4", 3), ("primOrSecondChild", "<u4"), ("d", "<u2", 3), ("geomID", "<u2")]
)
TRIANGLE = np.dtype([("v", "<u2", 3)])
MESH_INFO = np.dtype([("firstIndex", "<u4"), ("firstVertex", "<u4"), ("numTriangles", "<u4"), ("numVertices", "<u4")])
GEOM_REF = np.dtype([("index", "<u2"), ("type", "u1"), ("pad", "u1")])
MATERIAL = np.dtype(
    [("albedo", "<f4", 3), ("ior", "<f4"), ("emission", "<f4", 3), ("type", "<i4"), ("emissive", "u1"), ("pad", "u1", 3)]
)
HIT = np.dtype([("t", "<f4"), ("geom_id", "<u4"), ("prim_id", "<u4"), ("normal", "<f4", 3)])

assert RAY.itemsize == 32 and HIT_RECORD.itemsize == 64 and TRACE_RESULT.itemsize == 84
assert BVH_NODE.itemsize == 24 and TRIANGLE.itemsize == 6 and MESH_INFO.itemsize == 16
assert GEOM_REF.itemsize == 4 and MATERIAL.itemsize == 36 and HIT.itemsize == 24

INVALID_GEOM = 0xFFFF
INVALID_PRIM = 0xFFFFFFFF
FLAG_ERROR = 1
FLAG_ESCAPED = 2


# --------------------------------------------------------------------------------------------
# C structs
class SceneDesc(C.Structure):
    _fields_ = [
        ("geometry", C.c_void_p), ("num_geometry", C.c_uint32),
        ("mesh_info", C.c_void_p), ("num_meshes", C.c_uint32),
        ("mesh_tris", C.c_void_p), ("num_tris", C.c_uint32),
        ("mesh_verts", C.c_void_p), ("num_verts", C.c_uint32),
        ("mesh_normals", C.c_void_p), ("num_normals", C.c_uint32),
        ("mat_ids", C.c_void_p), ("num_mat_ids", C.c_uint32),
        ("materials", C.c_void_p), ("num_materials", C.c_uint32),
        ("bvh_nodes", C.c_void_p), ("num_bvh_nodes", C.c_uint32),
        ("max_leaf_depth", C.c_uint32),
        ("spheres", C.c_void_p), ("num_spheres", C.c_uint32),
        ("discs", C.c_void_p), ("num_discs", C.c_uint32),
        ("image_width", C.c_float), ("image_height", C.c_float),
        ("fov_radians", C.c_float), ("anti_alias_scale", C.c_float),
        ("max_path_length", C.c_uint32), ("roulette_start_depth", C.c_uint32),
        ("samples_per_pixel", C.c_uint32), ("rng_seed", C.c_uint64),
        ("path_trace", C.c_int32), ("device", C.c_int32),
    ]


class TraceParams(C.Structure):
    _fields_ = [
        ("light_pos", C.c_float * 3), ("ambient", C.c_float),
        ("first_sample", C.c_uint32), ("num_samples", C.c_uint32),
        ("rays_per_batch", C.c_uint32), ("traversal", C.c_uint32),
        ("scene_residency", C.c_uint32), ("samples_per_chunk", C.c_uint32),
        ("count_visits", C.c_uint32), ("primary_pass", C.c_uint32),
        ("batch_stride", C.c_uint32), ("first_batch", C.c_uint32), ("tail_bounce", C.c_uint32), ("chunk_overlap", C.c_uint32),
    ]


class TraceStats(C.Structure):
    _fields_ = [
        ("closest_hit_queries", C.c_uint64), ("occlusion_queries", C.c_uint64),
        ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
        ("samples", C.c_uint64), ("escaped_samples", C.c_uint64),
        ("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("trace_secs", C.c_double), ("kernel_launches", C.c_uint64),
        ("trace_kernel_ms", C.c_double), ("nif_kernel_ms", C.c_double), ("accumulate_kernel_ms", C.c_double),
        ("trace_kernel_launches", C.c_uint64), ("nif_kernel_launches", C.c_uint64),
        ("shade_kernel_ms", C.c_double), ("shade_kernel_launches", C.c_uint64),
    ]


class NifLayer(C.Structure):
    _fields_ = [
        ("in_features", C.c_uint32), ("out_features", C.c_uint32),
        ("kernel_f16", C.c_void_p), ("bias_f16", C.c_void_p), ("relu", C.c_int32),
    ]


class NifDesc(C.Structure):
    _fields_ = [
        ("embedding_dimension", C.c_uint32), ("num_layers", C.c_uint32),
        ("layers", C.POINTER(NifLayer)),
        ("max", C.c_float), ("mean", C.c_float * 3), ("log_tone_map", C.c_int32),
    ]


class NifMetadata(C.Structure):
    _fields_ = [
        ("embedding_dimension", C.c_uint32), ("hidden_size", C.c_uint32),
        ("image_shape", C.c_uint32 * 3),
        ("max", C.c_float), ("eps", C.c_float), ("mean", C.c_float * 3), ("log_tone_map", C.c_int32),
    ]


class KerasLayer(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("activation", C.c_char * 32), ("layer", NifLayer)]


RAY_CALLBACK = C.CFUNCTYPE(None, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p)

# Every symbol include/b200rt.h declares; tests check the built library exports each one.
B200RT_SYMBOLS = [
    "b200rt_abi_version", "b200rt_last_error", "b200rt_device_count", "b200rt_device_ordinal",
    "b200rt_scene_create", "b200rt_scene_destroy",
    "b200rt_scene_load_nif", "b200rt_scene_set_hdri_rotation", "b200rt_scene_set_max_nif_batch_size",
    "b200rt_nif_eval", "b200rt_trace", "b200rt_trace_device",
    "b200rt_get_trace_stats", "b200rt_get_trace_time_secs",
    "b200rt_intersect", "b200rt_occluded", "b200rt_host_register", "b200rt_host_unregister",
]
B200RT_SCENE_SYMBOLS = [
    "b200rt_host_scene_builtin", "b200rt_host_scene_import", "b200rt_host_scene_free",
    "b200rt_host_scene_desc", "b200rt_scene_last_error", "b200rt_build_bvh",
    "b200rt_init_ray_stream", "b200rt_scale_rgb", "b200rt_visualise_hits",
    "b200rt_write_exr", "b200rt_write_pfm", "b200rt_read_nif_metadata", "b200rt_sincos",
    "b200rt_scene_desc_from_blob", "b200rt_scene_blob_write",
    "b200rt_keras_hdf5_open", "b200rt_keras_hdf5_close", "b200rt_keras_hdf5_num_layers", "b200rt_keras_hdf5_layer",
    "b200rt_keras_hdf5_version", "b200rt_keras_last_error",
]

_lib = None
_scene_lib = None


def lib_path() -> Path:
    # B200RT_LIB selects another build of the same ABI (`make experiments`: libb200rt_exp.so); measurements only
    override = os.environ.get("B200RT_LIB")
    return Path(override) if override else PKG_DIR / "libb200rt.so"


def scene_lib_path() -> Path:
    return PKG_DIR / "libb200rt_scene.so"


def lib() -> C.CDLL:
    """The CUDA trace library. Raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise RuntimeError(f"{p} is missing: build it with `make` (or __graft_entry__.build()); "
                               "the trace path has no CPU or PyTorch fallback")
        L = C.CDLL(str(p), mode=C.RTLD_GLOBAL if hasattr(C, "RTLD_GLOBAL") else 0)
        L.b200rt_abi_version.restype = C.c_int
        L.b200rt_last_error.restype = C.c_char_p
        L.b200rt_device_count.restype = C.c_int
        L.b200rt_device_ordinal.argtypes = [C.c_int]
        L.b200rt_device_ordinal.restype = C.c_int
        L.b200rt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
        L.b200rt_scene_destroy.argtypes = [C.c_void_p]
        L.b200rt_scene_destroy.restype = None
        L.b200rt_scene_load_nif.argtypes = [C.c_void_p, C.POINTER(NifDesc)]
        L.b200rt_scene_set_hdri_rotation.argtypes = [C.c_void_p, C.c_float]
        L.b200rt_scene_set_max_nif_batch_size.argtypes = [C.c_void_p, C.c_size_t]
        L.b200rt_nif_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.b200rt_trace.argtypes = [C.c_void_p, C.POINTER(TraceParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.b200rt_trace_device.argtypes = [C.c_void_p, C.POINTER(TraceParams), C.c_void_p, C.c_size_t, C.c_void_p]
        L.b200rt_get_trace_stats.argtypes = [C.c_void_p, C.POINTER(TraceStats)]
        L.b200rt_get_trace_time_secs.argtypes = [C.c_void_p]
        L.b200rt_get_trace_time_secs.restype = C.c_double
        L.b200rt_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]
        L.b200rt_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.b200rt_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.b200rt_host_unregister.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def scene_lib() -> C.CDLL:
    global _scene_lib
    if _scene_lib is None:
        p = scene_lib_path()
        if not p.exists():
            raise RuntimeError(f"{p} is missing: build it with `make` (or __graft_entry__.build())")
        L = C.CDLL(str(p))
        L.b200rt_scene_last_error.restype = C.c_char_p
        L.b200rt_host_scene_builtin.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.b200rt_host_scene_import.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.b200rt_host_scene_free.argtypes = [C.c_void_p]
        L.b200rt_host_scene_free.restype = None
        L.b200rt_host_scene_desc.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.b200rt_build_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.b200rt_init_ray_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.b200rt_scale_rgb.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        L.b200rt_scale_rgb.restype = None
        L.b200rt_visualise_hits.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(SceneDesc), C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.b200rt_visualise_hits.restype = C.c_long
        L.b200rt_write_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.b200rt_write_pfm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.b200rt_read_nif_metadata.argtypes = [C.c_char_p, C.POINTER(NifMetadata)]
        L.b200rt_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.b200rt_scene_desc_from_blob.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(SceneDesc)]
        L.b200rt_scene_blob_write.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_size_t]
        L.b200rt_scene_blob_write.restype = C.c_size_t
        L.b200rt_sincos.restype = None
        L.b200rt_keras_hdf5_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.b200rt_keras_hdf5_close.argtypes = [C.c_void_p]
        L.b200rt_keras_hdf5_close.restype = None
        L.b200rt_keras_hdf5_num_layers.argtypes = [C.c_void_p]
        L.b200rt_keras_hdf5_num_layers.restype = C.c_uint32
        L.b200rt_keras_hdf5_layer.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(KerasLayer)]
        L.b200rt_keras_hdf5_version.argtypes = [C.c_void_p]
        L.b200rt_keras_hdf5_version.restype = C.c_char_p
        L.b200rt_keras_last_error.restype = C.c_char_p
        _scene_lib = L
    return _scene_lib


def ptr(a: np.ndarray | None):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
