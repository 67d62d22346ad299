"""Host-side scene handling: thin Python view over ``libb200rt_scene.so``.

Mirrors the reference's ``buildSceneDescription`` + ``buildSceneData`` pair
(src/app_utils.cpp:252-364): pick a built-in scene or import a file, flatten it into the unified
arrays and build the compact BVH. The arrays are exposed as zero-copy numpy views.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np

from . import _capi as capi

ASSETS = capi.REPO_ROOT / "assets"
DEFAULT_MESH = ASSETS / "monkey_bust.glb"

VISUALISE_MODES = {"rgb": 0, "id": 1, "normal": 2, "tfar": 3, "color": 4, "hitpoint": 5}


def _view(addr, count, dtype):
    if not addr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class HostScene:
    """Owns the flattened scene arrays + BVH; ``desc`` is ready for ``b200rt_scene_create``."""

    def __init__(self, handle):
        self._h = handle
        self.desc = capi.SceneDesc()
        rc = capi.scene_lib().b200rt_host_scene_desc(self._h, C.byref(self.desc))
        if rc != 0:
            raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
        d = self.desc
        self.geometry = _view(d.geometry, d.num_geometry, capi.GEOM_REF)
        self.mesh_info = _view(d.mesh_info, d.num_meshes, capi.MESH_INFO)
        self.mesh_tris = _view(d.mesh_tris, d.num_tris, capi.TRIANGLE)
        self.mesh_verts = _view(d.mesh_verts, d.num_verts, capi.VEC3)
        self.mesh_normals = _view(d.mesh_normals, d.num_normals, capi.VEC3)
        self.mat_ids = _view(d.mat_ids, d.num_mat_ids, np.dtype("<u4"))
        self.materials = _view(d.materials, d.num_materials, capi.MATERIAL)
        self.bvh_nodes = _view(d.bvh_nodes, d.num_bvh_nodes, capi.BVH_NODE)
        self.spheres = _view(d.spheres, d.num_spheres * 4, np.dtype("<f4")).reshape(-1, 4)
        self.discs = _view(d.discs, d.num_discs * 7, np.dtype("<f4")).reshape(-1, 7)

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def builtin(cls, name: str = "box", mesh_file: str | Path | None = None) -> "HostScene":
        """``--scene box|box-simple|spheres`` (trace.cpp:360)."""
        mesh = str(mesh_file if mesh_file is not None else DEFAULT_MESH).encode()
        h = C.c_void_p()
        rc = capi.scene_lib().b200rt_host_scene_builtin(name.encode(), mesh, C.byref(h))
        if rc != 0:
            raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
        return cls(h)

    @classmethod
    def from_file(cls, path: str | Path, load_normals: bool = False) -> "HostScene":
        """``--mesh-file`` (trace.cpp:349-353)."""
        h = C.c_void_p()
        rc = capi.scene_lib().b200rt_host_scene_import(str(path).encode(), int(load_normals), C.byref(h))
        if rc != 0:
            raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
        return cls(h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                capi.scene_lib().b200rt_host_scene_free(h)
            except Exception:
                pass

    # -- render parameters (SceneRef scalars, trace.cpp:468-488) -----------------------------
    def configure(self, width: int, height: int, *, path_trace: bool = True, samples: int = 256,
                  seed: int = 1442, anti_alias: float = 0.25, max_path_length: int = 10,
                  roulette_start_depth: int = 3, device: int = -1) -> "HostScene":
        d = self.desc
        d.image_width, d.image_height = float(width), float(height)
        d.anti_alias_scale = anti_alias
        d.max_path_length = max_path_length
        d.roulette_start_depth = roulette_start_depth
        d.samples_per_pixel = samples
        d.rng_seed = seed
        d.path_trace = int(path_trace)
        d.device = device
        return self

    @property
    def fov(self) -> float:
        return self.desc.fov_radians

    def stats(self) -> dict:
        d = self.desc
        return {"meshes": d.num_meshes, "triangles": d.num_tris, "vertices": d.num_verts,
                "spheres": d.num_spheres, "discs": d.num_discs, "bvh_nodes": d.num_bvh_nodes,
                "bvh_bytes": d.num_bvh_nodes * 24, "max_leaf_depth": d.max_leaf_depth}


class BlobScene:
    """A scene described by the reference's serialised ``SceneRef`` byte stream (zero copy: ``desc`` points into the
    blob). Spheres / discs travel beside the stream, as in the reference (src/IpuScene.cpp:208-215)."""

    def __init__(self, blob: np.ndarray, spheres=None, discs=None, *, path_trace: bool = True, seed: int = 1442,
                 device: int = -1):
        raw = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob.view(np.uint8).ravel()
        self._store = np.zeros(raw.size + 16, dtype=np.uint8)  # own 16-byte aligned copy
        off = (-self._store.ctypes.data) % 16
        self.blob = self._store[off:off + raw.size]
        self.blob[:] = raw
        self.desc = capi.SceneDesc()
        rc = capi.scene_lib().b200rt_scene_desc_from_blob(capi.ptr(self.blob), self.blob.size, C.byref(self.desc))
        if rc != 0:
            raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
        self.spheres = np.ascontiguousarray(spheres if spheres is not None else np.zeros((0, 4)), dtype=np.float32)
        self.discs = np.ascontiguousarray(discs if discs is not None else np.zeros((0, 7)), dtype=np.float32)
        d = self.desc
        d.spheres = self.spheres.ctypes.data if self.spheres.size else None
        d.num_spheres = self.spheres.shape[0]
        d.discs = self.discs.ctypes.data if self.discs.size else None
        d.num_discs = self.discs.shape[0]
        d.path_trace, d.rng_seed, d.device = int(path_trace), seed, device

    @property
    def fov(self) -> float:
        return self.desc.fov_radians


def scene_blob(scene) -> np.ndarray:
    """Serialise ``scene.desc`` the way the reference's ``Serialiser<16>`` does."""
    lib = capi.scene_lib()
    n = lib.b200rt_scene_blob_write(C.byref(scene.desc), None, 0)
    out = np.zeros(n, dtype=np.uint8)
    lib.b200rt_scene_blob_write(C.byref(scene.desc), capi.ptr(out), n)
    return out


def init_ray_stream(width: int, height: int, fov: float, window=None) -> np.ndarray:
    """initPerspectiveRayStream + zeroRgb (src/app_utils.cpp:19-53); window = (w, h, col, row)."""
    w, h, c, r = window if window is not None else (width, height, 0, 0)
    rays = np.zeros(w * h, dtype=capi.TRACE_RESULT)
    rc = capi.scene_lib().b200rt_init_ray_stream(capi.ptr(rays), width, height, w, h, c, r, fov)
    if rc != 0:
        raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
    return rays


def scale_rgb(rays: np.ndarray, scale: float) -> None:
    capi.scene_lib().b200rt_scale_rgb(capi.ptr(rays), rays.size, scale)


def visualise_hits(rays: np.ndarray, scene: HostScene, mode: str, width: int, height: int):
    """visualiseHits (src/app_utils.cpp:61-127): returns (H x W x 3 BGR float image, hit count)."""
    img = np.zeros((height, width, 3), dtype=np.float32)
    hits = capi.scene_lib().b200rt_visualise_hits(capi.ptr(rays), rays.size, C.byref(scene.desc),
                                                  VISUALISE_MODES[mode], capi.ptr(img), width, height)
    if hits < 0:
        raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
    return img, hits


def write_exr(path, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, dtype=np.float32)
    rc = capi.scene_lib().b200rt_write_exr(str(path).encode(), capi.ptr(img), img.shape[1], img.shape[0])
    if rc != 0:
        raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())


def write_pfm(path, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, dtype=np.float32)
    rc = capi.scene_lib().b200rt_write_pfm(str(path).encode(), capi.ptr(img), img.shape[1], img.shape[0])
    if rc != 0:
        raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())


def host_sincos(x: float):
    s, c = C.c_float(), C.c_float()
    capi.scene_lib().b200rt_sincos(x, C.byref(s), C.byref(c))
    return s.value, c.value
