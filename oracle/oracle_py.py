"""ctypes loader for the two CPU checkers (oracle/oracle_api.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from ipu_ray_lib_b200 import _capi as capi

ORACLE_DIR = Path(__file__).resolve().parent
PORT_LIB = ORACLE_DIR / "liboracle_port.so"
REF_LIB = ORACLE_DIR / "_ref" / "liboracle_ref.so"


class Oracle:
    """One of the two checkers. kind = 'port' (restatement) or 'reference' (reference sources)."""

    def __init__(self, kind: str):
        self.kind = kind
        self.prefix = "orc_" if kind == "port" else "ref_"
        path = PORT_LIB if kind == "port" else REF_LIB
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (make -C oracle)")
        self.lib = C.CDLL(str(path))
        f = self._f
        f("kind").restype = C.c_char_p
        assert f("kind")().decode() == kind
        f("shadow_trace").argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_float,
                                      C.c_int, C.c_void_p]
        f("path_trace").argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                    C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        f("intersect").argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        f("occluded").argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        f("sincos").argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        f("uniform_stream").argtypes = [C.c_uint64, C.c_size_t, C.c_void_p]
        f("raw_stream").argtypes = [C.c_uint64, C.c_size_t, C.c_void_p]
        f("sample_diffuse").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f("dielectric").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        f("reflect").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f("offset_ray").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f("pixel_to_ray_dir").argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_void_p]
        f("round_to_half_not_smaller").argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        f("camera_sample").argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p,
                                       C.c_size_t, C.c_void_p]
        f("nif_eval").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        f("dir_to_uv").argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- renders ----
    def shadow_trace(self, scene, rays, light=(18.0, 257.0, -1060.0), ambient=0.05, threads=0):
        counters = np.zeros(8, dtype=np.uint64)
        lp = np.asarray(light, dtype=np.float32)
        rc = self._f("shadow_trace")(C.byref(scene.desc), capi.ptr(rays), rays.size, capi.ptr(lp), ambient, threads,
                                     capi.ptr(counters))
        assert rc == 0
        return self._counters(counters)

    def path_trace(self, scene, rays, first_sample=0, num_samples=None, nif=None, hdri_rotation=0.0, threads=0):
        counters = np.zeros(8, dtype=np.uint64)
        n = scene.desc.samples_per_pixel if num_samples is None else num_samples
        nif_ptr, keep = None, None
        if nif is not None:
            d, keep = nif.to_desc()
            nif_ptr = C.cast(C.pointer(d), C.c_void_p)
        rc = self._f("path_trace")(C.byref(scene.desc), capi.ptr(rays), rays.size, first_sample, n, nif_ptr,
                                   hdri_rotation, threads, capi.ptr(counters))
        assert rc == 0
        return self._counters(counters)

    def path_trace_as_shipped(self, scene, rays, num_samples, threads=0):
        """Reference build only: renderCPU's loop as shipped (one global generator, draws inside `omp critical`, serial
        camera-ray regeneration per sample; trace.cpp:236-245). Not reproducible; for timing."""
        assert self.kind == "reference"
        fn = self.lib.ref_path_trace_as_shipped
        fn.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
        counters = np.zeros(8, dtype=np.uint64)
        assert fn(C.byref(scene.desc), capi.ptr(rays), rays.size, num_samples, threads, capi.ptr(counters)) == 0
        return self._counters(counters)

    def intersect(self, scene, rays, threads=0):
        out = np.zeros(rays.size, dtype=capi.HIT)
        counters = np.zeros(8, dtype=np.uint64)
        assert self._f("intersect")(C.byref(scene.desc), capi.ptr(rays), rays.size, capi.ptr(out), threads,
                                    capi.ptr(counters)) == 0
        return out, self._counters(counters)

    def occluded(self, scene, rays, threads=0):
        out = np.zeros(rays.size, dtype=np.uint8)
        assert self._f("occluded")(C.byref(scene.desc), capi.ptr(rays), rays.size, capi.ptr(out), threads) == 0
        return out

    @staticmethod
    def _counters(c):
        return {"closest_hit_queries": int(c[0]), "occlusion_queries": int(c[1]), "node_visits": int(c[2]),
                "prim_tests": int(c[3]), "samples": int(c[4]), "escaped_samples": int(c[5])}

    # ---- known-answer helpers ----
    def sincos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        s, c = np.zeros_like(x), np.zeros_like(x)
        self._f("sincos")(capi.ptr(x), x.size, capi.ptr(s), capi.ptr(c))
        return s, c

    def uniform_stream(self, seed, n):
        out = np.zeros(n, dtype=np.float32)
        self._f("uniform_stream")(seed, n, capi.ptr(out))
        return out

    def raw_stream(self, seed, n):
        out = np.zeros(n, dtype=np.uint64)
        self._f("raw_stream")(seed, n, capi.ptr(out))
        return out

    def sample_diffuse(self, normals, u12):
        normals = np.ascontiguousarray(normals, np.float32); u12 = np.ascontiguousarray(u12, np.float32)
        out = np.zeros_like(normals)
        self._f("sample_diffuse")(capi.ptr(normals), capi.ptr(u12), normals.shape[0], capi.ptr(out))
        return out

    def dielectric(self, dirs, normals, ior_u1):
        dirs = np.ascontiguousarray(dirs, np.float32); normals = np.ascontiguousarray(normals, np.float32)
        ior_u1 = np.ascontiguousarray(ior_u1, np.float32)
        out = np.zeros_like(dirs); refr = np.zeros(dirs.shape[0], np.uint8)
        self._f("dielectric")(capi.ptr(dirs), capi.ptr(normals), capi.ptr(ior_u1), dirs.shape[0], capi.ptr(out),
                              capi.ptr(refr))
        return out, refr

    def reflect(self, dirs, normals):
        dirs = np.ascontiguousarray(dirs, np.float32); normals = np.ascontiguousarray(normals, np.float32)
        out = np.zeros_like(dirs)
        self._f("reflect")(capi.ptr(dirs), capi.ptr(normals), dirs.shape[0], capi.ptr(out))
        return out

    def offset_ray(self, origins, dirs, normals):
        a = [np.ascontiguousarray(x, np.float32) for x in (origins, dirs, normals)]
        out = np.zeros_like(a[0])
        self._f("offset_ray")(capi.ptr(a[0]), capi.ptr(a[1]), capi.ptr(a[2]), a[0].shape[0], capi.ptr(out))
        return out

    def pixel_to_ray_dir(self, xy, w, h, tan_theta):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros((xy.shape[0], 3), np.float32)
        self._f("pixel_to_ray_dir")(capi.ptr(xy), xy.shape[0], w, h, tan_theta, capi.ptr(out))
        return out

    def round_to_half_not_smaller(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(x.size, np.uint16)
        self._f("round_to_half_not_smaller")(capi.ptr(x), x.size, capi.ptr(out))
        return out

    def camera_sample(self, seed, w, h, fov, aa, row_col_sample):
        rcs = np.ascontiguousarray(row_col_sample, np.uint32)
        out = np.zeros((rcs.shape[0], 3), np.float32)
        self._f("camera_sample")(seed, w, h, fov, aa, capi.ptr(rcs), rcs.shape[0], capi.ptr(out))
        return out

    def nif_eval(self, nif, uv, threads=0):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        d, keep = nif.to_desc()
        rc = self._f("nif_eval")(C.byref(d), capi.ptr(uv), uv.shape[0], capi.ptr(out), threads)
        if rc != 0:
            raise RuntimeError(f"{self.kind} oracle has no NIF")
        return out

    def nif_eval_partials(self, nif, uv, half_chunk=16, threads=0):
        """NIF forward with fp16 partials modelled (the IPU's partialsType=half, src/IpuScene.cpp:256-262). Port only."""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        d, keep = nif.to_desc()
        f = self._f("nif_eval_partials")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        assert f(C.byref(d), capi.ptr(uv), uv.shape[0], capi.ptr(out), threads, half_chunk) == 0
        return out

    def dir_to_uv(self, dirs, rotation_radians=0.0):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros((dirs.shape[0], 2), np.float32)
        self._f("dir_to_uv")(capi.ptr(dirs), dirs.shape[0], rotation_radians, capi.ptr(out))
        return out


def have_ref() -> bool:
    return REF_LIB.exists()


def have_port() -> bool:
    return PORT_LIB.exists()
