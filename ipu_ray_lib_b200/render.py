"""Python mirror of the reference's device orchestration object for the trace path.

``B200Scene`` has the public surface of ``IpuScene`` (include/IpuScene.hpp:31-56): constructed
from a scene + ray stream, optional NIF model / HDRI rotation, ``execute()`` renders the stream in
place, ``get_trace_time_secs()`` reports the timed span. Everything runs through the C ABI of
``include/b200rt.h``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .scene import HostScene


class B200Error(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc != 0:
        raise B200Error(f"b200rt error {rc}: {capi.lib().b200rt_last_error().decode()}")


class B200Scene:
    """One replica of the trace path bound to one GPU (IpuScene, src/IpuScene.cpp:24-62)."""

    def __init__(self, scene: HostScene, ray_callback=None, rays_per_worker: int = 1):
        self.host_scene = scene
        self.ray_callback = ray_callback
        self.rays_per_worker = rays_per_worker
        self._handle = C.c_void_p()
        _check(capi.lib().b200rt_scene_create(C.byref(scene.desc), C.byref(self._handle)))
        self._nif_keepalive = None

    def close(self) -> None:
        h, self._handle = self._handle, C.c_void_p()
        if h:
            capi.lib().b200rt_scene_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- NIF environment light (IpuScene::loadNifModel & friends) -----------------------------
    def load_nif_model(self, model) -> bool:
        """``model`` is an :class:`ipu_ray_lib_b200.nif.NifWeights`."""
        desc, keep = model.to_desc()
        self._nif_keepalive = keep
        _check(capi.lib().b200rt_scene_load_nif(self._handle, C.byref(desc)))
        return True

    def set_hdri_rotation(self, degrees: float) -> None:
        _check(capi.lib().b200rt_scene_set_hdri_rotation(self._handle, degrees))

    def set_max_nif_batch_size(self, rays_per_batch: int) -> None:
        _check(capi.lib().b200rt_scene_set_max_nif_batch_size(self._handle, rays_per_batch))

    def set_available_memory_proportion(self, proportion: float) -> None:
        """Accepted for interface parity (IpuScene::setAvailableMemoryProportion); no effect on B200."""

    def nif_eval(self, uv: np.ndarray) -> np.ndarray:
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), dtype=np.float32)
        _check(capi.lib().b200rt_nif_eval(self._handle, capi.ptr(uv), uv.shape[0], capi.ptr(out)))
        return out

    # -- tracing -----------------------------------------------------------------------------
    @staticmethod
    def make_params(**kw) -> capi.TraceParams:
        p = capi.TraceParams()
        for k, v in kw.items():
            if k == "light_pos":
                p.light_pos[0], p.light_pos[1], p.light_pos[2] = v
            else:
                setattr(p, k, v)
        return p

    def execute(self, rays: np.ndarray, **params) -> np.ndarray:
        """Render a host-resident TraceResult stream in place (IpuScene::execute)."""
        assert rays.dtype == capi.TRACE_RESULT and rays.flags["C_CONTIGUOUS"]
        p = self.make_params(**params)
        if "rays_per_batch" not in params:
            p.rays_per_batch = 8640 * self.rays_per_worker
        cb = None
        if self.ray_callback is not None:
            user_cb = self.ray_callback

            def _trampoline(idx, ptr, n, _user):
                buf = (C.c_char * (n * capi.TRACE_RESULT.itemsize)).from_address(ptr)
                user_cb(idx, np.frombuffer(buf, dtype=capi.TRACE_RESULT, count=n))

            cb = capi.RAY_CALLBACK(_trampoline)
        _check(capi.lib().b200rt_trace(self._handle, C.byref(p), capi.ptr(rays), rays.size,
                                       C.cast(cb, C.c_void_p) if cb else None, None))
        return rays

    def execute_device(self, device_ptr: int, num_rays: int, stream: int = 0, **params) -> None:
        """Render a TraceResult stream already resident in HBM (e.g. a torch uint8 tensor's data_ptr) on `stream`
        (a cudaStream_t handle; 0 = the legacy default stream, i.e. torch's default stream)."""
        p = self.make_params(**params)
        _check(capi.lib().b200rt_trace_device(self._handle, C.byref(p), C.c_void_p(device_ptr), num_rays,
                                              C.c_void_p(stream) if stream else None))

    def intersect(self, rays: np.ndarray, traversal: int = 0) -> np.ndarray:
        assert rays.dtype == capi.RAY
        out = np.zeros(rays.size, dtype=capi.HIT)
        _check(capi.lib().b200rt_intersect(self._handle, capi.ptr(rays), rays.size, capi.ptr(out), traversal))
        return out

    def occluded(self, rays: np.ndarray) -> np.ndarray:
        assert rays.dtype == capi.RAY
        out = np.zeros(rays.size, dtype=np.uint8)
        _check(capi.lib().b200rt_occluded(self._handle, capi.ptr(rays), rays.size, capi.ptr(out)))
        return out

    def stats(self) -> dict:
        s = capi.TraceStats()
        _check(capi.lib().b200rt_get_trace_stats(self._handle, C.byref(s)))
        return {k: getattr(s, k) for k, _ in capi.TraceStats._fields_ if k != "reserved"}

    def get_trace_time_secs(self) -> float:
        return capi.lib().b200rt_get_trace_time_secs(self._handle)
