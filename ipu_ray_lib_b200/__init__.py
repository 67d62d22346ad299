"""B200-native ray-data-parallel trace path of markp-gc/ipu_ray_lib.

Only what the hot path needs lives here: ``csrc/`` (hand-written sm_100a CUDA kernels and the
C ABI of ``include/b200rt.h``), ``host/`` (C++ mirror of the reference's IpuScene / trace CLI and
the host-side scene utilities) and thin ctypes views used by tests and ``bench.py``.
"""
from . import _capi  # noqa: F401
from .scene import HostScene, init_ray_stream, scale_rgb, visualise_hits  # noqa: F401

__all__ = ["HostScene", "init_ray_stream", "scale_rgb", "visualise_hits"]
