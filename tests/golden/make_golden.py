"""Generates tests/golden/reference_vectors.npz from the REFERENCE build of the oracle.

Run in the build container (needs oracle/_ref/liboracle_ref.so, i.e. /root/reference):
    python tests/golden/make_golden.py
The vectors are outputs of the reference's own kernel sources (src/Mesh.cpp, src/Primitives.cpp,
src/CompactBVH2Node.cpp, ext/math/sincos.cpp and the headers they include) driven by
oracle/ref_driver.cpp; they pin oracle/oracle_port.cpp and the CUDA path on machines where the
reference tree does not exist (the GPU box).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from ipu_ray_lib_b200 import _capi as capi, scene  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402


def edge_case_rays(rng, n_random, bounds_lo, bounds_hi):
    """Random rays through the scene volume plus the edge cases the slab/primitive tests care about."""
    rays = np.zeros(n_random + 64, dtype=capi.RAY)
    o = rng.uniform(bounds_lo, bounds_hi, (rays.size, 3)).astype(np.float32)
    d = rng.standard_normal((rays.size, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # axis-parallel directions (zero components -> inf/NaN slab arithmetic)
    for k in range(24):
        d[n_random + k] = 0
        d[n_random + k, k % 3] = 1.0 if (k // 3) % 2 == 0 else -1.0
    # one zero component
    for k in range(24, 48):
        d[n_random + k, k % 3] = 0.0
        d[n_random + k] /= np.linalg.norm(d[n_random + k])
    # negative zero component, un-normalised directions
    d[n_random + 48:n_random + 56, 1] = -0.0
    d[n_random + 56:] *= 3.5
    rays["origin"] = o
    rays["direction"] = d.astype(np.float32)
    rays["tMin"] = 0.0
    rays["tMax"] = np.inf
    rays["tMax"][::7] = 400.0  # finite tMax on some
    return rays


def main():
    ref = Oracle("reference")
    out = {}
    rng = np.random.default_rng(20221017)
    for name, (w, h) in {"box": (40, 40), "spheres": (32, 32)}.items():
        s = scene.HostScene.builtin(name)
        out[f"{name}_bvh_sha1"] = np.frombuffer(hashlib.sha1(s.bvh_nodes.tobytes()).digest(), dtype=np.uint8)
        s.configure(w, h, path_trace=False)
        rays = scene.init_ray_stream(w, h, s.fov)
        ref.shadow_trace(s, rays, threads=1)
        out[f"{name}_shadow"] = rays.view(np.uint8).copy()
        s.configure(w, h, path_trace=True, samples=5)
        rays = scene.init_ray_stream(w, h, s.fov)
        counters = ref.path_trace(s, rays, threads=1)
        out[f"{name}_path"] = rays.view(np.uint8).copy()
        out[f"{name}_path_counters"] = np.array([counters["closest_hit_queries"], counters["prim_tests"],
                                                 counters["samples"], counters["escaped_samples"]], dtype=np.uint64)
        lo, hi = ((-300, -300, -1400), (300, 300, -700)) if name == "box" else ((-4, -2, -8), (4, 3, 0))
        q = edge_case_rays(rng, 448, lo, hi)
        hits, _ = ref.intersect(s, q, threads=1)
        out[f"{name}_query_rays"] = q.view(np.uint8).copy()
        out[f"{name}_query_hits"] = hits.view(np.uint8).copy()
        out[f"{name}_query_occluded"] = ref.occluded(s, q, threads=1)

    x = np.concatenate([rng.uniform(-50, 50, 200), [0.0, 0.7, -0.7, np.pi, 2 * np.pi, 1e-8, 100.5]]).astype(np.float32)
    s_, c_ = ref.sincos(x)
    out["sincos_x"], out["sincos_s"], out["sincos_c"] = x, s_, c_
    out["uniform_1442"] = ref.uniform_stream(1442, 64)
    out["raw_1442"] = ref.raw_stream(1442, 64)
    nrm = rng.standard_normal((64, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[0] = (0, 0, 1); nrm[1] = (1, 0, 0); nrm[2] = (0, -1, 0)
    u12 = rng.uniform(0, 1, (64, 2)).astype(np.float32)
    u12[0] = (0.3, 0.7); u12[1] = (0.5, 0.5); u12[2] = (0.0, 1.0)
    out["diffuse_normals"], out["diffuse_u"], out["diffuse_out"] = nrm, u12, ref.sample_diffuse(nrm, u12)
    dirs = rng.standard_normal((64, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    ior_u1 = np.stack([np.full(64, 1.52, np.float32), rng.uniform(0, 1, 64).astype(np.float32)], axis=1)
    d_out, refr = ref.dielectric(dirs, nrm, ior_u1)
    out["dielectric_dirs"], out["dielectric_ior_u1"], out["dielectric_out"], out["dielectric_refracted"] = dirs, ior_u1, d_out, refr
    out["reflect_out"] = ref.reflect(dirs, nrm)
    org = rng.uniform(-500, 500, (64, 3)).astype(np.float32)
    out["offset_origins"], out["offset_out"] = org, ref.offset_ray(org, dirs, nrm)
    xy = rng.uniform(0, 1440, (64, 2)).astype(np.float32)
    out["p2r_xy"], out["p2r_out"] = xy, ref.pixel_to_ray_dir(xy, 1440.0, 1440.0, 0.41421357)
    hx = np.concatenate([rng.uniform(0, 700, 120), [0.1, 65504.0, 1e-6, 0.0]]).astype(np.float32)
    out["half_x"], out["half_out"] = hx, ref.round_to_half_not_smaller(hx)
    rcs = np.stack([rng.integers(0, 1440, 64), rng.integers(0, 1440, 64), rng.integers(0, 1000, 64)], axis=1).astype(np.uint32)
    out["camera_rcs"], out["camera_out"] = rcs, ref.camera_sample(1442, 1440, 1440, 0.7853982, 0.25, rcs)
    path = Path(__file__).with_name("reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
