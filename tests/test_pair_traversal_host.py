"""The traversal core of the CUDA kernels (csrc/rt_prims.h: pair table, NaN-free fast slab test, near-first closest
hit, any hit), compiled for the HOST by g++ from the same header (tests/host_pair_check.cpp), against the oracle.
Runs without a GPU. Bit-exact: t, ids and normals."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import CustomScene, make_rays
from ipu_ray_lib_b200 import _capi as capi, scene

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "tests" / "libhostpair.so"


@pytest.fixture(scope="module")
def hostpair():
    src = ROOT / "tests" / "host_pair_check.cpp"
    hdrs = list((ROOT / "ipu_ray_lib_b200" / "csrc").glob("*.h*"))
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in [src, *hdrs]):
        subprocess.run(["make", "-C", str(ROOT), "tests/libhostpair.so"], check=True, capture_output=True)
    L = C.CDLL(str(LIB))
    L.hostpair_last_error.restype = C.c_char_p
    L.hostpair_intersect.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.hostpair_stream_intersect.argtypes = L.hostpair_intersect.argtypes
    L.hostpair_occluded.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p]
    L.hostpair_validate.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    return L


# "pair": pair_closest_hit (shadow / bare-query kernels); "stream": the step functions of wf_trace_kernel
# (stream_begin / stream_trav / stream_leaf / stream_pop; window tMin = 0, tMax = inf as in the bounce loop)
KINDS = ["pair", "stream"]


def run_intersect(L, s, rays, kind="pair"):
    out = np.zeros(rays.size, dtype=capi.HIT)
    cnt = np.zeros(4, dtype=np.uint64)
    fn = L.hostpair_intersect if kind == "pair" else L.hostpair_stream_intersect
    assert fn(C.byref(s.desc), capi.ptr(rays), rays.size, capi.ptr(out), capi.ptr(cnt)) == 0, L.hostpair_last_error()
    return out, cnt


def root_box(s):
    mn = s.bvh_nodes["min"][0].astype(np.float32)
    ext = np.frombuffer(s.bvh_nodes["d"][0].tobytes(), dtype=np.float16).astype(np.float32)
    return mn, ext


def random_rays(n, seed, lo, hi, unit=True):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    if unit:
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    return make_rays(o, d.astype(np.float32))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", ["box", "spheres", "box-simple"])
def test_closest_hit_matches_oracle_on_random_unit_rays(hostpair, port, name, kind):
    s = scene.HostScene.builtin(name)
    mn, ext = root_box(s)
    rays = random_rays(60000, 7, mn - 0.1 * ext, mn + 1.1 * ext)
    got, cnt = run_intersect(hostpair, s, rays, kind)
    want, _ = port.intersect(s, rays)
    assert got.tobytes() == want.tobytes()
    assert cnt[2] == rays.size  # every one of these queries is NaN-free: the fast slab test was used
    assert (got["geom_id"] != capi.INVALID_GEOM).mean() > 0.1


@pytest.mark.parametrize("kind", KINDS)
def test_camera_and_axis_aligned_rays(hostpair, port, box_scene, kind):
    """Camera rays of the shadow-trace configuration include exact zeros in the direction (the centre column/row):
    those queries must take the NaN-preserving slab test and still agree bit for bit."""
    w = h = 128
    box_scene.configure(w, h, path_trace=False)
    stream = scene.init_ray_stream(w, h, box_scene.fov)
    rays = make_rays(stream["h"]["r"]["origin"], stream["h"]["r"]["direction"])
    assert (rays["direction"] == 0).any()
    axis = make_rays(np.tile(np.array([[0, 0, 0], [10, 20, -300], [0, 0, -500]], np.float32), (6, 1))[:18],
                     np.repeat(np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32), 3, axis=0))
    rays = np.concatenate([rays, axis])
    got, cnt = run_intersect(hostpair, box_scene, rays, kind)
    want, _ = port.intersect(box_scene, rays)
    assert got.tobytes() == want.tobytes()
    assert 0 < cnt[2] < rays.size


@pytest.mark.parametrize("kind", KINDS)
def test_rays_starting_on_node_planes_with_zero_direction_components(hostpair, port, box_scene, kind):
    """(bound - origin) * (1/0) = NaN: the case the fast slab test must never see."""
    nodes = box_scene.bvh_nodes
    rng = np.random.default_rng(3)
    pick = rng.integers(0, nodes.size, 4000)
    o = nodes["min"][pick].astype(np.float32).copy()
    d = rng.normal(size=(pick.size, 3)).astype(np.float32)
    d[np.arange(pick.size), rng.integers(0, 3, pick.size)] = 0.0
    d[::7] *= -0.0  # negative zeros too
    rays = make_rays(o, d)
    got, cnt = run_intersect(hostpair, box_scene, rays, kind)
    want, _ = port.intersect(box_scene, rays)
    assert cnt[2] == 0
    # These rays are built to graze: they start ON a box corner, so some hit a shared triangle edge with bit-equal t in
    # both triangles. Near-first order may then keep the other triangle of the pair (DESIGN.md "Traversal-order
    # caveat"); t itself must still agree everywhere, and anything but such a tie must agree in every field.
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    differ = np.nonzero((got["prim_id"] != want["prim_id"]) | (got["geom_id"] != want["geom_id"]))[0]
    assert differ.size <= rays.size // 1000
    same = np.setdiff1d(np.arange(rays.size), differ)
    assert got[same].tobytes() == want[same].tobytes()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("fixture", ["dae_scene", "hdri_scene"])
def test_imported_scenes(hostpair, port, fixture, request, kind):
    s = request.getfixturevalue(fixture)
    mn, ext = root_box(s)
    rays = random_rays(40000, 11, mn - 0.2 * ext, mn + 1.2 * ext)
    got, _ = run_intersect(hostpair, s, rays, kind)
    want, _ = port.intersect(s, rays)
    assert got.tobytes() == want.tobytes()


def test_any_hit_matches_oracle(hostpair, port, box_scene):
    mn, ext = root_box(box_scene)
    rays = random_rays(50000, 5, mn - 0.1 * ext, mn + 1.1 * ext)
    rays["tMax"] = np.random.default_rng(1).uniform(10, 900, rays.size).astype(np.float32)
    out = np.zeros(rays.size, np.uint8)
    assert hostpair.hostpair_occluded(C.byref(box_scene.desc), capi.ptr(rays), rays.size, capi.ptr(out)) == 0
    want = port.occluded(box_scene, rays)
    assert np.array_equal(out, want) and 0.1 < out.mean() < 0.9


@pytest.mark.parametrize("kind", KINDS)
def test_single_leaf_tree_and_equal_t_ties(hostpair, port, kind):
    one = CustomScene(spheres=[(0, 0, -5, 1)])
    rays = make_rays([[0, 0, 0], [0, 3, 0]], [[0, 0, -1], [0, 0, -1]])
    got, _ = run_intersect(hostpair, one, rays, kind)
    want, _ = port.intersect(one, rays)
    assert got.tobytes() == want.tobytes() and got["geom_id"][0] == 0
    # two coincident quads: every hit is an exact tie between two leaves
    q = np.array([[-1, -1, -4], [1, -1, -4], [1, 1, -4], [-1, 1, -4]], np.float32)
    tie = CustomScene(meshes=[(q, [[0, 1, 2], [0, 2, 3]]), (q, [[0, 1, 2], [0, 2, 3]])])
    rng = np.random.default_rng(0)
    rays = make_rays(np.zeros((3000, 3), np.float32), np.c_[rng.uniform(-.3, .3, (3000, 2)), -np.ones(3000)].astype(np.float32))
    got, _ = run_intersect(hostpair, tie, rays, kind)
    want, _ = port.intersect(tie, rays)
    assert got.tobytes() == want.tobytes() and (got["geom_id"] != capi.INVALID_GEOM).mean() > 0.5


def test_malformed_node_arrays_are_rejected(hostpair, box_scene):
    """What b200rt_scene_create validates before uploading anything (serialised scenes come from outside)."""
    def check(mutate, expect):
        nodes = box_scene.bvh_nodes.copy()
        mutate(nodes)
        d = capi.SceneDesc.from_buffer_copy(box_scene.desc)
        d.bvh_nodes = nodes.ctypes.data
        assert hostpair.hostpair_validate(C.byref(d), None, None, None) == -1
        assert expect in hostpair.hostpair_last_error().decode()

    inner = int(np.nonzero(box_scene.bvh_nodes["geomID"] == 0xFFFF)[0][5])
    leaf = int(np.nonzero(box_scene.bvh_nodes["geomID"] != 0xFFFF)[0][5])
    check(lambda n: n["primOrSecondChild"].__setitem__(inner, n.size + 3), "child index out of range")
    check(lambda n: n["primOrSecondChild"].__setitem__(inner, inner), "second child")
    check(lambda n: n["geomID"].__setitem__(leaf, 200), "geomID out of range")
    check(lambda n: n["primOrSecondChild"].__setitem__(leaf, 1 << 20), "primID outside")
    n_pairs, depth, finite = C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert hostpair.hostpair_validate(C.byref(box_scene.desc), C.byref(n_pairs), C.byref(depth), C.byref(finite)) == 0
    assert n_pairs.value == (box_scene.bvh_nodes.size - 1) // 2 and finite.value == 1
    assert depth.value <= box_scene.desc.max_leaf_depth
