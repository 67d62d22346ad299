"""Host-side producers of the trace path's inputs: built-in scenes, GLB import, BVH builder, ray-gen, AOVs."""
import numpy as np
import pytest

from ipu_ray_lib_b200 import _capi as capi, scene


def half_to_float(bits):
    return np.asarray(bits, np.uint16).view(np.float16).astype(np.float32)


def node_bounds(nodes):
    lo = nodes["min"].astype(np.float32)
    hi = lo + half_to_float(nodes["d"])
    return lo, hi


def test_box_scene_matches_survey_inventory(box_scene):
    st = box_scene.stats()
    # 6 box meshes + 2 GLB meshes, 4032 triangles, 11 984 vertices, 4035 leaves -> 8069 nodes
    assert st["meshes"] == 8 and st["triangles"] == 4032 and st["vertices"] == 11984
    assert st["spheres"] == 2 and st["discs"] == 1 and st["bvh_nodes"] == 8069 and st["bvh_bytes"] == 193656
    assert list(box_scene.mat_ids) == [4, 0, 1, 2, 0, 5, 0, 0, 3, 7, 6]
    assert [int(t) for t in box_scene.geometry["type"]] == [0] * 8 + [1, 1, 2]
    assert [int(m["numTriangles"]) for m in box_scene.mesh_info] == [2, 6, 2, 2, 10, 10, 64, 3936]
    assert box_scene.materials.size == 8
    m = box_scene.materials
    assert int(m["type"][3]) == 2 and int(m["type"][5]) == 1 and bool(m["emissive"][4]) and bool(m["emissive"][6])
    assert np.allclose(m["emission"][4], [(100 * 15.6 + 100 * 18.4) / 255, (100 * 8 + 74.5 * 15.6) / 255, 57.3 * 8 / 255], rtol=1e-6)
    # camera-space transform (scene_utils.cpp:474-508): spheres at x = -(450-278), z = -(90+800)
    assert np.array_equal(box_scene.spheres, np.float32([[-172, -236, -890, 37], [-72, -236, -890, 37]]))
    d = box_scene.discs[0]
    assert d[0] == -1 and np.signbit(d[2]) and d[3] == 60 and abs(d[4] - 277.9998) < 1e-3 and d[6] == -1050
    assert abs(box_scene.fov - np.pi / 4) < 1e-7


def test_spheres_scene(spheres_scene):
    st = spheres_scene.stats()
    assert st["spheres"] == 5 and st["discs"] == 1 and st["meshes"] == 0 and st["bvh_nodes"] == 11
    assert list(spheres_scene.mat_ids) == [0, 1, 2, 3, 4, 5]
    assert abs(spheres_scene.fov - np.pi / 2) < 1e-7


def test_box_simple_has_only_the_box():
    s = scene.HostScene.builtin("box-simple")
    assert s.stats()["meshes"] == 6 and s.stats()["spheres"] == 0 and s.stats()["triangles"] == 32


def test_unknown_scene_is_an_error():
    with pytest.raises(RuntimeError, match="Invalid scene selection"):
        scene.HostScene.builtin("nope")


@pytest.mark.parametrize("fixture", ["box_scene", "spheres_scene", "dae_scene", "hdri_scene"])
def test_bvh_is_preorder_binary_and_conservative(fixture, request):
    s = request.getfixturevalue(fixture)
    nodes = s.bvh_nodes
    n = nodes.size
    lo, hi = node_bounds(nodes)
    leaf = nodes["geomID"] != 0xFFFF
    assert leaf.sum() * 2 - 1 == n
    # walk pre-order: first child = index + 1, second child index stored; verify containment and depth
    max_depth, seen_leaves = 0, []
    stack = [(0, 1)]
    while stack:
        i, depth = stack.pop()
        max_depth = max(max_depth, depth)
        if leaf[i]:
            seen_leaves.append((int(nodes["geomID"][i]), int(nodes["primOrSecondChild"][i])))
            continue
        c0, c1 = i + 1, int(nodes["primOrSecondChild"][i])
        assert i < c0 < c1 < n
        for c in (c0, c1):
            # each node rounds its own fp16 extent up, so a child may poke out by < 1 fp16 ulp of its extent
            slack = half_to_float(nodes["d"][c]) * np.float32(2.0 ** -10) + np.float32(1e-4)
            assert np.all(lo[c] >= lo[i]) and np.all(hi[c] <= hi[i] + slack)
        stack.append((c1, depth + 1))
        stack.append((c0, depth + 1))
    assert max_depth == s.desc.max_leaf_depth <= 64
    # every primitive appears in exactly one leaf
    expect = []
    for g, geom in enumerate(s.geometry):
        if geom["type"] == 0:
            expect += [(g, t) for t in range(int(s.mesh_info[int(geom["index"])]["numTriangles"]))]
        else:
            expect.append((g, 0))
    assert sorted(seen_leaves) == sorted(expect)


def test_leaf_boxes_contain_their_triangles_with_fp16_extents_rounded_up(box_scene):
    s = box_scene
    nodes = s.bvh_nodes
    lo, hi = node_bounds(nodes)
    verts = np.stack([s.mesh_verts["x"], s.mesh_verts["y"], s.mesh_verts["z"]], axis=1)
    for i in np.nonzero(nodes["geomID"] != 0xFFFF)[0][::7]:
        g = int(nodes["geomID"][i])
        if s.geometry["type"][g] != 0:
            continue
        mi = s.mesh_info[int(s.geometry["index"][g])]
        tri = s.mesh_tris["v"][int(mi["firstIndex"]) + int(nodes["primOrSecondChild"][i])]
        p = verts[int(mi["firstVertex"]) + tri.astype(np.int64)]
        assert np.array_equal(lo[i], p.min(0))          # min is stored exactly in fp32
        assert np.all(hi[i] >= p.max(0))                # extent never rounded down (roundToHalfNotSmaller)
        exact = p.max(0) - p.min(0)
        assert np.all(half_to_float(nodes["d"][i]) - exact <= np.maximum(exact * 2.0 ** -10, 2.0 ** -24))


def test_round_to_half_not_smaller_matches_oracle(port):
    import ctypes as C
    rng = np.random.default_rng(3)
    n = 300
    lo = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    ext = rng.uniform(0, 700, (n, 3)).astype(np.float32)
    b = np.concatenate([lo, lo + ext], axis=1).astype(np.float32)
    ids = np.stack([np.zeros(n), np.arange(n)], axis=1).astype(np.uint32)
    nodes = np.zeros(2 * n - 1, capi.BVH_NODE)
    depth = C.c_uint32()
    assert capi.scene_lib().b200rt_build_bvh(capi.ptr(b), capi.ptr(ids), n, capi.ptr(nodes), C.byref(depth)) == 2 * n - 1
    leaves = nodes[nodes["geomID"] != 0xFFFF]
    order = leaves["primOrSecondChild"].astype(np.int64)
    exact = (b[order, 3:] - b[order, :3]).astype(np.float32)
    assert np.array_equal(leaves["d"].ravel(), port.round_to_half_not_smaller(exact.ravel()))


def test_bvh_rejects_extents_beyond_fp16():
    import ctypes as C
    b = np.float32([[0, 0, 0, 70000, 1, 1], [1, 1, 1, 2, 2, 2]])
    ids = np.uint32([[0, 0], [0, 1]])
    nodes = np.zeros(3, capi.BVH_NODE)
    rc = capi.scene_lib().b200rt_build_bvh(capi.ptr(b), capi.ptr(ids), 2, capi.ptr(nodes), None)
    assert rc < 0 and b"fp16" in capi.scene_lib().b200rt_scene_last_error()


def test_ray_stream_matches_oracle_ray_directions(port, box_scene):
    w, h = 97, 61  # non-square, odd sizes
    rays = scene.init_ray_stream(w, h, box_scene.fov)
    assert rays.size == w * h
    s, c = port.sincos(np.float32([box_scene.fov / 2]))
    tan_theta = np.float32(s[0]) / np.float32(c[0])
    xy = np.stack([rays["p"][:, 1], rays["p"][:, 0]], axis=1)  # (col, row)
    assert np.array_equal(rays["h"]["r"]["direction"], port.pixel_to_ray_dir(xy, float(w), float(h), float(tan_theta)))
    assert np.array_equal(rays["p"][:w, 1], np.arange(w, dtype=np.float32)) and rays["p"][w, 0] == 1.0
    assert np.all(rays["h"]["geomID"] == 0xFFFF) and np.all(rays["h"]["primID"] == 0xFFFFFFFF)
    assert np.all(rays["h"]["normal"] == np.float32([0, 0, 1])) and np.all(np.isinf(rays["h"]["r"]["tMax"]))
    assert np.all(rays["rgb"] == 0) and np.all(rays["h"]["flags"] == 0)


def test_crop_window_uses_full_image_coordinates(box_scene):
    full = scene.init_ray_stream(64, 48, box_scene.fov)
    crop = scene.init_ray_stream(64, 48, box_scene.fov, window=(16, 8, 20, 10))  # w,h,col,row
    assert crop.size == 128
    sel = full.reshape(48, 64)[10:18, 20:36].ravel()
    assert crop.tobytes() == sel.tobytes()


def test_host_sincos_matches_oracle(port):
    x = np.random.default_rng(1).uniform(-30, 30, 2000).astype(np.float32)
    s, c = port.sincos(x)
    for xi, si, ci in zip(x[:500], s, c):
        hs, hc = scene.host_sincos(float(xi))
        assert np.float32(hs) == si and np.float32(hc) == ci


def test_visualise_modes_and_image_writers(tmp_path, port, box_scene):
    w = h = 48
    box_scene.configure(w, h, path_trace=False)
    rays = scene.init_ray_stream(w, h, box_scene.fov)
    port.shadow_trace(box_scene, rays)
    hit = rays["h"]["geomID"] != 0xFFFF
    img, hits = scene.visualise_hits(rays, box_scene, "normal", w, h)
    assert hits == hit.sum() and img.shape == (h, w, 3)
    assert np.array_equal(img.reshape(-1, 3)[hit][:, ::-1], rays["h"]["normal"][hit])  # stored BGR
    assert np.all(img.reshape(-1, 3)[~hit] == 0)
    ids, _ = scene.visualise_hits(rays, box_scene, "id", w, h)
    flat = ids.reshape(-1, 3)
    assert np.array_equal(flat[hit][:, 0], rays["h"]["geomID"][hit] + 1.0)
    assert np.array_equal(flat[hit][:, 2], box_scene.mat_ids[rays["h"]["geomID"][hit]] + 1.0)
    tfar, _ = scene.visualise_hits(rays, box_scene, "tfar", w, h)
    assert np.array_equal(tfar[..., 0].ravel(), rays["h"]["r"]["tMax"])
    rgb, _ = scene.visualise_hits(rays, box_scene, "rgb", w, h)
    # PFM round trip (bottom-up RGB)
    scene.write_pfm(tmp_path / "a.pfm", rgb)
    raw = (tmp_path / "a.pfm").read_bytes()
    header_end = raw.index(b"-1.0\n") + 5
    data = np.frombuffer(raw[header_end:], np.float32).reshape(h, w, 3)[::-1, :, ::-1]
    assert np.array_equal(data, rgb)
    # EXR: check the structure we wrote (magic, channel list, pixel payload of the first scanline)
    scene.write_exr(tmp_path / "a.exr", rgb)
    exr = (tmp_path / "a.exr").read_bytes()
    assert exr[:4] == bytes([0x76, 0x2F, 0x31, 0x01]) and b"channels\x00chlist\x00" in exr
    assert len(exr) > h * w * 12
    first_row_b = np.frombuffer(exr[-(h * (8 + 12 * w)):][8:8 + 4 * w], np.float32)
    assert np.array_equal(first_row_b, rgb[0, :, 0])
    cv2 = pytest.importorskip("cv2")
    import os
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        back = cv2.imread(str(tmp_path / "a.exr"), cv2.IMREAD_UNCHANGED)
    except cv2.error:
        back = None
    if back is not None:
        assert np.array_equal(back, rgb)


def test_nif_metadata_reader():
    import ctypes as C
    md = capi.NifMetadata()
    path = capi.REPO_ROOT / "assets/nif/urban_alley_01_4k_fp16_yuv/assets.extra/nif_metadata.txt"
    assert capi.scene_lib().b200rt_read_nif_metadata(str(path).encode(), C.byref(md)) == 0
    assert md.embedding_dimension == 12 and md.hidden_size == 320 and list(md.image_shape) == [2048, 4096, 3]
    assert md.log_tone_map == 1 and abs(md.max - 3.4299468994140625) < 1e-7
    assert abs(md.mean[0] - (-2.3514461517333984 - 1e-8)) < 1e-6  # eps folded in (NifMetaData.cpp:48-53)


# ---- the reference's serialised SceneRef (Serialiser<16>) --------------------------------------------------------
def _oracle_blob(orc, s):
    import ctypes as C

    from ipu_ray_lib_b200 import _capi as capi
    f = orc._f("serialise_scene")
    f.restype = C.c_size_t
    f.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t]
    n = f(C.byref(s.desc), None, 0)
    buf = np.zeros(n, np.uint8)
    assert f(C.byref(s.desc), capi.ptr(buf), n) == n
    return buf


@pytest.mark.parametrize("fixture", ["box_scene", "spheres_scene", "dae_scene"])
def test_serialised_scene_round_trip_and_matches_the_reference_writer(fixture, request, port):
    """b200rt_scene_blob_write == the restated writer (== the reference's own Serialiser<16>, next test), and
    b200rt_scene_desc_from_blob reads it back in place: same arrays, same scalars."""
    from ipu_ray_lib_b200.scene import BlobScene, scene_blob
    s = request.getfixturevalue(fixture)
    s.configure(640, 480, samples=77, max_path_length=7, roulette_start_depth=2, anti_alias=0.5)
    blob = scene_blob(s)
    assert blob.tobytes() == _oracle_blob(port, s).tobytes()
    b = BlobScene(blob, spheres=s.spheres, discs=s.discs)
    for name in ("num_geometry", "num_meshes", "num_tris", "num_verts", "num_normals", "num_mat_ids", "num_materials",
                 "num_bvh_nodes", "max_leaf_depth", "image_width", "image_height", "fov_radians", "anti_alias_scale",
                 "max_path_length", "roulette_start_depth", "samples_per_pixel", "num_spheres", "num_discs"):
        assert getattr(b.desc, name) == getattr(s.desc, name), name
    import ctypes as C
    for ptr, count, size in (("geometry", "num_geometry", 4), ("mesh_info", "num_meshes", 16), ("mesh_tris", "num_tris", 6),
                             ("mesh_verts", "num_verts", 12), ("mesh_normals", "num_normals", 12),
                             ("mat_ids", "num_mat_ids", 4), ("materials", "num_materials", 36), ("bvh_nodes", "num_bvh_nodes", 24)):
        n = getattr(s.desc, count) * size
        if n:
            pa = C.cast(getattr(b.desc, ptr), C.c_void_p).value
            pb = C.cast(getattr(s.desc, ptr), C.c_void_p).value
            assert C.string_at(pa, n) == C.string_at(pb, n), ptr
            assert b.blob.ctypes.data <= pa < b.blob.ctypes.data + b.blob.size  # zero copy: points into the blob


def test_serialised_scene_is_the_reference_byte_stream(ref, port, box_scene, dae_scene):
    for s in (box_scene, dae_scene):
        s.configure(1440, 1440, samples=1000)
        assert _oracle_blob(ref, s).tobytes() == _oracle_blob(port, s).tobytes()


def test_malformed_serialised_scenes_are_rejected(box_scene):
    from ipu_ray_lib_b200.scene import BlobScene, scene_blob
    blob = scene_blob(box_scene)
    with pytest.raises(RuntimeError, match="truncated"):
        BlobScene(blob[:1000])
    with pytest.raises(RuntimeError, match="trailing"):
        BlobScene(np.concatenate([blob, np.zeros(8, np.uint8)]))
    bad = blob.copy()
    bad[:4] = 255  # absurd element count
    with pytest.raises(RuntimeError, match="truncated"):
        BlobScene(bad)
