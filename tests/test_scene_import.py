"""The COLLADA scene importer (host/scene_import.cpp) behind `--mesh-file` (src/scene_utils.cpp:151-317).

assimp is not available, so what is checked is (a) the inventory SURVEY.md appendix D probed from the files,
(b) the reference's material heuristics, (c) the transform chain on a hand-written document whose answer is known
in closed form, (d) that both CPU checkers agree on the imported arrays.
"""
import numpy as np
import pytest

from conftest import ROOT, assert_streams_identical
from ipu_ray_lib_b200 import HostScene, scene

DIFFUSE, SPECULAR, REFRACTIVE = 0, 1, 2


def test_test_scene_inventory(dae_scene):
    st = dae_scene.stats()
    assert st["triangles"] == 8474 and st["bvh_nodes"] == 2 * 8474 - 1 and st["bvh_bytes"] == 406728
    assert st["spheres"] == 0 and st["discs"] == 0
    # ten instances, two of them share `sandy`: one output mesh per material, ascending material index
    assert st["meshes"] == 9 and dae_scene.mat_ids.tolist() == list(range(9))
    assert dae_scene.mesh_normals.size == dae_scene.mesh_verts.size  # --load-normals
    assert abs(dae_scene.fov - np.deg2rad(45.0)) < 1e-6
    m = dae_scene.materials
    # library order: green_glass clear_glass metal badge green_glass.001 red_light blue_light sandy base_reflective
    assert m["type"].tolist() == [REFRACTIVE, REFRACTIVE, SPECULAR, DIFFUSE, REFRACTIVE, DIFFUSE, DIFFUSE, DIFFUSE,
                                  SPECULAR]
    assert m["emissive"].tolist() == [0, 0, 0, 0, 0, 1, 1, 0, 0]
    np.testing.assert_allclose(m["ior"][:3], [1.41, 1.5, 1.45])
    # the two lights: file emission x the hand-edited shininess (10 and 350)
    assert m["emission"][5][0] == np.float32(1.0) * np.float32(10.0)
    assert m["emission"][6][2] == np.float32(1.0) * np.float32(350.0)
    np.testing.assert_allclose(m["albedo"][0], [0.03948122, 0.3343189, 0.03890199])
    n = np.stack([dae_scene.mesh_normals[k] for k in "xyz"], -1)
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)


def test_hdri_scene_inventory(hdri_scene):
    st = hdri_scene.stats()
    assert st["triangles"] == 5656 and st["bvh_nodes"] == 11311 and st["bvh_bytes"] == 271464
    assert st["meshes"] == 5 and hdri_scene.mat_ids.tolist() == [0, 1, 2, 3, 4]
    assert hdri_scene.mesh_normals.size == 0  # normals only with --load-normals
    assert abs(hdri_scene.fov - np.deg2rad(54.43)) < 1e-4
    assert hdri_scene.materials["type"].tolist() == [REFRACTIVE, SPECULAR, REFRACTIVE, DIFFUSE, SPECULAR]
    assert not hdri_scene.materials["emissive"].any()  # lit by the environment only


def test_import_is_deterministic_and_normals_do_not_move_vertices(dae_scene):
    again = HostScene.from_file(ROOT / "assets" / "test_scene.dae", load_normals=True)
    assert again.mesh_verts.tobytes() == dae_scene.mesh_verts.tobytes()
    assert again.bvh_nodes.tobytes() == dae_scene.bvh_nodes.tobytes()
    bare = HostScene.from_file(ROOT / "assets" / "test_scene.dae", load_normals=False)
    assert bare.mesh_verts.tobytes() == dae_scene.mesh_verts.tobytes()
    assert bare.mesh_tris.tobytes() == dae_scene.mesh_tris.tobytes()
    assert bare.mesh_normals.size == 0


def test_camera_looks_down_minus_z_at_the_scene(dae_scene):
    """After import the camera sits at the origin looking along -z (src/scene_utils.cpp:286-314)."""
    v = np.stack([dae_scene.mesh_verts[k] for k in "xyz"], -1)
    assert (v[:, 2] < 0).mean() > 0.99
    # the ground plane extends below the camera, the lights above it
    assert v[:, 1].min() < -0.5 < 0.5 < v[:, 1].max()


DOC = """<?xml version="1.0"?>
<!-- hand-written: exercises polylist, nested nodes, translate/rotate/scale, a quad, a default material -->
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
  <asset><up_axis>Y_UP</up_axis></asset>
  <library_cameras><camera id="cam"><optics><technique_common><perspective>
     <yfov>90</yfov><aspect_ratio>1</aspect_ratio></perspective></technique_common></optics></camera></library_cameras>
  <library_effects>
    <effect id="fx"><profile_COMMON><technique sid="common"><phong>
      <emission><color>0 0 0 1</color></emission><diffuse><color>0.25 0.5 0.75 1</color></diffuse>
      <reflectivity><float>0.5</float></reflectivity></phong></technique></profile_COMMON></effect>
    <effect id="glow"><profile_COMMON><technique sid="common"><lambert>
      <emission><color>1 2 3 1</color></emission></lambert></technique></profile_COMMON></effect>
  </library_effects>
  <library_materials>
    <material id="m0" name="mirror &amp; co"><instance_effect url="#fx"/></material>
    <material id="m1" name="lamp"><instance_effect url="#glow"/></material>
  </library_materials>
  <library_geometries><geometry id="quad"><mesh>
    <source id="quad-pos"><float_array id="a" count="12">-1 -1 0  1 -1 0  1 1 0  -1 1 0</float_array>
      <technique_common><accessor source="#a" count="4" stride="3"/></technique_common></source>
    <source id="quad-n"><float_array id="b" count="3">0 0 1</float_array>
      <technique_common><accessor source="#b" count="1" stride="3"/></technique_common></source>
    <vertices id="quad-v"><input semantic="POSITION" source="#quad-pos"/></vertices>
    <polylist material="S" count="1"><input semantic="VERTEX" source="#quad-v" offset="0"/>
      <input semantic="NORMAL" source="#quad-n" offset="1"/><vcount>4</vcount><p>0 0 1 0 2 0 3 0</p></polylist>
  </mesh></geometry></library_geometries>
  <library_visual_scenes><visual_scene id="S0">
    <node id="outer"><translate>0 0 -10</translate>
      <node id="inner"><rotate>0 0 1 90</rotate><scale>2 1 1</scale>
        <instance_geometry url="#quad"><bind_material><technique_common>
          <instance_material symbol="S" target="#MATERIAL"/></technique_common></bind_material></instance_geometry>
      </node></node>
    <node id="camnode"><translate>0 0 5</translate><instance_camera url="#cam"/></node>
  </visual_scene></library_visual_scenes>
  <scene><instance_visual_scene url="#S0"/></scene>
</COLLADA>
"""


@pytest.mark.parametrize("target,want_type,want_emissive", [("m0", SPECULAR, 0), ("m1", DIFFUSE, 1), ("nothing", DIFFUSE, 0)])
def test_handwritten_document_transform_chain_and_materials(tmp_path, target, want_type, want_emissive):
    f = tmp_path / "quad.dae"
    f.write_text(DOC.replace("#MATERIAL", "#" + target))
    s = HostScene.from_file(f, load_normals=True)
    assert s.stats()["triangles"] == 2 and s.stats()["vertices"] == 4 and s.stats()["meshes"] == 1
    # quad -> scale x by 2 -> rotate 90 deg about z: (x, y) -> (-y, 2x) -> translate z-10 ; camera at z=+5 looking
    # down -z with y up, then the reference's (-x, y, -z) flip twice over (camera matrix x/z axes, handedness swap)
    v = np.stack([s.mesh_verts[k] for k in "xyz"], -1)
    want = np.array([[1, -2, -15], [1, 2, -15], [-1, 2, -15], [-1, -2, -15]], dtype=np.float32)
    np.testing.assert_allclose(v, want, atol=1e-5)
    n = np.stack([s.mesh_normals[k] for k in "xyz"], -1)
    np.testing.assert_allclose(n, np.tile([0, 0, 1], (4, 1)), atol=1e-6)  # facing the camera
    assert s.mesh_tris.view(np.uint16).tolist() == [0, 1, 2, 0, 2, 3]  # fan triangulation
    assert abs(s.fov - np.pi / 2) < 1e-6  # yfov 90 at aspect 1
    mat = s.materials[s.mat_ids[0]]
    assert mat["type"] == want_type and mat["emissive"] == want_emissive
    if target == "m1":  # no <shininess>: the importer default (10) scales the emission, like with assimp
        assert mat["emission"].tolist() == [10.0, 20.0, 30.0]
    if target == "nothing":  # unbound symbol -> appended default material (0.6 grey)
        assert s.mat_ids[0] == 2 and mat["albedo"].tolist() == [np.float32(0.6)] * 3


@pytest.mark.parametrize("text,msg", [
    ("<COLLADA><asset></COLLADA>", "XML parse error"),
    ("<COLLADA", "XML parse error"),
    ("<notcollada/>", "Could not load scene file"),
    (DOC.replace('<instance_camera url="#cam"/>', ""), "No camera found"),
    (DOC.replace("<p>0 0 1 0 2 0 3 0</p>", "<p>0 0 1 0 2 0 9 0</p>"), "beyond <source> count"),
    (DOC.replace("<vcount>4</vcount>", "<vcount>5</vcount>"), "shorter than <vcount>"),
])
def test_malformed_documents_are_errors(tmp_path, text, msg):
    f = tmp_path / "bad.dae"
    f.write_text(text)
    with pytest.raises(RuntimeError, match=msg):
        HostScene.from_file(f)


def test_files_the_reference_rejects():
    with pytest.raises(RuntimeError, match="No camera found"):  # src/scene_utils.cpp:176-180
        HostScene.from_file(ROOT / "assets" / "monkey_bust.glb")
    with pytest.raises(RuntimeError, match="Could not load scene file"):
        HostScene.from_file(ROOT / "assets" / "missing.dae")
    with pytest.raises(RuntimeError, match="Could not load scene file"):
        HostScene.from_file(ROOT / "assets" / "nif_metadata.txt")


def test_checkers_agree_on_imported_scenes(port, ref, dae_scene):
    """Restatement vs the reference's own compiled kernels on an imported scene, interpolated normals included."""
    w = h = 48
    dae_scene.configure(w, h, path_trace=True, samples=4, seed=7)
    base = scene.init_ray_stream(w, h, dae_scene.fov)
    a, b = base.copy(), base.copy()
    ca = port.path_trace(dae_scene, a)
    cb = ref.path_trace(dae_scene, b)
    assert_streams_identical(a, b, "port vs reference build, test_scene.dae")
    for k in ("closest_hit_queries", "samples", "escaped_samples"):  # the reference build has no visit counters
        assert ca[k] == cb[k], k
    dae_scene.configure(w, h, path_trace=False)
    a, b = base.copy(), base.copy()
    port.shadow_trace(dae_scene, a, light=(0.0, 6.0, -3.0))
    ref.shadow_trace(dae_scene, b, light=(0.0, 6.0, -3.0))
    assert_streams_identical(a, b, "port vs reference build, shadow trace")
