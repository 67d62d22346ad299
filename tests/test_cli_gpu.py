"""The C++ surface on a GPU: the `trace` CLI (host/trace_main.cpp) through the C++ mirror of IpuScene (host/B200Scene.hpp:
replica threads, strided batches of one shared stream, callback numbering k * R + replica) against the oracle."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_streams_identical
from ipu_ray_lib_b200 import _capi as capi, scene

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
TRACE = ROOT / "ipu_ray_lib_b200" / "trace"


def run_trace(tmp_path, *flags):
    out = tmp_path / "rays.bin"
    cmd = [str(TRACE), "-o", str(tmp_path / "img"), "--builtin-mesh", str(ROOT / "assets" / "monkey_bust.glb"),
           "--save-ray-stream", str(out), "--log-level", "debug", *flags]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.fromfile(out, dtype=capi.TRACE_RESULT), r.stderr


@pytest.mark.parametrize("ipus", [1, 2, 8])
def test_cli_path_trace_matches_oracle_for_any_replica_count(tmp_path, port, ipus):
    """--ipus N is clamped to the B200s present; whatever N ends up being, the image is the 1-replica image bit for bit
    (per-(pixel,sample) RNG streams), the callback sees every batch exactly once under the reference's numbering."""
    w, h, spp = 320, 200, 5
    rays, log = run_trace(tmp_path, "--scene", "box", "-w", str(w), "-h", str(h), "--samples", str(spp), "--ipus", str(ipus),
                          "--ipu-ray-callback", "--seed", "1442")
    s = scene.HostScene.builtin("box").configure(w, h, path_trace=True, samples=spp, seed=1442)
    want = scene.init_ray_stream(w, h, s.fov)
    port.path_trace(s, want)
    assert_streams_identical(rays, want, f"trace --ipus {ipus}")
    batches = sorted(int(m) for m in re.findall(r"Application callback received batch (\d+)", log))
    assert batches == list(range((w * h + 8639) // 8640))
    assert (tmp_path / "img_rgb_b200.exr").exists()


def test_cli_shadow_trace_normals_and_imported_scene(tmp_path, port):
    w, h = 256, 192
    rays, _ = run_trace(tmp_path, "--scene", "box", "-w", str(w), "-h", str(h), "--render-mode", "shadow-trace",
                        "--visualise", "normal", "--ipus", "2")
    s = scene.HostScene.builtin("box").configure(w, h, path_trace=False)
    want = scene.init_ray_stream(w, h, s.fov)
    port.shadow_trace(s, want)
    assert_streams_identical(rays, want, "trace shadow-trace")
    assert (tmp_path / "img_normal_b200.exr").exists()
    rays, _ = run_trace(tmp_path, "--mesh-file", str(ROOT / "assets" / "test_scene.dae"), "--load-normals", "-w", "160",
                        "-h", "120", "--samples", "3", "--ipus", "1")
    s = scene.HostScene.from_file(ROOT / "assets" / "test_scene.dae", load_normals=True).configure(160, 120, path_trace=True, samples=3)
    want = scene.init_ray_stream(160, 120, s.fov)
    port.path_trace(s, want)
    assert_streams_identical(rays, want, "trace --mesh-file test_scene.dae --load-normals")


def test_cli_loads_a_keras_h5_nif_like_the_reference(tmp_path, port):
    """trace --nif-hdri <assets.extra>: nif_metadata.txt + converted.hdf5 (src/IpuScene.cpp:174-187), the weights read
    by the HDF5 reader; the NIF-lit image agrees with the oracle within the NIF tolerance and the hit records exactly."""
    import shutil

    from ipu_ray_lib_b200.nif import NifWeights

    extra = tmp_path / "assets.extra"
    extra.mkdir()
    md = ROOT / "assets/nif/urban_alley_01_4k_fp16_yuv/assets.extra/nif_metadata.txt"
    shutil.copy(md, extra / "nif_metadata.txt")
    nif = NifWeights.from_metadata(md, seed=99)
    nif.save(extra / "converted.hdf5")
    w, h, spp = 200, 120, 4
    rays, log = run_trace(tmp_path, "--scene", "spheres", "-w", str(w), "-h", str(h), "--samples", str(spp), "--ipus", "1",
                          "--nif-hdri", str(extra), "--hdri-rotation", "35", "--max-nif-batch-size", "4096")
    assert "Loaded NIF model" in log
    s = scene.HostScene.builtin("spheres").configure(w, h, path_trace=True, samples=spp, seed=1442)
    want = scene.init_ray_stream(w, h, s.fov)
    port.path_trace(s, want, nif=NifWeights.load(extra / "converted.hdf5"), hdri_rotation=35.0)
    a, b = rays.copy(), want.copy()
    a["rgb"] = 0
    b["rgb"] = 0
    assert a.tobytes() == b.tobytes()
    err = np.abs(rays["rgb"].astype(np.float64) - want["rgb"]).sum() / np.abs(want["rgb"]).sum()
    assert err < 8e-4, err
