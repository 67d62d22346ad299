import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything under test is native and built in-tree; build once if a library is missing."""
    from ipu_ray_lib_b200 import _capi as capi
    from oracle import oracle_py

    if not (capi.lib_path().exists() and capi.scene_lib_path().exists() and oracle_py.PORT_LIB.exists()):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "reference_vectors.npz")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle_py import Oracle

    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own kernel sources (oracle/_ref). Absent only where it was never built."""
    from oracle import oracle_py

    if not oracle_py.have_ref():
        pytest.skip("oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
    return oracle_py.Oracle("reference")


@pytest.fixture(scope="session")
def box_scene():
    from ipu_ray_lib_b200 import HostScene

    return HostScene.builtin("box")


@pytest.fixture(scope="session")
def spheres_scene():
    from ipu_ray_lib_b200 import HostScene

    return HostScene.builtin("spheres")


def words(rays):
    """TraceResult stream as [n, 21] uint32 words for exact comparisons."""
    return rays.view(np.uint32).reshape(rays.size, 21)


WORD_NAMES = (["rgb.x", "rgb.y", "rgb.z", "row", "col", "o.x", "o.y", "o.z", "tMin", "d.x", "d.y", "d.z", "tMax",
               "primID", "n.x", "n.y", "n.z", "thr.x", "thr.y", "thr.z", "geom|flags"])


def assert_streams_identical(a, b, what=""):
    if a.tobytes() == b.tobytes():
        return
    wa, wb = words(a), words(b)
    bad = np.nonzero((wa != wb).any(axis=1))[0]
    per_word = {WORD_NAMES[i]: int(c) for i, c in enumerate((wa != wb).sum(axis=0)) if c}
    raise AssertionError(f"{what}: {bad.size}/{a.size} rays differ; per field {per_word}; first {bad[:5]}\n"
                         f"a={a[bad[0]]}\nb={b[bad[0]]}")


@pytest.fixture(scope="session")
def dae_scene():
    """assets/test_scene.dae with interpolated normals (BASELINE.json config 3)."""
    from ipu_ray_lib_b200 import HostScene

    return HostScene.from_file(ROOT / "assets" / "test_scene.dae", load_normals=True)


@pytest.fixture(scope="session")
def hdri_scene():
    """assets/hdri_test.dae, open sky (BASELINE.json config 5)."""
    from ipu_ray_lib_b200 import HostScene

    return HostScene.from_file(ROOT / "assets" / "hdri_test.dae", load_normals=False)
