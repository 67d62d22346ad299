"""The C-ABI library loads, exports every symbol the headers declare, and fails loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from ipu_ray_lib_b200 import _capi as capi

ROOT = Path(__file__).resolve().parents[1]


def _declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-z0-9_]+)\s*\(", text)))


def test_headers_and_symbol_lists_agree():
    assert _declared("b200rt.h") == sorted(capi.B200RT_SYMBOLS)
    assert _declared("b200rt_scene.h") == sorted(capi.B200RT_SCENE_SYMBOLS)


def test_trace_library_exports_every_declared_symbol():
    lib = capi.lib()
    for name in _declared("b200rt.h"):
        assert hasattr(lib, name), name
    assert lib.b200rt_abi_version() == 1


def test_scene_library_exports_every_declared_symbol():
    lib = capi.scene_lib()
    for name in _declared("b200rt_scene.h"):
        assert hasattr(lib, name), name


def test_wire_layouts_match_reference_sizes():
    # SURVEY.md Appendix C (probed from the reference headers)
    assert capi.TRACE_RESULT.itemsize == 84 and capi.TRACE_RESULT.fields["p"][1] == 12
    assert capi.TRACE_RESULT.fields["h"][1] == 20
    hr = capi.HIT_RECORD
    assert hr.itemsize == 64
    assert (hr.fields["primID"][1], hr.fields["normal"][1], hr.fields["throughput"][1], hr.fields["geomID"][1],
            hr.fields["flags"][1]) == (32, 36, 48, 60, 62)
    bn = capi.BVH_NODE
    assert bn.itemsize == 24 and bn.fields["primOrSecondChild"][1] == 12 and bn.fields["d"][1] == 16
    assert bn.fields["geomID"][1] == 22
    m = capi.MATERIAL
    assert m.itemsize == 36 and m.fields["ior"][1] == 12 and m.fields["emission"][1] == 16
    assert m.fields["type"][1] == 28 and m.fields["emissive"][1] == 32


def test_c_struct_sizes():
    # x86-64 SysV layout of the headers' structs
    assert C.sizeof(capi.TraceParams) == 4 * 4 + 4 * 7 + 4 * 5
    assert C.sizeof(capi.TraceStats) == 8 * 6 + 8 * 4 + 8 + 8 * 3 + 8 * 2 + 8 * 2
    assert C.sizeof(capi.NifLayer) == 32
    assert C.sizeof(capi.SceneDesc) % 8 == 0


def test_ctypes_mirrors_follow_the_header_field_by_field():
    """The ctypes structures name the same fields in the same order as include/b200rt.h (a renamed or inserted field
    in one of them would otherwise pass the size checks)."""
    import re
    from pathlib import Path
    hdr = (Path(__file__).resolve().parents[1] / "include" / "b200rt.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def header_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:  # "type a, b" declares two fields
                first, *rest = decl.split(",")
                for d in [first.split()[-1]] + [r.strip() for r in rest]:
                    out.append(re.sub(r"\[.*", "", d.lstrip("*")))
        return out

    for cname, mirror in (("b200rt_trace_params", capi.TraceParams), ("b200rt_trace_stats", capi.TraceStats),
                          ("b200rt_scene_desc", capi.SceneDesc)):
        assert header_fields(cname) == [f for f, _ in mirror._fields_], cname


def test_bad_arguments_are_rejected_with_a_message(box_scene):
    lib = capi.lib()
    out = C.c_void_p()
    assert lib.b200rt_scene_create(None, C.byref(out)) == -1
    assert b"null" in lib.b200rt_last_error()
    d = capi.SceneDesc()
    C.memmove(C.byref(d), C.byref(box_scene.desc), C.sizeof(d))
    d.num_bvh_nodes = 0
    assert lib.b200rt_scene_create(C.byref(d), C.byref(out)) == -1
    assert b"BVH" in lib.b200rt_last_error()
    C.memmove(C.byref(d), C.byref(box_scene.desc), C.sizeof(d))
    d.num_mat_ids = 3  # fewer materials than primitives: the reference throws std::logic_error
    assert lib.b200rt_scene_create(C.byref(d), C.byref(out)) == -1
    assert b"material" in lib.b200rt_last_error()
    C.memmove(C.byref(d), C.byref(box_scene.desc), C.sizeof(d))
    d.path_trace, d.max_path_length = 1, 0
    assert lib.b200rt_scene_create(C.byref(d), C.byref(out)) == -1
    assert b"max_path_length" in lib.b200rt_last_error()
    assert not out.value


@pytest.mark.skipif(capi.lib().b200rt_device_count() > 0, reason="CPU-only behaviour")
def test_no_cpu_fallback_without_a_gpu(box_scene):
    """The product path must fail loudly when there is no B200: no CPU/oracle fallback exists."""
    lib = capi.lib()
    out = C.c_void_p()
    rc = lib.b200rt_scene_create(C.byref(box_scene.desc), C.byref(out))
    assert rc == -2 and not out.value
    assert b"no CUDA device" in lib.b200rt_last_error() or b"CUDA" in lib.b200rt_last_error()


def test_product_never_links_or_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may reference it."""
    offenders = []
    for p in list((ROOT / "ipu_ray_lib_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc"}:
            t = p.read_text(errors="ignore")
            # code references (includes, imports, library names); prose mentions in comments are fine
            if re.search(r"#include[^\n]*oracle|oracle_py|liboracle|oracle_api\.h|from oracle|import oracle", t):
                offenders.append(str(p))
    assert not offenders, offenders
    import subprocess

    needed = subprocess.run(["readelf", "-d", str(capi.lib_path())], capture_output=True, text=True).stdout
    assert "oracle" not in needed
