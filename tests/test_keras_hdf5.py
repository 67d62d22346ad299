"""N3: the reference's NIF container, a Keras "h5" model (src/keras/Hdf5Model.cpp:8-133), read by the self-contained
HDF5 reader (host/keras_hdf5.cpp) through the C ABI; fixtures are written by ipu_ray_lib_b200/keras_h5.py following the
HDF5 File Format Specification (h5py / libhdf5 do not exist here, so the reader is NOT pinned against a file produced by
Keras itself: the file layout is the spec's classic one h5py emits by default)."""
import ctypes as C
import json
import struct

import numpy as np
import pytest

from ipu_ray_lib_b200 import _capi as capi
from ipu_ray_lib_b200.keras_h5 import model_config, write_keras_h5
from ipu_ray_lib_b200.nif import DenseLayer, NifWeights


def same(a: NifWeights, b: NifWeights):
    assert len(a.layers) == len(b.layers)
    for x, y in zip(a.layers, b.layers):
        assert x.kernel.shape == y.kernel.shape and x.kernel.tobytes() == y.kernel.tobytes()
        assert (x.bias is None) == (y.bias is None) and x.relu == y.relu
        if x.bias is not None:
            assert x.bias.tobytes() == y.bias.tobytes()


@pytest.mark.parametrize("float32,vlen", [(False, False), (True, False), (False, True)])
def test_round_trip_of_the_headline_model(tmp_path, float32, vlen):
    w = NifWeights.synthetic(seed=1442)
    p = tmp_path / "converted.hdf5"
    write_keras_h5(p, w, float32=float32, vlen_config=vlen)
    assert p.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    back = NifWeights.load(p, metadata_path=capi.REPO_ROOT / "assets/nif/urban_alley_01_4k_fp16_yuv/assets.extra/nif_metadata.txt")
    same(w, back)  # fp16 -> (fp32) -> fp16 is exact
    assert back.embedding_dimension == 12 and abs(back.max - 3.4299468994140625) < 1e-6 and back.log_tone_map
    assert [l.kernel.shape for l in back.layers] == [(48, 320), (320, 320), (320, 320), (368, 320), (320, 320), (320, 320), (320, 3)]


def test_biasless_layers_many_groups_and_float32_rounding(tmp_path):
    """More layers than one symbol-table node holds (the B-tree has several leaves), a layer without bias, and float32
    weights that are NOT representable in fp16: the loader rounds to nearest even like numpy."""
    rng = np.random.default_rng(5)
    w = NifWeights(embedding_dimension=4)
    width = 16
    for i in range(11):
        out = 3 if i == 10 else 24
        w.layers.append(DenseLayer(rng.standard_normal((width, out)).astype(np.float16),
                                   None if i % 3 == 1 else rng.standard_normal(out).astype(np.float16), i != 10))
        width = out
    p32 = tmp_path / "m32.h5"
    exact = [rng.standard_normal(l.kernel.shape).astype(np.float32) for l in w.layers]

    class Wide:  # same shapes, float32 payloads
        embedding_dimension = 4
        layers = [DenseLayer(k, l.bias, l.relu) for k, l in zip(exact, w.layers)]

    for l, k in zip(Wide.layers, exact):
        l.kernel = _F32(k)
    write_keras_h5(p32, Wide, float32=True)
    back = NifWeights.load(p32, metadata_path=tmp_path / "none.txt")
    for l, k in zip(back.layers, exact):
        assert l.kernel.tobytes() == k.astype(np.float16).tobytes()
    assert [l.bias is None for l in back.layers] == [i % 3 == 1 for i in range(11)]
    assert back.embedding_dimension == 4


class _F32(np.ndarray):
    """float32 array whose astype(float32) keeps full precision (write_keras_h5 casts with astype(dt))."""

    def __new__(cls, a):
        return np.asarray(a, np.float32).view(cls)


def test_model_config_rules_of_the_reference_loader(tmp_path):
    w = NifWeights.synthetic(seed=3, hidden=32, hidden_layers=2, concat_at=1)
    p = tmp_path / "m.h5"
    write_keras_h5(p, w)
    lib = capi.scene_lib()
    h = C.c_void_p()
    assert lib.b200rt_keras_hdf5_open(str(p).encode(), C.byref(h)) == 0
    assert lib.b200rt_keras_hdf5_version(h) == b"2.9.0" and lib.b200rt_keras_hdf5_num_layers(h) == 3
    kl = capi.KerasLayer()
    assert lib.b200rt_keras_hdf5_layer(h, 1, C.byref(kl)) == 0
    assert kl.name == b"dense_1" and kl.activation == b"relu" and kl.layer.in_features == 32 + 48  # the concat layer
    assert lib.b200rt_keras_hdf5_layer(h, 2, C.byref(kl)) == 0 and kl.activation == b"linear" and kl.layer.relu == 0
    assert lib.b200rt_keras_hdf5_layer(h, 3, C.byref(kl)) != 0
    lib.b200rt_keras_hdf5_close(h)

    # a Sequential model, and a layer class the reference rejects (Hdf5Model.cpp:17-20, :44-50)
    raw = p.read_bytes()
    cfg = json.dumps(model_config(w.layers, w.embedding_dimension, ["dense", "dense_1", "dense_2"], "float16")).encode()
    assert raw.count(cfg) == 1
    for old, new, msg in [(b'"class_name": "Functional"', b'"class_name": "Sequential"', b"Functional"),
                          (b'"class_name": "Concatenate"', b'"class_name": "BatchNormal"', b"not supported by Hdf5Model loader")]:
        assert len(old) == len(new)
        q = tmp_path / "bad.h5"
        q.write_bytes(raw.replace(cfg, cfg.replace(old, new, 1)))
        assert lib.b200rt_keras_hdf5_open(str(q).encode(), C.byref(h)) != 0
        assert msg in lib.b200rt_keras_last_error()


def test_malformed_files_fail_cleanly(tmp_path):
    lib = capi.scene_lib()
    h = C.c_void_p()
    w = NifWeights.synthetic(seed=3, hidden=32, hidden_layers=2, concat_at=1)
    p = tmp_path / "m.h5"
    write_keras_h5(p, w)
    raw = bytearray(p.read_bytes())
    cases = {"not hdf5": b"PK\x03\x04" + bytes(600), "empty": b"", "truncated": bytes(raw[:len(raw) // 2]),
             "root header points past the end": bytes(raw[:64]) + struct.pack("<Q", len(raw) + 4096) + bytes(raw[72:])}
    # every dataset renamed: the lookup of kernel:0 fails
    cases["missing dataset"] = bytes(raw).replace(b"kernel:0", b"kernel:9")
    for name, data in cases.items():
        q = tmp_path / "bad.h5"
        q.write_bytes(data)
        assert lib.b200rt_keras_hdf5_open(str(q).encode(), C.byref(h)) != 0, name
        assert lib.b200rt_keras_last_error().startswith(b"hdf5:"), name
    assert lib.b200rt_keras_hdf5_open(str(tmp_path / "absent.h5").encode(), C.byref(h)) != 0
    rng = np.random.default_rng(0)
    for _ in range(200):  # random corruption never crashes the reader
        bad = bytearray(raw)
        for k in rng.integers(0, len(bad), 8):
            bad[k] = rng.integers(0, 256)
        q = tmp_path / "fuzz.h5"
        q.write_bytes(bytes(bad))
        if lib.b200rt_keras_hdf5_open(str(q).encode(), C.byref(h)) == 0:
            lib.b200rt_keras_hdf5_close(h)
