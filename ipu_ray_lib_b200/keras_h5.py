"""Writer for the Keras "h5" layout the reference loads its NIF from (``<assets.extra>/converted.hdf5``,
src/IpuScene.cpp:177 -> src/keras/Hdf5Model.cpp:61-85): root attributes ``keras_version`` / ``backend`` /
``model_config`` (JSON of a "Functional" model) and ``/model_weights/<layer>/<layer>/{kernel:0,bias:0}``.

h5py / libhdf5 are not available in this environment, so the file is assembled directly from the HDF5 File Format
Specification 3.0 in the "classic" layout h5py emits by default (superblock version 0, version-1 object headers,
symbol-table groups with a v1 B-tree + local heap, contiguous little-endian float datasets, scalar string
attributes). It is the single on-disk container of NIF weights for both the C++ loader
(host/keras_hdf5.cpp, used by ``B200Scene::loadNifModel`` and ``NifWeights.load``) and Python.
"""
from __future__ import annotations

import json
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class _Image:
    def __init__(self):
        self.buf = bytearray(96)  # superblock written last

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        at = len(self.buf)
        self.buf += data
        return at

    def patch(self, at: int, data: bytes):
        self.buf[at:at + len(data)] = data


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(img: _Image, messages: list[bytes], continuation: list[bytes] | None = None) -> int:
    """Version-1 object header; `continuation` messages go to a second block reached through message 0x0010."""
    first = list(messages)
    cont_at = None
    if continuation:
        block = b"".join(continuation)
        cont_at = img.alloc(block)
        first.append(_message(0x0010, struct.pack("<QQ", cont_at, len(block))))
    body = b"".join(first)
    n = len(first) + (len(continuation) if continuation else 0)
    return img.alloc(struct.pack("<BxHII4x", 1, n, 1, len(body)) + body)


def _dataspace(dims) -> bytes:
    return struct.pack("<BBBx4x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)


def _float_type(size: int) -> bytes:
    # class 1 (floating point), version 1; little-endian, IEEE: sign position in bits 8-15 of the class bit field
    if size == 2:
        sign, exp_loc, exp_size, man_size, bias = 15, 10, 5, 10, 15
    else:
        sign, exp_loc, exp_size, man_size, bias = 31, 23, 8, 23, 127
    bits = 0x20 | (sign << 8)  # mantissa normalisation = implied msb, byte order LE, padding 0
    return struct.pack("<B3BI", 0x11, bits & 0xff, (bits >> 8) & 0xff, 0, size) + \
        struct.pack("<HHBBBBI", 0, size * 8, exp_loc, exp_size, 0, man_size, bias)


def _string_type(size: int) -> bytes:
    return struct.pack("<B3BI", 0x13, 0x00, 0, 0, size)  # class 3, null-terminated ASCII


def _vlen_string_type() -> bytes:
    # class 9 variable length, type = string (1), base type = 1-byte string
    return struct.pack("<B3BI", 0x19, 0x01, 0, 0, 16) + _string_type(1)


def _attribute(name: str, type_msg: bytes, data: bytes) -> bytes:
    nm = name.encode() + b"\0"
    space = _dataspace(())
    return _message(0x000C, struct.pack("<BxHHH", 1, len(nm), len(type_msg), len(space)) + _pad8(nm) + _pad8(type_msg) +
                    _pad8(space) + data)


def _group(img: _Image, children: dict[str, int], extra_messages=(), continuation=None, leaf_k: int = 2) -> int:
    """Old-style group: local heap with the names, symbol nodes of at most 2 * leaf_k entries under one B-tree node."""
    names = sorted(children)
    heap = bytearray(b"\0" * 8)  # offset 0 = empty string
    offsets = {}
    for n in names:
        offsets[n] = len(heap)
        heap += _pad8(n.encode() + b"\0")
    heap += b"\0" * 16  # one free block
    seg_at = img.alloc(bytes(heap))
    free_off = len(heap) - 16
    img.patch(seg_at + free_off, struct.pack("<QQ", 1, 16))
    heap_at = img.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, seg_at))
    snods, keys = [], [0]
    per = 2 * leaf_k
    for i in range(0, max(len(names), 1), per):
        part = names[i:i + per]
        body = b"SNOD" + struct.pack("<BxH", 1, len(part))
        for n in part:
            body += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
        body += b"\0" * ((per - len(part)) * 40)
        snods.append(img.alloc(body))
        keys.append(offsets[part[-1]] if part else 0)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", keys[0])
    for s, k in zip(snods, keys[1:]):
        tree += struct.pack("<QQ", s, k)
    tree_at = img.alloc(tree)
    msgs = [_message(0x0011, struct.pack("<QQ", tree_at, heap_at))] + list(extra_messages)
    return _object_header(img, msgs, continuation)


def _dataset(img: _Image, array: np.ndarray) -> int:
    a = np.ascontiguousarray(array)
    assert a.dtype in (np.float16, np.float32)
    data_at = img.alloc(a.astype(a.dtype.newbyteorder("<")).tobytes())
    msgs = [_message(0x0001, _dataspace(a.shape)), _message(0x0003, _float_type(a.dtype.itemsize)),
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_at, a.nbytes))]
    return _object_header(img, msgs)


def model_config(layers, embedding_dimension: int, names, dtype: str) -> dict:
    """A Keras Functional config in the shape the NIF trainer saves: InputLayer, Dense..., one Concatenate."""
    feat = 4 * embedding_dimension
    cfg = [{"class_name": "InputLayer", "config": {"batch_input_shape": [None, feat], "dtype": dtype, "name": "input_1"},
            "name": "input_1", "inbound_nodes": []}]
    width, prev = feat, "input_1"
    for name, l in zip(names, layers):
        k_in, k_out = l.kernel.shape
        if k_in == width + feat:
            cat = f"concatenate_{name}"
            cfg.append({"class_name": "Concatenate", "config": {"name": cat, "axis": -1, "dtype": dtype}, "name": cat,
                        "inbound_nodes": [[[prev, 0, 0, {}], ["input_1", 0, 0, {}]]]})
            prev = cat
        cfg.append({"class_name": "Dense",
                    "config": {"name": name, "trainable": True, "dtype": dtype, "units": int(k_out),
                               "activation": "relu" if l.relu else "linear", "use_bias": l.bias is not None},
                    "name": name, "inbound_nodes": [[[prev, 0, 0, {}]]]})
        prev, width = name, k_out
    return {"class_name": "Functional",
            "config": {"name": "nif", "layers": cfg, "input_layers": [["input_1", 0, 0]], "output_layers": [[prev, 0, 0]]},
            "keras_version": "2.9.0", "backend": "tensorflow"}


def write_keras_h5(path, weights, *, float32: bool = False, vlen_config: bool = False) -> None:
    """Write ``weights`` (an :class:`ipu_ray_lib_b200.nif.NifWeights`) as a Keras h5 model file.

    float32: store the datasets as float32 (the loader rounds them to fp16); vlen_config: store ``model_config`` as a
    variable-length string through the global heap (what h5py does for ``str`` values) instead of a fixed-length one.
    """
    img = _Image()
    names = ["dense"] + [f"dense_{i}" for i in range(1, len(weights.layers))]
    dt = np.float32 if float32 else np.float16
    layer_groups = {}
    for name, l in zip(names, weights.layers):
        inner = {"kernel:0": _dataset(img, l.kernel.astype(dt))}
        if l.bias is not None:
            inner["bias:0"] = _dataset(img, l.bias.astype(dt))
        layer_groups[name] = _group(img, {name: _group(img, inner)})
    weights_group = _group(img, layer_groups)
    cfg = json.dumps(model_config(weights.layers, weights.embedding_dimension, names, "float32" if float32 else "float16")).encode()

    def fixed(name, value: bytes):
        return _attribute(name, _string_type(len(value) + 1), _pad8(value + b"\0"))

    if vlen_config:
        obj = struct.pack("<HH4xQ", 1, 1, len(cfg)) + _pad8(cfg)
        free = struct.pack("<HH4xQ", 0, 0, 0)
        size = 16 + len(obj) + len(free)
        gcol = img.alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + obj + free)
        cfg_attr = _attribute("model_config", _vlen_string_type(), struct.pack("<IQI", len(cfg), gcol, 1))
    else:
        cfg_attr = fixed("model_config", cfg)
    root = _group(img, {"model_weights": weights_group},
                  extra_messages=[fixed("keras_version", b"2.9.0"), fixed("backend", b"tensorflow")],
                  continuation=[cfg_attr])
    eof = len(img.buf)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 2, 16, 0) + \
        struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + struct.pack("<QQII16x", 0, root, 0, 0)
    assert len(sb) == 96
    img.patch(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(img.buf))
