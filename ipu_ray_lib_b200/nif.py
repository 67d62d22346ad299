"""NIF environment-light weights: container, synthetic generator and file format.

The reference loads a Keras h5 file (src/keras/Hdf5Model.cpp) that is missing from its checkout
(.MISSING_LARGE_BLOBS). This module keeps the layer list data-driven exactly like the reference (a list
of Dense layers: fp16 kernel [in,out], optional fp16 bias, relu/linear) and provides
  * ``NifWeights.synthetic`` — fixed-seed random weights in the architecture the shipped metadata
    implies (embedding 12, hidden 320, 6 hidden layers with the encoded input concatenated into the
    4th, 3 linear outputs), and
  * ``NifWeights.load`` / ``save`` — the reference's own container, a Keras h5 model file, read by the C++ reader of
    ``host/keras_hdf5.cpp`` (shared with ``B200Scene::loadNifModel``) and written by ``keras_h5.py``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import _capi as capi


@dataclass
class DenseLayer:
    kernel: np.ndarray            # [in, out] float16
    bias: np.ndarray | None       # [out] float16
    relu: bool


@dataclass
class NifWeights:
    embedding_dimension: int
    layers: list = field(default_factory=list)
    max: float = 3.4299468994140625
    mean: tuple = (-2.3514461517333984 - 1e-8, -2.2660605907440186 - 1e-8, -1.9648972749710083 - 1e-8)
    log_tone_map: bool = True

    @classmethod
    def synthetic(cls, seed: int = 1442, embedding_dimension: int = 12, hidden: int = 320,
                  hidden_layers: int = 6, concat_at: int = 3, scale: float = 1.0) -> "NifWeights":
        rng = np.random.default_rng(seed)
        feat = 4 * embedding_dimension
        w = cls(embedding_dimension=embedding_dimension)
        width = feat
        for i in range(hidden_layers):
            k_in = width + (feat if i == concat_at else 0)
            kern = (rng.standard_normal((k_in, hidden)) * (scale * np.sqrt(2.0 / k_in))).astype(np.float16)
            bias = (rng.standard_normal(hidden) * 0.05).astype(np.float16)
            w.layers.append(DenseLayer(kern, bias, True))
            width = hidden
        kern = (rng.standard_normal((width, 3)) * (0.5 / np.sqrt(width))).astype(np.float16)
        bias = (rng.standard_normal(3) * 0.1).astype(np.float16)
        w.layers.append(DenseLayer(kern, bias, False))
        return w

    @classmethod
    def from_metadata(cls, metadata_path, seed: int = 1442, **kw) -> "NifWeights":
        """Synthetic weights shaped by a nif_metadata.txt (IpuScene::loadNifModel, src/IpuScene.cpp:174-187)."""
        md = capi.NifMetadata()
        rc = capi.scene_lib().b200rt_read_nif_metadata(str(metadata_path).encode(), C.byref(md))
        if rc != 0:
            raise RuntimeError(capi.scene_lib().b200rt_scene_last_error().decode())
        w = cls.synthetic(seed=seed, embedding_dimension=md.embedding_dimension, hidden=md.hidden_size or 320, **kw)
        w.max = float(md.max)
        w.mean = tuple(float(x) for x in md.mean)
        w.log_tone_map = bool(md.log_tone_map)
        return w

    def save(self, path, *, float32: bool = False) -> None:
        """Write the layer list as a Keras h5 model (the container the reference loads, src/keras/Hdf5Model.cpp)."""
        from .keras_h5 import write_keras_h5

        write_keras_h5(path, self, float32=float32)

    @classmethod
    def load(cls, path, metadata_path=None) -> "NifWeights":
        """Load a Keras h5 model through the C++ reader (host/keras_hdf5.cpp). Decode parameters come from
        ``nif_metadata.txt`` (default: next to the file), as in IpuScene::loadNifModel (src/IpuScene.cpp:174-187)."""
        lib = capi.scene_lib()
        h = C.c_void_p()
        if lib.b200rt_keras_hdf5_open(str(path).encode(), C.byref(h)) != 0:
            raise RuntimeError(lib.b200rt_keras_last_error().decode())
        try:
            layers = []
            for i in range(lib.b200rt_keras_hdf5_num_layers(h)):
                kl = capi.KerasLayer()
                if lib.b200rt_keras_hdf5_layer(h, i, C.byref(kl)) != 0:
                    raise RuntimeError(lib.b200rt_keras_last_error().decode())
                L = kl.layer
                k = np.ctypeslib.as_array(C.cast(L.kernel_f16, C.POINTER(C.c_uint16)), (L.in_features, L.out_features))
                kern = k.view(np.float16).copy()
                bias = None
                if L.bias_f16:
                    bias = np.ctypeslib.as_array(C.cast(L.bias_f16, C.POINTER(C.c_uint16)), (L.out_features,)).view(np.float16).copy()
                layers.append(DenseLayer(kern, bias, bool(L.relu)))
        finally:
            lib.b200rt_keras_hdf5_close(h)
        md_path = Path(metadata_path) if metadata_path else Path(path).with_name("nif_metadata.txt")
        if md_path.exists():
            md = capi.NifMetadata()
            if lib.b200rt_read_nif_metadata(str(md_path).encode(), C.byref(md)) != 0:
                raise RuntimeError(lib.b200rt_scene_last_error().decode())
            w = cls(embedding_dimension=md.embedding_dimension, max=float(md.max), mean=tuple(float(x) for x in md.mean),
                    log_tone_map=bool(md.log_tone_map))
        else:
            w = cls(embedding_dimension=layers[0].kernel.shape[0] // 4)
        w.layers = layers
        return w

    def flops_per_sample(self) -> int:
        """sum(2*K*N + N_bias) as NifModel::analyseModel (src/neural_networks/NifModel.cpp:123-145)."""
        return sum(2 * l.kernel.shape[0] * l.kernel.shape[1] + (l.kernel.shape[1] if l.bias is not None else 0)
                   for l in self.layers)

    def weight_stream_bytes_per_tile(self) -> int:
        """fp16 bytes of weight images one 128-row tile of the tensor-core kernel streams from L2 (csrc/nif_tc.cuh): per
        layer (K rounded up to 16 + one 16-row bias slice) x (N padded to 16, hidden widths to 160 or 320) x 2 B."""
        total = 0
        for l in self.layers:
            k, n = l.kernel.shape
            kpad = (k + 15) // 16 * 16 + 16
            npad = (n + 15) // 16 * 16
            if npad > 16:
                npad = 160 if npad <= 160 else 320
            total += kpad * npad * 2
        return total

    def to_desc(self):
        """(b200rt_nif_desc, keepalive) for the C ABI / oracle."""
        n = len(self.layers)
        arr = (capi.NifLayer * n)()
        keep = [arr]
        for i, l in enumerate(self.layers):
            k = np.ascontiguousarray(l.kernel, dtype=np.float16)
            keep.append(k)
            arr[i].in_features, arr[i].out_features = k.shape
            arr[i].kernel_f16 = k.ctypes.data
            if l.bias is not None:
                b = np.ascontiguousarray(l.bias, dtype=np.float16)
                keep.append(b)
                arr[i].bias_f16 = b.ctypes.data
            else:
                arr[i].bias_f16 = None
            arr[i].relu = int(l.relu)
        d = capi.NifDesc()
        d.embedding_dimension = self.embedding_dimension
        d.num_layers = n
        d.layers = C.cast(arr, C.POINTER(capi.NifLayer))
        d.max = self.max
        d.mean[0], d.mean[1], d.mean[2] = self.mean
        d.log_tone_map = int(self.log_tone_map)
        return d, keep
