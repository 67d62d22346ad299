"""Measured GPU-vs-oracle NIF error (what tests/test_nif.py's tolerances are derived from)."""
import sys, numpy as np
sys.path.insert(0, '.')
from ipu_ray_lib_b200 import HostScene
from ipu_ray_lib_b200.nif import NifWeights, DenseLayer
from ipu_ray_lib_b200.render import B200Scene
from oracle.oracle_py import Oracle
port = Oracle("port")
rng = np.random.default_rng(2)
uv = rng.uniform(0, 1, (100_000, 2)).astype(np.float32)
models = {"320x6 (headline)": NifWeights.synthetic(seed=1442), "64x3": NifWeights.synthetic(seed=5, hidden=64, hidden_layers=3, concat_at=1),
          "200x4 (padded to 320)": NifWeights.synthetic(seed=6, hidden=200, hidden_layers=4, concat_at=2),
          "100x2 (padded to 160)": NifWeights.synthetic(seed=7, hidden=100, hidden_layers=2, concat_at=1)}
with B200Scene(HostScene.builtin("box").configure(64, 64)) as g:
    for name, w in models.items():
        g.load_nif_model(w)
        got, want = g.nif_eval(uv), port.nif_eval(w, uv)
        rel = np.abs(got - want) / np.abs(want)
        half = port.nif_eval_partials(w, uv, half_chunk=16)
        relh = np.abs(got - half) / np.abs(half)
        print(f"{name}: GPU vs fp32-accumulate oracle max rel {rel.max():.3e} mean {rel.mean():.3e} p99.9 {np.quantile(rel, .999):.3e} | "
              f"GPU vs fp16-partials model max {relh.max():.3e} mean {relh.mean():.3e}")
