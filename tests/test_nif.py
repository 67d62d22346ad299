"""NIF environment light: CPU restatement sanity (not gpu) and GPU-vs-oracle parity within a stated tolerance (gpu).

Tolerance statement: features, weights and layer outputs are fp16 on both sides; the oracle accumulates each
dot product in fp32 in k order, the GPU in tile order (tensor cores) — pre-rounding sums differ by ~1e-6
relative, which flips an fp16 rounding of a layer output now and then (1 fp16 ulp = 1e-3 relative of ONE of 320
activations). Measured on a B200 over 100 k random lookups of the headline model (scripts/nif_tolerance.py,
profiles/r02_nif_tolerance.txt): max relative error 3.4e-3, mean 1.6e-4, 99.9th percentile 1.9e-3. The tests require
max < 1.5e-2 and mean < 8e-4 (about 5x the measurement).
The part nothing can pin (the reference has no CPU NIF, and the IPU accumulated in fp16 partials,
src/IpuScene.cpp:256-262) is bounded with a model of that option: test_fp16_partials_model_bounds_the_unpinned_gap.
"""
import numpy as np
import pytest

from ipu_ray_lib_b200 import _capi as capi, scene
from ipu_ray_lib_b200.nif import NifWeights

MAX_REL, MEAN_REL = 1.5e-2, 8e-4


def numpy_nif(w: NifWeights, uv):
    """Independent numpy statement of the same forward pass (fp64 accumulation) — loose cross-check of the oracle."""
    uv = np.asarray(uv, np.float32)
    E = w.embedding_dimension
    un = ((uv - np.float32(1)) * np.float32(2)).astype(np.float32)
    coeff = (2.0 ** np.arange(E)).astype(np.float32)
    a = (un[:, :, None] * coeff).astype(np.float16).astype(np.float32)  # [n, 2, E]
    feat = np.concatenate([np.sin(a[:, 0]), np.sin(a[:, 1]), np.cos(a[:, 0]), np.cos(a[:, 1])], axis=1).astype(np.float16)
    x = feat.astype(np.float64)
    for l in w.layers:
        if x.shape[1] != l.kernel.shape[0]:
            x = np.concatenate([x, feat.astype(np.float64)], axis=1)
        y = x @ l.kernel.astype(np.float64)
        if l.bias is not None:
            y = y + l.bias.astype(np.float64)
        if l.relu:
            y = np.maximum(y, 0)
        x = y.astype(np.float16).astype(np.float64)
    y = x * w.max + np.asarray(w.mean)
    return np.exp(y) if w.log_tone_map else y


def test_synthetic_model_shape_matches_shipped_metadata():
    w = NifWeights.from_metadata(capi.REPO_ROOT / "assets/nif/urban_alley_01_4k_fp16_yuv/assets.extra/nif_metadata.txt")
    assert w.embedding_dimension == 12
    assert [l.kernel.shape for l in w.layers] == [(48, 320), (320, 320), (320, 320), (368, 320), (320, 320), (320, 320), (320, 3)]
    assert w.flops_per_sample() == 1089283  # ~1.09 MFLOP / sample
    assert abs(w.max - 3.4299468994140625) < 1e-6 and w.log_tone_map


def test_oracle_nif_against_numpy(port):
    w = NifWeights.synthetic(seed=7)
    uv = np.random.default_rng(0).uniform(0, 1, (3000, 2)).astype(np.float32)
    got = port.nif_eval(w, uv)
    want = numpy_nif(w, uv)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < MAX_REL and rel.mean() < MEAN_REL
    assert np.all(np.isfinite(got)) and got.min() > 0


def test_fp16_partials_model_bounds_the_unpinned_gap(port):
    """The reference ran the MLP with partialsType = half on the IPU (src/IpuScene.cpp:256-262); this build accumulates
    in fp32 (TMEM). The oracle can model fp16 partials (running sum rounded to fp16 every 16 products): the two
    contracts differ by ~2e-3 mean / ~1.4e-2 max relative in the decoded radiance -- that is the size of what stays
    unpinned, an order of magnitude above the GPU-vs-oracle error and far below Monte-Carlo noise at 1000 spp."""
    w = NifWeights.synthetic(seed=1442)
    uv = np.random.default_rng(2).uniform(0, 1, (20000, 2)).astype(np.float32)
    a, b = port.nif_eval(w, uv), port.nif_eval_partials(w, uv, half_chunk=16)
    rel = np.abs(a - b) / np.abs(a)
    assert 5e-4 < rel.mean() < 4e-3 and rel.max() < 3e-2
    assert np.array_equal(port.nif_eval_partials(w, uv, half_chunk=0), a)


def test_oracle_nif_is_data_driven(port):
    """Bias-less layers, no skip connection, linear decode (log_tone_map off) all follow the layer list."""
    rng = np.random.default_rng(1)
    from ipu_ray_lib_b200.nif import DenseLayer
    w = NifWeights(embedding_dimension=4, max=2.0, mean=(0.1, 0.2, 0.3), log_tone_map=False)
    w.layers = [DenseLayer((rng.standard_normal((16, 32)) * 0.3).astype(np.float16), None, True),
                DenseLayer((rng.standard_normal((32, 3)) * 0.3).astype(np.float16), None, False)]
    uv = rng.uniform(0, 1, (500, 2)).astype(np.float32)
    got, want = port.nif_eval(w, uv), numpy_nif(w, uv)
    assert np.allclose(got, want, rtol=2e-2, atol=2e-3)


def test_dir_to_uv(port):
    d = np.float32([[0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1]])
    uv = port.dir_to_uv(d)
    assert np.allclose(uv[:, 0], [0, 1, .5, .5, .5, .5], atol=1e-6)          # u = acos(y)/pi
    assert np.allclose(uv[2:, 1], [0, .25, .5, .75], atol=1e-6)              # v = atan2(z,x)/2pi wrapped to [0,1]
    rot = port.dir_to_uv(d[2:3], rotation_radians=np.pi / 2)
    assert abs(rot[0, 1] - 0.25) < 1e-6


@pytest.mark.gpu
def test_gpu_nif_eval_matches_oracle(port, box_scene):
    from ipu_ray_lib_b200.render import B200Scene
    w = NifWeights.synthetic(seed=1442)
    rng = np.random.default_rng(2)
    uv = rng.uniform(0, 1, (100_000, 2)).astype(np.float32)
    uv[:7] = [(0, 0), (1, 1), (0.5, 0.5), (1, 0), (0, 1), (0.999999, 0.25), (0.25, 1e-7)]
    want = port.nif_eval(w, uv)
    with B200Scene(box_scene.configure(64, 64)) as g:
        g.load_nif_model(w)
        for n in (1, 127, 128, 129, 5000, 100_000):  # ragged M tiles
            got = g.nif_eval(uv[:n])
            rel = np.abs(got - want[:n]) / np.abs(want[:n])
            assert rel.max() < MAX_REL and rel.mean() < MEAN_REL, (n, rel.max(), rel.mean())


@pytest.mark.gpu
def test_gpu_nif_other_architectures(port, box_scene):
    from ipu_ray_lib_b200.nif import DenseLayer
    from ipu_ray_lib_b200.render import B200Scene
    rng = np.random.default_rng(3)
    uv = rng.uniform(0, 1, (20_000, 2)).astype(np.float32)
    small = NifWeights.synthetic(seed=5, hidden=64, hidden_layers=3, concat_at=1)
    nobias = NifWeights(embedding_dimension=12, max=1.5, mean=(0.0, 0.1, 0.2), log_tone_map=False)
    nobias.layers = [DenseLayer((rng.standard_normal((48, 128)) * 0.2).astype(np.float16), None, True),
                     DenseLayer((rng.standard_normal((128, 128)) * 0.12).astype(np.float16), None, True),
                     DenseLayer((rng.standard_normal((128, 3)) * 0.1).astype(np.float16), None, False)]
    odd = NifWeights.synthetic(seed=6, hidden=200, hidden_layers=4, concat_at=2)   # zero-padded to 320 columns
    narrow = NifWeights.synthetic(seed=7, hidden=100, hidden_layers=2, concat_at=1)  # zero-padded to 160 columns
    ragged = NifWeights.synthetic(seed=8, hidden=77, hidden_layers=3, concat_at=1)   # not a multiple of 16
    with B200Scene(box_scene.configure(64, 64)) as g:
        for w in (small, nobias, odd, narrow, ragged):
            g.load_nif_model(w)
            got, want = g.nif_eval(uv), port.nif_eval(w, uv)
            assert np.allclose(got, want, rtol=MAX_REL, atol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunk", [("spheres", 0), ("spheres", 3), ("box", 5)])
def test_nif_lit_path_trace_matches_oracle(port, name, chunk):
    """Hit records are bit-exact (the NIF does not feed back into the paths); rgb within the NIF tolerance."""
    from ipu_ray_lib_b200.render import B200Scene
    w, h, spp = 80, 64, 7
    s = scene.HostScene.builtin(name).configure(w, h, path_trace=True, samples=spp, seed=31)
    nif = NifWeights.synthetic(seed=1442)
    base = scene.init_ray_stream(w, h, s.fov)
    want = base.copy()
    cw = port.path_trace(s, want, nif=nif, hdri_rotation=110.0)
    with B200Scene(s) as g:
        g.load_nif_model(nif)
        g.set_hdri_rotation(110.0)
        got = base.copy()
        g.execute(got, samples_per_chunk=chunk)
        st = g.stats()
        g.set_max_nif_batch_size(1000)  # IpuScene::setMaxNifBatchSize: serial NIF launches of <= 1000 rays, same image
        again = base.copy()
        g.execute(again, samples_per_chunk=chunk)
        assert again.tobytes() == got.tobytes() and g.stats()["nif_kernel_launches"] > st["nif_kernel_launches"]
    assert st["escaped_samples"] == cw["escaped_samples"] and st["samples"] == cw["samples"]
    a = got.copy(); b = want.copy()
    a["rgb"] = 0; b["rgb"] = 0
    assert a.tobytes() == b.tobytes(), "hit records must stay bit-exact with the NIF enabled"
    num = np.abs(got["rgb"].astype(np.float64) - want["rgb"]).sum()
    assert num / np.abs(want["rgb"]).sum() < MEAN_REL
    # per-pixel: allow the rare fp16 argument flip caused by libm-vs-CUDA acos/atan2 ulp differences
    rel = np.abs(got["rgb"] - want["rgb"]).max(axis=1) / np.maximum(np.abs(want["rgb"]).max(axis=1), 1e-6)
    assert np.mean(rel > MAX_REL) < 5e-3


@pytest.mark.gpu
def test_config2_nif_lit_window_of_the_full_frame(port):
    """BASELINE.json config 2 WITH its environment light on a 1440 x 48 window of the 1440 x 1440 frame (seed 1442,
    synthetic weights of seed 1442, the bench's model): hit records bit-exact, rgb within the stated NIF tolerance,
    escaped-sample count equal to the oracle's."""
    from ipu_ray_lib_b200.render import B200Scene
    spp = 6
    s = scene.HostScene.builtin("box").configure(1440, 1440, path_trace=True, samples=spp, seed=1442)
    nif = NifWeights.synthetic(seed=1442)
    base = scene.init_ray_stream(1440, 1440, s.fov, window=(1440, 48, 0, 696))
    want = base.copy()
    cw = port.path_trace(s, want, nif=nif)
    with B200Scene(s) as g:
        g.load_nif_model(nif)
        got = base.copy()
        g.execute(got)
        st = g.stats()
    assert st["escaped_samples"] == cw["escaped_samples"] and st["closest_hit_queries"] == cw["closest_hit_queries"]
    a = got.copy(); b = want.copy()
    a["rgb"] = 0; b["rgb"] = 0
    assert a.tobytes() == b.tobytes()
    assert np.abs(got["rgb"].astype(np.float64) - want["rgb"]).sum() / np.abs(want["rgb"]).sum() < MEAN_REL
    rel = np.abs(got["rgb"] - want["rgb"]).max(axis=1) / np.maximum(np.abs(want["rgb"]).max(axis=1), 1e-6)
    assert np.mean(rel > MAX_REL) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spp,chunk,residency", [("box", 96, 80, 13, 2, 0), ("spheres", 80, 64, 9, 1, 0),
                                                        ("box", 200, 120, 12, 4, 1), ("box", 64, 48, 5, 8, 0)])
def test_chunk_overlap_is_bit_identical(name, w, h, spp, chunk, residency):
    """b200rt_trace_params.chunk_overlap: the NIF + accumulate of chunk c on a second stream beside the trace / shade
    kernels of chunk c + 1 (two sets of per-sample records, accumulates in chunk order) leaves the same bytes in the
    stream as the serialised chunks -- every field, rgb included -- for many small chunks, an odd chunk count, a single
    chunk (nothing to overlap), an explicit shared-memory residency and serial NIF micro-batches."""
    from ipu_ray_lib_b200.render import B200Scene
    s = scene.HostScene.builtin(name).configure(w, h, path_trace=True, samples=spp, seed=77)
    nif = NifWeights.synthetic(seed=1442)
    base = scene.init_ray_stream(w, h, s.fov)
    with B200Scene(s) as g:
        g.load_nif_model(nif)
        g.set_hdri_rotation(35.0)
        serial = base.copy()
        g.execute(serial, samples_per_chunk=chunk, chunk_overlap=1, scene_residency=residency)
        launches = g.stats()["kernel_launches"]
        for mode in (2, 0):  # on, auto
            got = base.copy()
            g.execute(got, samples_per_chunk=chunk, chunk_overlap=mode, scene_residency=residency)
            assert got.tobytes() == serial.tobytes(), f"chunk_overlap={mode}"
            assert g.stats()["kernel_launches"] == launches
        g.set_max_nif_batch_size(500)
        got = base.copy()
        g.execute(got, samples_per_chunk=chunk, chunk_overlap=2, scene_residency=residency)
        assert got.tobytes() == serial.tobytes(), "chunk_overlap with NIF micro-batches"
        # sample ranges compose across calls with the overlap on
        half = base.copy()
        g.execute(half, samples_per_chunk=chunk, chunk_overlap=2, first_sample=0, num_samples=spp // 2)
        g.execute(half, samples_per_chunk=chunk, chunk_overlap=2, first_sample=spp // 2, num_samples=spp - spp // 2)
        assert half["rgb"].tobytes() == serial["rgb"].tobytes()
