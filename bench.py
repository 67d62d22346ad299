#!/usr/bin/env python
"""bench.py — benchmark of the trace path (BASELINE.json: path-trace Mrays/s & samples/s at 1/2/4/8 B200 vs CPU).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path, BASELINE configs[1] (the headline)
    python bench.py --config {1,2,3,4,5} ...                   # the other named configurations (1-based, see CONFIGS)
    python bench.py --impl reference --gpus N --steps K ...     # the reference's own CPU kernels on the host cores

A STEP is one complete render of the workload. Default = configs[1] of BASELINE.json: the built-in box scene,
path-traced with the NIF environment light at 1440x1440, 1000 spp (maxPathLength 10, roulette start 3, anti-alias 0.25,
seed 1442; synthetic fixed-seed NIF weights because the trained ones are missing from the reference checkout).
With N > 1 the same image is partitioned ray-data-parallel: the TraceResult stream is cut into the reference's ray
batches (8640 rays) and batch i goes to rank i % N (src/IpuScene.cpp:676-684); scene/BVH/NIF weights are replicated;
there is no collective on the data path, only the final framebuffer gather (rgb) to rank 0 over NCCL, inside the step.

Numbers on the JSON line:
  value   Mrays/s = BVH queries actually issued (closest-hit + occlusion, counted on the device) / s,
          rays already resident in HBM, device-timed with CUDA events on the launching stream, max over ranks.
  e2e     same metric through the public C-ABI call with HOST buffers: every step streams its shard of the ray stream
          host->device, renders, streams it back (b200rt_trace: tiles pipelined over three CUDA streams).
  samples_per_s  pixels * spp / s for the same timed region.
  roofline / roofline_kernels, cpu_baseline, clocks: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from ipu_ray_lib_b200 import HostScene, _capi as capi, init_ray_stream  # noqa: E402
from ipu_ray_lib_b200.nif import NifWeights  # noqa: E402
from ipu_ray_lib_b200.parallel import batch_owner_mask  # noqa: E402

METRIC = "path_trace_mrays_per_s"
UNIT = "Mrays/s"
RAYS_PER_BATCH = 8640

# BASELINE.json `configs`, 1-based. (steps, warmup) are the defaults when the command line gives none.
CONFIGS = {
    1: dict(label="configs[0]", scene="box", mesh=None, normals=False, mode="shadow", width=1440, height=1440, spp=1,
            nif=False, steps=20, warmup=3,
            what="built-in 'box' scene, --render-mode shadow-trace (traceShadowRay: closest hit + shadow ray), 1 spp"),
    2: dict(label="configs[1]", scene="box", mesh=None, normals=False, mode="path", width=1440, height=1440, spp=1000,
            nif=True, steps=3, warmup=3,
            what="built-in 'box' scene, path-trace RGB with NIF HDRI environment light (synthetic weights, seed 1442)"),
    3: dict(label="configs[2]", scene=None, mesh="assets/test_scene.dae", normals=True, mode="path", width=1440,
            height=1440, spp=4000, nif=False, steps=1, warmup=3,
            what="assets/test_scene.dae --load-normals (8474 triangles, glass / mirror / emitters; 407 KB BVH: L2-resident), path-trace RGB"),
    4: dict(label="configs[3]", scene="box", mesh=None, normals=False, mode="path", width=3840, height=2160, spp=256,
            nif=False, steps=2, warmup=3,
            what="assets/monkey_bust.glb inside the built-in 'box' scene (the GLB has no camera, SURVEY.md 8d), path-trace RGB, "
                 "reference default 256 spp"),
    5: dict(label="configs[4]", scene=None, mesh="assets/hdri_test.dae", normals=False, mode="path", width=8192,
            height=8192, spp=1000, nif=True, steps=1, warmup=3,
            what="assets/hdri_test.dae NIF-lit (open sky; synthetic weights, seed 1442), ray stream host-resident and "
                 "streamed through the GPU in tiles"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration, 1-based")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--samples", type=int, default=None, help="override the configuration's spp (labelled in `config`)")
    ap.add_argument("--scene", default=None)
    ap.add_argument("--no-nif", action="store_true", help="diagnostic only: drop the NIF environment light")
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--residency", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="samples per NIF wavefront chunk (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the bounded CPU sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.cfg = cfg
    args.steps = cfg["steps"] if args.steps is None else args.steps
    args.warmup = cfg["warmup"] if args.warmup is None else args.warmup
    args.width = cfg["width"] if args.width is None else args.width
    args.height = cfg["height"] if args.height is None else args.height
    args.spp_named = cfg["spp"]
    args.samples = cfg["spp"] if args.samples is None else args.samples
    args.mode = cfg["mode"]
    args.scene = cfg["scene"] if args.scene is None else args.scene
    args.use_nif = cfg["nif"] and not args.no_nif
    return args


def load_scene(args, device=-1):
    cfg = args.cfg
    if cfg["mesh"] and args.scene is None:
        s = HostScene.from_file(ROOT / cfg["mesh"], load_normals=cfg["normals"])
    else:
        s = HostScene.builtin(args.scene)
    return s.configure(args.width, args.height, path_trace=args.mode == "path", samples=max(args.samples, 1), seed=1442,
                       device=device)


def config_dict(args, n_gpus, reference_sample=None):
    cfg = args.cfg
    named = (args.width, args.height, args.samples) == (cfg["width"], cfg["height"], cfg["spp"])
    d = {
        "workload": f"{cfg['what']}, {args.width}x{args.height}, {args.samples} spp (BASELINE.json {cfg['label']}"
                    + ("" if named else f"; REDUCED from {cfg['width']}x{cfg['height']}, {cfg['spp']} spp") + ")",
        "baseline_config": cfg["label"], "scene": args.scene or cfg["mesh"], "mode": args.mode,
        "width": args.width, "height": args.height, "spp": args.samples,
        "max_path_length": 10, "roulette_start_depth": 3, "anti_alias": 0.25, "seed": 1442,
        "nif": bool(args.use_nif),
        "parallelism": f"ray-data-parallel: 8640-ray batches, batch i -> rank i % {n_gpus}, scene replicated",
        "l2_policy": ("inputs larger than L2: a fresh copy of the ray stream per step from a ring of pre-initialised "
                      "buffers (174 MB each vs 126 MB L2)" if args.mode == "shadow" else
                      "inputs larger than L2 (ray stream + >= 4 GB of path state per chunk vs 126 MB L2); no flush between steps"),
    }
    if reference_sample is not None:
        # the CPU arm renders a bounded sample of this workload with the reference's kernels; it has no NIF stage
        d["nif"] = False
        d["reference_sample"] = reference_sample
        d["workload"] += " -- CPU arm: bounded sample, see reference_sample; the reference CPU path has no NIF stage"
    return d


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        self.device_index = device_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        timed = [r for t, r in self.rows if t0 is not None and t1 is not None and t0 <= t <= t1]
        note = None
        if not timed and self.rows:
            # timed region shorter than one nvidia-smi period: fall back to the samples taken under the same load
            # during the warm-up steps that precede it
            timed = [r for _, r in self.rows]
            note = "timed region shorter than one sampling period: samples are from the warm-up steps (same load)"
        for r in timed:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """Per-launch DRAM traffic of the dominant kernels from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text())
    return {}


# ------------------------------------------------------------------------------------------------
def cpu_baseline(args, scene, nif, kind_pref=("port",), seconds=12.0):
    """The CPU checker timed on this box's host cores on a BOUNDED sample of the same workload."""
    from oracle import oracle_py

    kind = next((k for k in kind_pref if (k == "port" and oracle_py.have_port()) or (k == "reference" and oracle_py.have_ref())), None)
    if kind is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "no oracle library built"}
    orc = oracle_py.Oracle(kind)
    cores = os.cpu_count() or 1
    use_nif = nif if (kind == "port" and args.use_nif) else None
    w, h = args.width, args.height
    rng = np.random.default_rng(0)
    # bounded sample: rows of the frame (stratified over its height), generated directly as crop windows so that even
    # the 8192^2 stream never has to exist on the host for this leg
    def rows_sample(n_rows):
        picks = np.sort(rng.choice(h, size=min(n_rows, h), replace=False))
        return np.concatenate([init_ray_stream(w, h, scene.fov, window=(w, 1, 0, int(r))) for r in picks])

    if args.mode == "shadow":
        def run(n_rows, _spp):
            sel = rows_sample(n_rows)
            t0 = time.perf_counter()
            cnt = orc.shadow_trace(scene, sel, threads=cores)
            return time.perf_counter() - t0, cnt, sel.size
        what = "traceShadowRay"
    else:
        def run(n_rows, spp):
            sel = rows_sample(n_rows)
            t0 = time.perf_counter()
            cnt = orc.path_trace(scene, sel, first_sample=0, num_samples=spp, nif=use_nif, threads=cores)
            return time.perf_counter() - t0, cnt, sel.size
        what = "pathTrace"

    spp = 1 if args.mode == "shadow" else 4
    dt, cnt, npix = run(8, spp)  # probe
    per_row = max(dt, 1e-6) / 8
    rows = int(max(8, min(h, seconds / per_row)))
    if rows == h and args.mode != "shadow":  # fast host: keep every row and raise the sample count instead
        spp = int(max(4, min(args.samples, spp * seconds / (per_row * h))))
    dt, cnt, npix = run(rows, spp)
    q = cnt["closest_hit_queries"] + cnt["occlusion_queries"]
    samples = cnt["samples"] if args.mode != "shadow" else npix
    return {
        "value": q / dt / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
        "samples_per_s": samples / dt, "seconds": dt,
        "sample": f"{what} over {rows} of the {h} rows of the {w}x{h} frame (stratified) x {spp} spp = {samples} samples, "
                  f"{q} BVH queries in {dt:.2f} s; OpenMP dynamic schedule over rays"
                  + ("" if args.mode == "shadow" else ", per-(pixel,sample) RNG streams")
                  + ("; NIF evaluated in fp32 on the CPU for escaped rays" if use_nif is not None else
                     ("" if args.mode == "shadow" else
                      "; no NIF stage (the reference CPU path has none: escaped rays just stop, trace.cpp:171-174)")),
    }


def run_reference(args):
    """--impl reference: the reference's own CPU kernel sources (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = load_scene(args)
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals, base = [], None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args, scene, None, kind_pref=("reference", "port"), seconds=per_step)
        if i >= args.warmup:
            vals.append(base)
    v = float(np.mean([b["value"] for b in vals]))
    sps = float(np.mean([b["samples_per_s"] for b in vals]))
    shipped = as_shipped_cpu(args, scene) if (base["kind"] == "reference" and args.mode != "shadow") else None
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([b["seconds"] for b in vals])) * 1e3,
        "ms_per_full_step_extrapolated": args.width * args.height * max(args.samples, 1) / sps * 1e3,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus, reference_sample=base["sample"]),
        "samples_per_s": sps,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if shipped is not None:
        line["cpu_reference_as_shipped"] = shipped
    print(json.dumps(line))


def as_shipped_cpu(args, scene, seconds=4.0):
    """SURVEY.md 8d "(A)": the reference's CPU path-trace loop the way it ships -- one global generator, every draw inside
    `omp critical`, camera rays regenerated serially per sample (trace.cpp:236-245) -- timed on a small bounded sample.
    `value` of the reference arm stays the fair number (same kernels, one RNG stream per (pixel, sample), no lock)."""
    from oracle import oracle_py

    orc = oracle_py.Oracle("reference")
    cores = os.cpu_count() or 1
    w, h = args.width, args.height
    rng = np.random.default_rng(1)

    def run(n_rows, spp):
        picks = np.sort(rng.choice(h, size=min(n_rows, h), replace=False))
        sel = np.concatenate([init_ray_stream(w, h, scene.fov, window=(w, 1, 0, int(r))) for r in picks])
        t0 = time.perf_counter()
        cnt = orc.path_trace_as_shipped(scene, sel, spp, threads=cores)
        return time.perf_counter() - t0, cnt

    dt, cnt = run(4, 2)
    rows = int(max(4, min(h, 4 * seconds / max(dt, 1e-6))))
    dt, cnt = run(rows, 2)
    q = cnt["closest_hit_queries"]
    return {"value": q / dt / 1e6, "unit": UNIT, "cores": cores, "samples_per_s": cnt["samples"] / dt, "seconds": dt,
            "sample": f"{rows} of the {h} rows x 2 spp = {cnt['samples']} samples, {q} BVH queries in {dt:.2f} s; global RNG in "
                      f"omp critical(sample), serial camera-ray regeneration per sample, omp schedule(auto)"}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from ipu_ray_lib_b200.render import B200Scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.lib().b200rt_device_count() < 1:
        raise SystemExit("bench.py: no B200 visible; the trace path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, h, spp = args.width, args.height, args.samples
    shadow = args.mode == "shadow"
    scene = load_scene(args, device=local_rank)
    nif = NifWeights.synthetic(seed=1442) if args.use_nif else None
    n_total = w * h
    mine = batch_owner_mask(n_total, RAYS_PER_BATCH, world, rank)
    # this rank's shard of the stream, generated row by row so the full 8192^2 stream never sits on the host twice
    shard = np.concatenate([init_ray_stream(w, h, scene.fov, window=(w, r1 - r0, 0, r0))[mine[r0 * w:r1 * w]]
                            for r0, r1 in ((r, min(r + 256, h)) for r in range(0, h, 256))])
    n_local = shard.size
    shard_bytes = n_local * capi.TRACE_RESULT.itemsize

    g = B200Scene(scene)
    if nif is not None:
        g.load_nif_model(nif)
    params = dict(traversal=args.traversal, scene_residency=args.residency, samples_per_chunk=args.chunk)

    pristine = torch.from_numpy(shard.view(np.uint8)).cuda()
    # Single-pass (shadow-trace) steps consume their input, so every timed step gets its own pre-initialised copy of the
    # stream (up to 32; beyond that the copies are refreshed inside the timed region and the line says so).
    ring_n = min(max(args.steps, 1), 32) if shadow else 1
    works = [torch.empty_like(pristine) for _ in range(ring_n)]
    work = works[0]
    gather_list = None
    rgb_counts = [int(batch_owner_mask(n_total, RAYS_PER_BATCH, world, r).sum()) for r in range(world)]
    stream = torch.cuda.current_stream().cuda_stream

    def gather_rgb(buf):
        """Final framebuffer gather (rgb of every ray) to rank 0 over NCCL."""
        if world == 1:
            return
        # NCCL's gather wants equal sizes; shards differ by at most one ray batch, so every rank sends max(shard) rows
        # (its own rgb, zero-padded) and rank 0 keeps the first rgb_counts[r] rows of rank r's slice
        rows = max(rgb_counts)
        rgb = torch.zeros(rows, 12, dtype=torch.uint8, device="cuda")
        rgb[:n_local] = buf.view(n_local, 84)[:, :12]
        nonlocal gather_list
        if rank == 0 and gather_list is None:
            gather_list = [torch.empty(rows, 12, dtype=torch.uint8, device="cuda") for _ in rgb_counts]
        dist.gather(rgb, gather_list if rank == 0 else None, dst=0)

    totals = {"queries": 0, "samples": 0, "escaped": 0, "launches": 0, "kernel_ms": 0.0}

    def device_step(record, k=0):
        buf = works[k % ring_n]
        if not shadow or k >= ring_n:
            buf.copy_(pristine)
        g.execute_device(buf.data_ptr(), n_local, stream=stream, **params)
        gather_rgb(buf)
        if record:
            st = g.stats()
            totals["queries"] += st["closest_hit_queries"] + st["occlusion_queries"]
            totals["samples"] += st["samples"] if not shadow else n_local
            totals["escaped"] += st["escaped_samples"]
            totals["launches"] += st["kernel_launches"]
            totals["kernel_ms"] += st["kernel_ms"]

    # ---- device-resident arm ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:  # started before the warm-up so that it is running when the clock starts
        for _ in range(args.warmup):
            device_step(False)
        for b in works:
            b.copy_(pristine)
        barrier()
        clocks.mark_start()
        ev0.record()
        for k in range(args.steps):
            device_step(True, k)
        ev1.record()
        barrier()
        clocks.mark_end()
    dev_ms = ev0.elapsed_time(ev1)

    # ---- end-to-end arm: host buffers through b200rt_trace ----
    # Every timed step renders a freshly initialised ray stream that already sits in page-locked host memory (a ring of
    # streams prepared before the clock starts; rgb is a running sum, so a path-traced stream that comes round again
    # after the ring wraps is the same work). Inside the timed region: H2D of the stream, all kernels, D2H of the results.
    e2e_ring = max(1, min(args.steps, 32 if shadow else 4, int((8 << 30) // max(shard_bytes, 1))))
    ring = [torch.empty(n_local * 84, dtype=torch.uint8).pin_memory().numpy().view(capi.TRACE_RESULT) for _ in range(e2e_ring)]
    e2e_steps = args.steps if (not shadow or args.steps <= e2e_ring) else e2e_ring

    def e2e_step(buf):
        g.execute(buf, **params)
        if world > 1:
            work.view(n_local, 84)[:, :12].copy_(torch.from_numpy(buf.view(np.uint8).reshape(n_local, 84)[:, :12]).cuda())
            gather_rgb(work)

    ring[0][:] = shard
    e2e_step(ring[0])  # warm-up of the host-buffer path (staging buffers, first-touch)
    for buf in ring:
        buf[:] = shard
    barrier()
    t0 = time.perf_counter()
    e2e_queries = 0
    for k in range(e2e_steps):
        e2e_step(ring[k % len(ring)])
        st = g.stats()
        e2e_queries += st["closest_hit_queries"] + st["occlusion_queries"]
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    # ---- reduce over ranks: time = max, work = sum ----
    t = torch.tensor([dev_ms, e2e_ms, totals["kernel_ms"]], dtype=torch.float64, device="cuda")
    c = torch.tensor([totals["queries"], totals["samples"], totals["escaped"], totals["launches"], e2e_queries],
                     dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, kernel_ms = (float(x) for x in t.tolist())
    queries, samples, escaped, launches, e2e_queries = (int(x) for x in c.tolist())

    if rank == 0:
        peaks = measured_peaks()
        value = queries / (dev_ms * 1e-3) / 1e6
        e2e_value = e2e_queries / (e2e_ms * 1e-3) / 1e6
        # per-kernel breakdown from one extra instrumented step (not timed)
        prof = kernel_breakdown(g, work, pristine, n_local, stream, params, shadow)
        flops_per_lookup = nif.flops_per_sample() if nif is not None else 0
        nif_tile_bytes = nif.weight_stream_bytes_per_tile() if nif is not None else 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
            "samples_per_s": samples / (dev_ms * 1e-3),
            "bvh_queries_per_sample": queries / max(samples, 1), "escaped_fraction": escaped / max(samples, 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_total * 84, "d2h_bytes_per_step": n_total * 84,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "note": "ray stream in page-locked host memory; tiles pipelined H2D || kernels || D2H inside b200rt_trace"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "peaks": peaks,
        }
        if shadow and args.steps > ring_n:
            line["config"]["l2_policy"] += f"; steps beyond {ring_n} refresh their copy inside the timed region"
        line.update(rooflines(prof, peaks, n_local, spp, flops_per_lookup, line["clocks"].get("sm_mhz"), shadow, nif_tile_bytes))
        if nif is not None and not shadow:
            line["chunk_overlap"] = {
                "timed_steps": "library default (on): NIF + accumulate of chunk c on a second stream beside the trace / shade "
                               "kernels of chunk c + 1; bit-identical frame",
                "ms_per_step": dev_ms / args.steps,
                "roofline_step": "one extra untimed step with chunk_overlap off (kernels serialised): per-kernel event "
                                 "spans, shares and roofline fractions are the kernels' own",
                "ms_serialised_step": prof["kernel_ms"],
            }
        if not args.skip_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, scene, nif, kind_pref=("port",), seconds=args.cpu_seconds)
        elif world > 1:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


def kernel_breakdown(g, work, pristine, n_local, stream, params, shadow):
    """Device time of the trace kernel alone vs the whole step, with CUDA events on the launching stream."""
    import torch

    out = {}
    # one more full step: the library brackets every kernel launch with CUDA events on the launching stream. This step
    # runs with chunk_overlap off: in the timed steps (library default) the NIF kernel of chunk c shares the SMs with
    # the trace / shade kernels of chunk c + 1, so a kernel's event span there is not the kernel's own time. Serialised,
    # the spans add up to the step and agree with the ncu launch list (which serialises kernels too).
    work.copy_(pristine)
    g.execute_device(work.data_ptr(), n_local, stream=stream, **dict(params, chunk_overlap=1))
    st = g.stats()
    for k in ("kernel_ms", "trace_kernel_ms", "nif_kernel_ms", "accumulate_kernel_ms", "shade_kernel_ms", "kernel_launches",
              "trace_kernel_launches", "nif_kernel_launches", "shade_kernel_launches", "escaped_samples", "samples"):
        out[k] = st[k]
    out["queries"] = st["closest_hit_queries"] + st["occlusion_queries"]
    # work counters (node visits / primitive tests) from a reduced-spp instrumented run, scaled per query
    work.copy_(pristine)
    extra = {} if shadow else {"num_samples": 8}
    g.execute_device(work.data_ptr(), n_local, stream=stream, **dict(params, count_visits=1, **extra))
    sc = g.stats()
    q8 = max(sc["closest_hit_queries"] + sc["occlusion_queries"], 1)
    out["node_visits_per_query"] = sc["node_visits"] / q8
    out["prim_tests_per_query"] = sc["prim_tests"] / q8
    torch.cuda.synchronize()
    return out


def rooflines(prof, peaks, n_local, spp, flops_per_lookup, sm_mhz=None, shadow=False, nif_tile_bytes=0):
    """Three regimes: ray-stream HBM, traversal issue rate, NIF tensor cores."""
    import torch

    step_s = prof["kernel_ms"] * 1e-3
    trace_s = max(prof["trace_kernel_ms"] * 1e-3, 1e-9)
    nif_s = max(prof["nif_kernel_ms"] * 1e-3, 1e-9)
    traffic = ncu_traffic()
    # (i) ray streams: 84 B read + 84 B written per ray per render (the reference's streaming contract)
    stream_bytes = 168.0 * n_local
    # (ii) traversal: fp32-pipe lane-ops per query = V*33 + T*100 (SURVEY.md §8d weights; fmad off, so mul and add
    #      issue separately); ceiling = SMs x 4 schedulers x 32 lanes x SM clock
    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    sm_clock_hz = (sm_mhz or 1965.0) * 1e6  # median SM clock sampled by nvidia-smi during the timed region
    lane_ops = prof["queries"] * (prof["node_visits_per_query"] * 33.0 + prof["prim_tests_per_query"] * 100.0)
    issue_peak = props.multi_processor_count * 4 * 32 * sm_clock_hz
    # (iii) NIF: sum(2KN + N) flops per escaped sample (NifModel.cpp:123-145)
    nif_flops = float(prof["escaped_samples"]) * flops_per_lookup
    nl = max(prof["nif_kernel_launches"], 1)
    tl = max(prof["trace_kernel_launches"], 1)
    wavefront = prof.get("shade_kernel_launches", 0) > 0
    trace_name = "shadow_trace_kernel" if shadow else ("wf_trace_kernel" if wavefront else "path_trace_kernel")
    tr = traffic.get(trace_name) or {}
    unit_note = "B per step, summed over the step's launches (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)"
    kernels = [
        {"kernel": trace_name, "bound": "issue", "share_of_step": trace_s / step_s,
         "launches_per_step": prof["trace_kernel_launches"], "avg_launch_ms": trace_s * 1e3 / tl,
         "algorithmic_lane_ops_per_launch": lane_ops / tl, "achieved": lane_ops / trace_s / 1e12,
         "peak": issue_peak / 1e12, "unit": "Tlane-op/s", "frac": lane_ops / trace_s / issue_peak,
         "node_visits_per_query": prof["node_visits_per_query"], "prim_tests_per_query": prof["prim_tests_per_query"],
         "sm_clock_mhz_for_peak": sm_clock_hz / 1e6, "traffic": tr.get("dram_bytes_per_step"), "traffic_unit": unit_note,
         "note": ("single pass over the TraceResult stream: closest hit + shadow ray per camera ray" if shadow else
                  "one launch per bounce and chunk; launches of late bounces are nearly empty" if wavefront else
                  "one launch per chunk: camera ray, traversal and shading of every bounce")},
    ]
    if not shadow:
        kernels.append(
            {"kernel": "nif_mlp_kernel", "bound": "tensor", "share_of_step": nif_s / step_s if nif_flops else 0.0,
             "launches_per_step": prof["nif_kernel_launches"], "avg_launch_ms": nif_s * 1e3 / nl,
             "algorithmic_flops_per_launch": nif_flops / nl, "achieved": nif_flops / nif_s / 1e12 if nif_flops else 0.0,
             "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
             "frac": nif_flops / nif_s / 1e12 / peaks["bf16_tflops_sustained"] if nif_flops else 0.0,
             "traffic": (traffic.get("nif_mlp_kernel") or {}).get("dram_bytes_per_step"), "traffic_unit": unit_note})
        if nif_flops and nif_tile_bytes:
            # the kernel's other resource: every CTA streams the whole weight set from L2 once per 128-row tile. Ceiling:
            # the L2 -> SM read rate, ~6300 B/cycle chip-wide (B300_MICROARCH.md "LTS throughput cap"). The kernel runs
            # close to it, but is not bound by it alone (cycles per tile do not change with fewer CTAs
            # sharing the L2 or with fewer bytes per tile).
            tiles = float(prof["escaped_samples"]) / 128.0
            l2_bytes = tiles * nif_tile_bytes
            l2_peak = 6300.0 * sm_clock_hz  # B/s
            kernels.append({"kernel": "nif_mlp_kernel (weight stream L2 -> SM)", "bound": "l2",
                            "weight_bytes_per_128_row_tile": nif_tile_bytes, "algorithmic_bytes_per_step": l2_bytes,
                            "achieved": l2_bytes / nif_s / 1e9, "peak": l2_peak / 1e9, "unit": "GB/s",
                            "frac": l2_bytes / nif_s / l2_peak,
                            "peak_source": "6300 B/cycle x sampled SM clock (B300_MICROARCH.md LTS cap; no measured B200 figure)"})
    if wavefront:
        # wf_shade: HBM-bound on the path record (dense, slot-indexed). Algorithmic bytes: every bounce query beyond a
        # path's first (queries - samples of them; each is also exactly one survivor of the previous bounce) reads 72 B
        # (8 B hit + origin, direction, throughput, RNG) and was written as a 96 B record (+ 1/d and shear constants of
        # the next query); every camera path reads its 8 B hit and writes the 12 B colour slot; every path ends once
        # with a 20 B throughput / escape record.
        shade_s = max(prof["shade_kernel_ms"] * 1e-3, 1e-9)
        bounce_q = prof["queries"] - prof["samples"]
        shade_bytes = bounce_q * (72.0 + 96.0) + prof["samples"] * (8.0 + 12.0 + 20.0)
        sl = max(prof["shade_kernel_launches"], 1)
        kernels.append({"kernel": "wf_shade_kernel", "bound": "hbm", "share_of_step": shade_s / step_s,
                        "launches_per_step": prof["shade_kernel_launches"], "avg_launch_ms": shade_s * 1e3 / sl,
                        "algorithmic_bytes_per_step": shade_bytes, "achieved": shade_bytes / shade_s / 1e9,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": shade_bytes / shade_s / 1e9 / peaks["hbm_gbs"],
                        "traffic": (traffic.get("wf_shade_kernel") or {}).get("dram_bytes_per_step"), "traffic_unit": unit_note})
    kernels.append(
        {"kernel": "TraceResult stream in/out", "bound": "hbm", "algorithmic_bytes_per_step": stream_bytes,
         "achieved": stream_bytes / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
         "frac": stream_bytes / step_s / 1e9 / peaks["hbm_gbs"],
         "note": ("168 B/ray/render; the single-pass kernel is bound by traversal issue slots, not by this stream "
                  "(see the issue entry)" if shadow else
                  "168 B/ray/render amortised over all spp: HBM is idle by design in a multi-sample render")})
    dominant = kernels[0] if (shadow or trace_s >= nif_s or not nif_flops) else kernels[1]
    roof = {"kernel": dominant["kernel"], "bound": dominant["bound"], "achieved": dominant["achieved"],
            "peak": dominant["peak"], "unit": dominant["unit"], "frac": dominant["frac"],
            "traffic": dominant.get("traffic"), "peak_source": peaks["source"],
            "share_of_step": dominant["share_of_step"]}
    return {"roofline": roof, "roofline_kernels": kernels}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
