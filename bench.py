#!/usr/bin/env python
"""bench.py — headline benchmark of the trace path (BASELINE.json: path-trace Mrays/s & samples/s, 1440^2, built-in scene).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU kernels on the host cores

A STEP is one complete render of the workload: configs[1] of BASELINE.json — the built-in box scene,
path-traced with the NIF environment light at 1440x1440, 1000 spp (maxPathLength 10, roulette start 3,
anti-alias 0.25, seed 1442; synthetic fixed-seed NIF weights because the trained ones are missing from
the reference checkout). With N > 1 the same image is partitioned ray-data-parallel: the TraceResult
stream is cut into the reference's ray batches (8640 rays) and batch i goes to rank i % N
(src/IpuScene.cpp:676-684); scene/BVH/NIF weights are replicated; there is no collective on the data
path, only the final framebuffer gather (rgb) to rank 0 over NCCL, which is inside the timed step.

Numbers on the JSON line:
  value   Mrays/s = BVH queries actually issued (closest-hit + occlusion, counted on the device) / s,
          rays already resident in HBM, device-timed with CUDA events on the launching stream, max over ranks.
  e2e     same metric through the public C-ABI call with HOST buffers: every step copies its shard of
          the ray stream host->device, renders, copies it back (b200rt_trace).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=1440)
    ap.add_argument("--height", type=int, default=1440)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--scene", default="box")
    ap.add_argument("--no-nif", action="store_true", help="diagnostic only: drop the NIF environment light")
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--residency", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="samples per NIF wavefront chunk (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the bounded CPU sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {
        "workload": f"built-in '{args.scene}' scene, path-trace RGB"
                    + ("" if args.no_nif else " with NIF HDRI environment light (synthetic weights, seed 1442)")
                    + f", {args.width}x{args.height}, {args.samples} spp (BASELINE.json configs[1])",
        "scene": args.scene, "width": args.width, "height": args.height, "spp": args.samples,
        "max_path_length": 10, "roulette_start_depth": 3, "anti_alias": 0.25, "seed": 1442,
        "nif": not args.no_nif,
        "parallelism": f"ray-data-parallel: 8640-ray batches, batch i -> rank i % {n_gpus}, scene replicated",
        "l2_policy": "inputs larger than L2 (174 MB ray stream per image vs 126 MB L2); no flush between steps",
    }


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
    use_nif = nif if (kind == "port" and not args.no_nif) else None
    w, h = args.width, args.height
    full = init_ray_stream(w, h, scene.fov)
    # bounded sample: every `stride`-th pixel of the full image (stratified over the whole frame), `spp` samples each
    rng = np.random.default_rng(0)

    def run(stride, spp):
        sel = np.ascontiguousarray(full[rng.integers(0, stride)::stride])
        t0 = time.perf_counter()
        cnt = orc.path_trace(scene, sel, first_sample=0, num_samples=spp, nif=use_nif, threads=cores)
        return time.perf_counter() - t0, cnt, sel.size

    dt, cnt, npix = run(256, 2)  # probe (~16k samples)
    rate = cnt["samples"] / max(dt, 1e-6)
    target = max(int(rate * seconds), 20000)
    spp = 4
    stride = max(1, int(w * h * spp / target))
    if stride == 1:  # fast host: keep every pixel and raise the sample count instead
        spp = int(max(4, min(args.samples, target // (w * h))))
    dt, cnt, npix = run(stride, spp)
    q = cnt["closest_hit_queries"] + cnt["occlusion_queries"]
    return {
        "value": q / dt / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
        "samples_per_s": cnt["samples"] / dt, "seconds": dt,
        "sample": f"{npix} pixels (every {stride}th of the {w}x{h} frame) x {spp} spp = {cnt['samples']} samples, "
                  f"{q} BVH queries in {dt:.2f} s; OpenMP dynamic schedule over rays, per-(pixel,sample) RNG streams"
                  + ("; NIF evaluated in fp32 on the CPU for escaped rays" if use_nif is not None else
                     "; no NIF stage (the reference CPU path has none: escaped rays just stop, trace.cpp:171-174)"),
    }


def run_reference(args):
    """--impl reference: the reference's own CPU kernel sources (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = HostScene.builtin(args.scene).configure(args.width, args.height, path_trace=True, samples=args.samples)
    nif = NifWeights.synthetic(seed=1442)
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals, base = [], None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args, scene, nif, kind_pref=("reference", "port"), seconds=per_step)
        if i >= args.warmup:
            vals.append(base)
    v = float(np.mean([b["value"] for b in vals]))
    sps = float(np.mean([b["samples_per_s"] for b in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([b["seconds"] for b in vals])) * 1e3,
        "ms_per_full_step_extrapolated": args.width * args.height * args.samples / sps * 1e3,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, args.gpus),
        "samples_per_s": sps,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    scene = HostScene.builtin(args.scene).configure(w, h, path_trace=True, samples=spp, seed=1442, device=local_rank)
    nif = None if args.no_nif else NifWeights.synthetic(seed=1442)
    full = init_ray_stream(w, h, scene.fov)
    mine = batch_owner_mask(full.size, RAYS_PER_BATCH, world, rank)
    shard = np.ascontiguousarray(full[mine])
    n_local = shard.size
    shard_bytes = n_local * capi.TRACE_RESULT.itemsize

    g = B200Scene(scene)
    if nif is not None:
        g.load_nif_model(nif)
    params = dict(traversal=args.traversal, scene_residency=args.residency, samples_per_chunk=args.chunk)

    pristine = torch.from_numpy(shard.view(np.uint8)).cuda()
    work = torch.empty_like(pristine)
    pinned = torch.from_numpy(shard.view(np.uint8).copy()).pin_memory()
    pinned_np = pinned.numpy().view(capi.TRACE_RESULT)
    gather_list = None
    rgb_counts = [int(batch_owner_mask(full.size, RAYS_PER_BATCH, world, r).sum()) for r in range(world)]
    stream = torch.cuda.current_stream().cuda_stream

    def gather_rgb():
        """Final framebuffer gather (rgb of every ray) to rank 0 over NCCL."""
        if world == 1:
            return
        rgb = work.view(n_local, 84)[:, :12].contiguous()
        nonlocal gather_list
        if rank == 0 and gather_list is None:
            gather_list = [torch.empty(c, 12, dtype=torch.uint8, device="cuda") for c in rgb_counts]
        dist.gather(rgb, gather_list if rank == 0 else None, dst=0)

    totals = {"queries": 0, "samples": 0, "escaped": 0, "launches": 0, "kernel_ms": 0.0}

    def device_step(record):
        work.copy_(pristine)
        g.execute_device(work.data_ptr(), n_local, stream=stream, **params)
        gather_rgb()
        if record:
            st = g.stats()
            totals["queries"] += st["closest_hit_queries"] + st["occlusion_queries"]
            totals["samples"] += st["samples"]
            totals["escaped"] += st["escaped_samples"]
            totals["launches"] += st["kernel_launches"]
            totals["kernel_ms"] += st["kernel_ms"]

    # ---- device-resident arm ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:  # started before the warm-up so that it is running when the clock starts
        for _ in range(args.warmup):
            device_step(False)
        barrier()
        clocks.mark_start()
        ev0.record()
        for _ in range(args.steps):
            device_step(True)
        ev1.record()
        barrier()
        clocks.mark_end()
    dev_ms = ev0.elapsed_time(ev1)

    # ---- end-to-end arm: host buffers through b200rt_trace ----
    # Every timed step renders a freshly initialised ray stream that already sits in page-locked host memory (a ring of
    # up to 4 streams prepared before the clock starts; rgb is a running sum, so a stream that comes round again after
    # the ring wraps is the same work). Inside the timed region: H2D of the stream, all kernels, D2H of the results.
    ring = [pinned_np] + [torch.empty(n_local * 84, dtype=torch.uint8).pin_memory().numpy().view(capi.TRACE_RESULT)
                          for _ in range(min(args.steps, 4) - 1)]

    def e2e_step(buf):
        g.execute(buf, **params)
        if world > 1:
            work.view(n_local, 84)[:, :12].copy_(torch.from_numpy(buf.view(np.uint8).reshape(n_local, 84)[:, :12]).cuda())
            gather_rgb()

    pinned_np[:] = shard
    e2e_step(pinned_np)  # warm-up of the host-buffer path (staging buffers, first-touch)
    for buf in ring:
        buf[:] = shard
    barrier()
    t0 = time.perf_counter()
    e2e_queries = 0
    for k in range(args.steps):
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
        prof = kernel_breakdown(g, work, pristine, n_local, stream, params, nif)
        flops_per_lookup = nif.flops_per_sample() if nif is not None else 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
            "samples_per_s": samples / (dev_ms * 1e-3),
            "bvh_queries_per_sample": queries / max(samples, 1), "escaped_fraction": escaped / max(samples, 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": shard_bytes * 1 if world == 1 else full.size * 84,
                    "d2h_bytes_per_step": shard_bytes * 1 if world == 1 else full.size * 84, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "peaks": peaks,
        }
        line.update(rooflines(prof, peaks, n_local, spp, flops_per_lookup, line["clocks"].get("sm_mhz")))
        if not args.skip_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, scene, nif, kind_pref=("port",), seconds=args.cpu_seconds)
        elif world > 1:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


def kernel_breakdown(g, work, pristine, n_local, stream, params, nif):
    """Device time of the trace kernel alone vs the whole step, with CUDA events on the launching stream."""
    import torch

    out = {}
    # one more full step: the library brackets every kernel launch with CUDA events on the launching stream
    work.copy_(pristine)
    g.execute_device(work.data_ptr(), n_local, stream=stream, **params)
    st = g.stats()
    for k in ("kernel_ms", "trace_kernel_ms", "nif_kernel_ms", "accumulate_kernel_ms", "shade_kernel_ms", "kernel_launches",
              "trace_kernel_launches", "nif_kernel_launches", "shade_kernel_launches", "escaped_samples", "samples"):
        out[k] = st[k]
    out["queries"] = st["closest_hit_queries"] + st["occlusion_queries"]
    # work counters (node visits / primitive tests) from a reduced-spp instrumented run, scaled per query
    work.copy_(pristine)
    g.execute_device(work.data_ptr(), n_local, stream=stream, **dict(params, count_visits=1, num_samples=8))
    sc = g.stats()
    q8 = max(sc["closest_hit_queries"] + sc["occlusion_queries"], 1)
    out["node_visits_per_query"] = sc["node_visits"] / q8
    out["prim_tests_per_query"] = sc["prim_tests"] / q8
    torch.cuda.synchronize()
    return out


def rooflines(prof, peaks, n_local, spp, flops_per_lookup, sm_mhz=None):
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
    trace_name = "wf_trace_kernel" if wavefront else "path_trace_kernel"
    tr = traffic.get(trace_name) or {}
    unit_note = "B per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)"
    kernels = [
        {"kernel": trace_name, "bound": "issue", "share_of_step": trace_s / step_s,
         "launches_per_step": prof["trace_kernel_launches"], "avg_launch_ms": trace_s * 1e3 / tl,
         "algorithmic_lane_ops_per_launch": lane_ops / tl, "achieved": lane_ops / trace_s / 1e12,
         "peak": issue_peak / 1e12, "unit": "Tlane-op/s", "frac": lane_ops / trace_s / issue_peak,
         "node_visits_per_query": prof["node_visits_per_query"], "prim_tests_per_query": prof["prim_tests_per_query"],
         "sm_clock_mhz_for_peak": sm_clock_hz / 1e6, "traffic": tr.get("dram_bytes_per_launch"), "traffic_unit": unit_note,
         "note": ("one launch per bounce and chunk; launches of late bounces are nearly empty" if wavefront else
                  "one launch per chunk: camera ray, traversal and shading of every bounce")},
        {"kernel": "nif_mlp_kernel", "bound": "tensor", "share_of_step": nif_s / step_s if nif_flops else 0.0,
         "launches_per_step": prof["nif_kernel_launches"], "avg_launch_ms": nif_s * 1e3 / nl,
         "algorithmic_flops_per_launch": nif_flops / nl, "achieved": nif_flops / nif_s / 1e12 if nif_flops else 0.0,
         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
         "frac": nif_flops / nif_s / 1e12 / peaks["bf16_tflops_sustained"] if nif_flops else 0.0,
         "traffic": (traffic.get("nif_mlp_kernel") or {}).get("dram_bytes_per_launch"), "traffic_unit": unit_note},
    ]
    if wavefront:
        # wf_shade: HBM-bound on the path record. Algorithmic bytes per path-bounce: 84 B read (queue id, hit, origin,
        # direction, throughput, RNG; only the 16 B hit at bounce 0, whose camera ray is recomputed) + 68 B written for
        # the ~70 % that survive (+ the 12 B colour slot at bounce 0).
        shade_s = max(prof["shade_kernel_ms"] * 1e-3, 1e-9)
        shade_bytes = (prof["queries"] - prof["samples"]) * 84.0 + prof["samples"] * (16.0 + 12.0) + prof["queries"] * 0.7 * 68.0
        sl = max(prof["shade_kernel_launches"], 1)
        kernels.append({"kernel": "wf_shade_kernel", "bound": "hbm", "share_of_step": shade_s / step_s,
                        "launches_per_step": prof["shade_kernel_launches"], "avg_launch_ms": shade_s * 1e3 / sl,
                        "algorithmic_bytes_per_launch": shade_bytes / sl, "achieved": shade_bytes / shade_s / 1e9,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": shade_bytes / shade_s / 1e9 / peaks["hbm_gbs"],
                        "traffic": (traffic.get("wf_shade_kernel") or {}).get("dram_bytes_per_launch"), "traffic_unit": unit_note})
    kernels.append(
        {"kernel": "TraceResult stream in/out", "bound": "hbm", "algorithmic_bytes_per_step": stream_bytes,
         "achieved": stream_bytes / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
         "frac": stream_bytes / step_s / 1e9 / peaks["hbm_gbs"],
         "note": "168 B/ray/render amortised over all spp: HBM is idle by design in a multi-sample render"})
    dominant = kernels[0] if trace_s >= nif_s or not nif_flops else kernels[1]
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
