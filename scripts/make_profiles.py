"""Turns the raw ncu output of scripts/gpu_prof.sh (gpurun_out/, scratch) into the summaries committed under profiles/:
   <tag>_launches.csv          the launch list itself (ncu --metrics gpu__time_duration.sum,dram bytes,... --clock-control none)
   <tag>_launches_summary.txt  per kernel: launches, total / average ms, share, DRAM bytes, lanes, issue-active
   <tag>_ncu_full_metrics.json key counters of the --set full captures (bounce-1 wf_trace / wf_shade, NIF kernel)
   traffic.json                DRAM bytes per STEP and kernel (sum over a step's launches), read by bench.py
   <tag>_sass_opcodes.txt      opcode histogram of the product kernels (cuobjdump -sass of the built library)
Usage: python scripts/make_profiles.py r02 [spp_of_launch_list=128] [spp_of_step=1000]"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
spp_list = float(sys.argv[2]) if len(sys.argv) > 2 else 128.0
spp_step = float(sys.argv[3]) if len(sys.argv) > 3 else 1000.0
out = ROOT / "profiles"
src = ROOT / "gpurun_out"


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("rt::", "")


# ---- launch list ----
rows = [r for r in csv.reader(open(src / f"{tag}_launches.csv")) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ix = {n: i for i, n in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows:
    launch.setdefault(r[ix["ID"]], {"kernel": short(r[ix["Kernel Name"]])})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
unit_ns = {r[ix["Metric Name"]]: r[ix["Metric Unit"]] for r in rows}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit_ns.get("gpu__time_duration.sum", "ns"), 1e-6)
agg = collections.OrderedDict()
for l in launch.values():
    a = agg.setdefault(l["kernel"], collections.Counter())
    a["n"] += 1
    a["ms"] += l.get("gpu__time_duration.sum", 0.0) * tscale
    a["rd"] += l.get("dram__bytes_read.sum", 0.0)
    a["wr"] += l.get("dram__bytes_write.sum", 0.0)
    a["lanes_w"] += l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0) * l.get("gpu__time_duration.sum", 0.0)
    a["issue_w"] += l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * l.get("gpu__time_duration.sum", 0.0)
    a["t"] += l.get("gpu__time_duration.sum", 0.0)
total = sum(a["ms"] for a in agg.values())
shutil.copy(src / f"{tag}_launches.csv", out / f"{tag}_launches.csv")
with open(out / f"{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 1 --warmup 0 --samples {int(spp_list)} --skip-cpu-baseline` with B200RT_OVERLAP=0 (chunks serialised, like bench.py's roofline step; {len(launch)} launches captured: bounce-0 to bounce-4 "
            f"trace / shade pairs, the tail launch, NIF and accumulate of each 32-spp chunk), --clock-control none; per-launch times under ncu are serialised and cold-cache:\n"
            f"# the SHARES are what must agree with bench.py's event timing, not the absolute times.\n"
            f"# kernel, launches, total ms, share, avg ms, DRAM read GB, DRAM write GB, lanes per instruction (time-weighted), issue-active % (time-weighted)\n")
    for k, a in agg.items():
        f.write(f"{k}, {a['n']}, {a['ms']:.2f}, {100 * a['ms'] / total:.1f}%, {a['ms'] / a['n']:.3f}, {a['rd'] / 1e9:.2f}, {a['wr'] / 1e9:.2f}, "
                f"{a['lanes_w'] / max(a['t'], 1):.1f}, {a['issue_w'] / max(a['t'], 1):.1f}\n")
print(open(out / f"{tag}_launches_summary.txt").read())

# ---- traffic per step ----
# The captured run holds more launches than one step (bench.py renders again for its per-kernel breakdown): cut the list
# into chunks (a chunk starts with the bounce-0 trace kernel and ends with wf_accumulate) and average over complete chunks.
chunks, cur_chunk = [], None
for l in launch.values():
    first = "wf_trace_kernel" in l["kernel"] and l["kernel"].rstrip(">").endswith("1")
    if first:
        cur_chunk = []
    if cur_chunk is not None:
        cur_chunk.append(l)
        if "wf_accumulate" in l["kernel"]:
            chunks.append(cur_chunk)
            cur_chunk = None
chunks_per_step = spp_step / 32.0
group = {"wf_trace_kernel": "wf_trace", "wf_shade_kernel": "wf_shade", "wf_tail_kernel": "wf_tail", "nif_mlp_kernel": "nif_mlp", "wf_accumulate_kernel": "wf_accumulate"}
traffic = {}
for name, pat in group.items():
    per_chunk = [sum(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in c if pat in l["kernel"]) for c in chunks]
    n_launch = [sum(1 for l in c if pat in l["kernel"]) for c in chunks]
    if not per_chunk:
        continue
    traffic[name] = {"dram_bytes_per_step": sum(per_chunk) / len(per_chunk) * chunks_per_step,
                     "dram_bytes_per_chunk": sum(per_chunk) / len(per_chunk),
                     "launches_per_chunk": sum(n_launch) / len(n_launch),
                     "source": f"profiles/{tag}_launches.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of a 32-spp "
                               f"chunk (mean of {len(chunks)} complete chunks), times the {chunks_per_step:g} chunks of a {int(spp_step)}-spp step"}
(out / "traffic.json").write_text(json.dumps(traffic, indent=1))
print({k: (round(v["dram_bytes_per_step"] / 1e9, 1), v["launches_per_chunk"]) for k, v in traffic.items()})

# ---- full captures ----
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
full = {}
for rep in sorted(src.glob(f"{tag}_full_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    d = dict(zip(r[0], r[2]))
    units = dict(zip(r[0], r[1]))
    full[rep.stem] = {"kernel": short(d.get("Kernel Name", "")), "capture": f"ncu --set full --clock-control none --import-source on, launch 2 of the kernel in a 32-spp chunk (gpurun_out/{rep.name}, scratch)"}
    for k in want:
        if k in d and d[k] not in ("", "n/a"):
            full[rep.stem][k] = {"value": d[k], "unit": units.get(k, "")}
    # everything tensor-pipe related the report has
    for k in d:  # the tensor-pipe activity counters the report has (names differ between ncu versions)
        if re.search(r"pipe_tensor.*(cycles_active|inst_executed).*(avg\.pct_of_peak_sustained_active|\.sum)$", k) and k not in full[rep.stem] \
                and d[k] not in ("", "n/a", "0"):
            full[rep.stem][k] = {"value": d[k], "unit": units.get(k, "")}
(out / f"{tag}_ncu_full_metrics.json").write_text(json.dumps(full, indent=1))
for k, v in full.items():
    print(k, {m: x["value"] for m, x in list(v.items())[:12] if isinstance(x, dict)})

# ---- SASS opcode histogram ----
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "ipu_ray_lib_b200" / "libb200rt.so")], capture_output=True, text=True).stdout
hist = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = short(cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
with open(out / f"{tag}_sass_opcodes.txt", "w") as f:
    f.write("# cuobjdump -sass ipu_ray_lib_b200/libb200rt.so: opcode counts per kernel (sm_100a). UTC*MMA / LDTM / UBLKCP are the\n"
            "# tcgen05 MMA, TMEM load and TMA bulk-copy instructions; UBLKCP also appears in shadow_stream_kernel (TMA-staged ray tiles).\n")
    for k, c in hist.items():
        keys = [o for o in c if re.match(r"UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|HMMA|FMNMX3|VOTE|LDS|STL|LDL|ATOM", o)]
        f.write(f"{k}: {sum(c.values())} instructions; " + ", ".join(f"{o} {c[o]}" for o in sorted(keys)) + "\n")
print("sass kernels:", len(hist))
