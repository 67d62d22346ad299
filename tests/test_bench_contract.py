"""The committed bench lines (profiles/r0N_bench.json, written by `python bench.py` on a B200) carries every key of the
measurement contract, and the reference arm prints the same line shape on the CPU."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]


def _check(line, reference=False):
    for k in REQUIRED:
        assert k in line, k
    assert line["metric"] == "path_trace_mrays_per_s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    if reference:
        assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0
        assert line["cpu_baseline"]["kind"] in ("reference", "port")
    else:
        assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        assert line["roofline"]["bound"] in ("hbm", "tensor", "issue")
        assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
        assert line["gpu_launches"] > 0 and line["warmup"] >= 3
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
        assert 0 < line["e2e"]["value"] <= line["value"] * 1.05


@pytest.mark.parametrize("name", ["r01_bench.json", "r02_bench.json"])
def test_committed_bench_line_has_the_contract_keys(name):
    text = (ROOT / "profiles" / name).read_text().strip().splitlines()[-1]
    _check(json.loads(text))


def test_reference_arm_prints_the_same_line_shape():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in [k for k in REQUIRED if k not in ("clocks", "roofline")]:
        assert k in line, k
    _check({**line, "clocks": {}, "roofline": {}}, reference=True)
