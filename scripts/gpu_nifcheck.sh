#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/nifcheck.log 2>&1
timeout 600 python -m pytest tests/test_nif.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -5
B200RT_NIF_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 2 --samples 64 --skip-cpu-baseline 2> gpurun_out/nifcheck.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), [(k['kernel'][:14], round(k.get('avg_launch_ms') or 0,3), round(k['frac'],3)) for k in d['roofline_kernels'][:4]])"
grep "nif profile" gpurun_out/nifcheck.err | tail -1 | cut -c1-330
timeout 300 python scripts/nif_tolerance.py 2>&1 | tail -4
