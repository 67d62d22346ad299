#!/bin/bash
# NIF kernel time against the number of CTAs (is it bound chip-wide by L2 or per SM?)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/nifgrid.log 2>&1
for g in 148 140 132 124 111 96 74; do
  echo -n "grid $g: "
  B200RT_NIF_GRID=$g B200RT_NIF_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline 2> gpurun_out/nifgrid_$g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=[k for k in d['roofline_kernels'] if k['kernel']=='nif_mlp_kernel'][0]
print('nif ms', round(k['avg_launch_ms'],3), 'frac', round(k['frac'],3), 'step ms', round(d['ms_per_step'],2), 'clock', d['clocks']['sm_mhz'])"
  grep "nif profile" gpurun_out/nifgrid_$g.err | tail -1 | cut -c1-400
done
