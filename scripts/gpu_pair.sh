#!/bin/bash
# experiment: CTA-pair NIF kernel (variants/libb200rt_pair.so, B200RT_NIF_PAIR=2) -- parity tests, then cycles per tile
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/pair_run.log 2>&1
export B200RT_LIB=$PWD/ipu_ray_lib_b200/variants/libb200rt_pair.so
B200RT_NIF_PAIR=2 timeout 300 python -m pytest tests/test_nif.py -m gpu -x -q 2>&1 | tail -15
for mode in 0 2; do
  echo "== pair mode $mode"
  B200RT_NIF_PAIR=$mode B200RT_NIF_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline 2> gpurun_out/pair_$mode.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=[k for k in d['roofline_kernels'] if k['kernel']=='nif_mlp_kernel'][0]
print('nif ms', round(k['avg_launch_ms'],3), 'frac', round(k['frac'],3), 'step ms', round(d['ms_per_step'],2))"
  grep "nif profile" gpurun_out/pair_$mode.err | tail -2 | cut -c1-400
  tail -3 gpurun_out/pair_$mode.err | cut -c1-300
done
