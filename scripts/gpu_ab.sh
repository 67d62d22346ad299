#!/bin/bash
# A/B of library builds / env settings on the short bench: lines "label env... lib"
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/ab_run.log 2>&1
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), [(k['kernel'][:9], round(k.get('avg_launch_ms') or 0,4), round(k['frac'],3)) for k in d['roofline_kernels'][:3]])
PY
}
while read -r label lib envs; do
  [ -z "$label" ] && continue
  env $envs B200RT_LIB=$PWD/ipu_ray_lib_b200/$lib timeout 600 python bench.py --steps 2 --warmup 3 --samples 64 --skip-cpu-baseline > gpurun_out/ab.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  summ "$label" gpurun_out/ab.json
done < scripts/ab_cases.txt
