#!/bin/bash
# A/B of library builds / env settings on per-kernel times: lines "label lib [ENV=val ...]" in scripts/ab_cases.txt
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/ab2_run.log 2>&1
for round in 1 2; do
while read -r label lib envs; do
  [ -z "$label" ] && continue
  echo -n "$label: "; env $envs B200RT_LIB=$PWD/ipu_ray_lib_b200/$lib timeout 600 python scripts/wf_kernel_times.py 64 5 2>&1 | tail -1
done < scripts/ab_cases.txt
done
