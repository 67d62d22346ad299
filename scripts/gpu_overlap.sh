#!/bin/bash
# A/B of the chunk overlap on the headline workload at reduced spp: lines "label lib [ENV=val ...]"
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/overlap_ab.log 2>&1
SPP=${SPP:-256}
while read -r label lib envs; do
  [ -z "$label" ] && continue
  echo -n "$label: "; env $envs B200RT_LIB=$PWD/ipu_ray_lib_b200/$lib timeout 300 python scripts/overlap_times.py $SPP 3 2>&1 | tail -2 | tr '\n' ' '; echo
done <<'CASES'
ov libb200rt.so
ov_tsh variants/libb200rt_tsh.so
l2_serial libb200rt.so WF_CHUNK_OVERLAP=1 WF_SCENE_RESIDENCY=2
l2_serial_tsh variants/libb200rt_tsh.so WF_CHUNK_OVERLAP=1 WF_SCENE_RESIDENCY=2
serial libb200rt.so WF_CHUNK_OVERLAP=1
serial_tsh variants/libb200rt_tsh.so WF_CHUNK_OVERLAP=1
CASES
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
