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
ov_mb2 variants/libb200rt_mb2.so
ov_mb3 variants/libb200rt_mb3.so
serial libb200rt.so WF_CHUNK_OVERLAP=1
serial_mb2 variants/libb200rt_mb2.so WF_CHUNK_OVERLAP=1
serial_mb3 variants/libb200rt_mb3.so WF_CHUNK_OVERLAP=1
CASES
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
