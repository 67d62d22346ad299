#!/bin/bash
# A/B of the chunk overlap on the headline workload at reduced spp: lines "label lib [ENV=val ...]"
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/overlap_ab.log 2>&1
SPP=${SPP:-256}
while read -r label lib envs; do
  [ -z "$label" ] && continue
  echo -n "$label: "; env $envs B200RT_LIB=$PWD/ipu_ray_lib_b200/$lib timeout 300 python scripts/overlap_times.py $SPP 3 2>&1 | tail -2 | tr '\n' ' '; echo
done <<'CASES'
ov_auto libb200rt.so
ov_c16 libb200rt.so WF_SAMPLES_PER_CHUNK=16
ov_c8 libb200rt.so WF_SAMPLES_PER_CHUNK=8
ov_c64 libb200rt.so WF_SAMPLES_PER_CHUNK=64
serial_c16 libb200rt.so WF_SAMPLES_PER_CHUNK=16 WF_CHUNK_OVERLAP=1
serial libb200rt.so WF_CHUNK_OVERLAP=1
CASES
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
