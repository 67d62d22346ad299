#!/bin/bash
# A/B of the chunk overlap on the headline workload at reduced spp: lines "label lib [ENV=val ...]"
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/overlap_ab.log 2>&1
SPP=${SPP:-256}
while read -r label lib envs; do
  [ -z "$label" ] && continue
  echo -n "$label: "; env $envs B200RT_LIB=$PWD/ipu_ray_lib_b200/$lib timeout 300 python scripts/overlap_times.py $SPP 3 2>&1 | tail -2 | tr '\n' ' '; echo
done <<'CASES'
base libb200rt.so
ov_l2 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 WF_CHECK=1
ov_l2_g136 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 B200RT_NIF_GRID=136
ov_l2_g128 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 B200RT_NIF_GRID=128
ov_l2_g120 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 B200RT_NIF_GRID=120
ov_l2_g112 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 B200RT_NIF_GRID=112
ov_l2_g100 libb200rt.so WF_CHUNK_OVERLAP=2 WF_SCENE_RESIDENCY=2 B200RT_NIF_GRID=100
ov_sh_g120 libb200rt.so WF_CHUNK_OVERLAP=2 B200RT_NIF_GRID=120
ov_sh_g100 libb200rt.so WF_CHUNK_OVERLAP=2 B200RT_NIF_GRID=100
base2 libb200rt.so
CASES
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
