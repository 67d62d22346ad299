#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
exec > gpurun_out/tail_ab.log 2>&1
for rep in 1 2; do
for t in 0 1; do
  for c in 3 4; do
    echo -n "config $c tail_env=$t: "
    if [ $t = 1 ]; then export B200RT_WF_TAIL=1; else unset B200RT_WF_TAIL; fi
    timeout 300 python bench.py --config $c --samples 64 --steps 2 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'], [(k['kernel'][:8], round(k.get('avg_launch_ms') or 0,3), k.get('launches_per_step')) for k in d['roofline_kernels'][:3]])"
  done
done
done
