#!/bin/bash
# round 2, wf_trace on the streaming-step core: parity, short benches, scheduler statistics
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), [(k['kernel'], round(k.get('avg_launch_ms') or 0,4), k.get('launches_per_step'), round(k['frac'],3)) for k in d['roofline_kernels'][:3]])
PY
}
timeout 600 python bench.py --steps 1 --warmup 3 --samples 64 --skip-cpu-baseline > gpurun_out/${TAG}_c2_64spp.json 2> gpurun_out/${TAG}_c2_64spp.err; summ c2 gpurun_out/${TAG}_c2_64spp.json
B200RT_WF_PHASE_STATS=1 timeout 300 python scripts/wf_phase_stats.py > gpurun_out/${TAG}_phase_stats.log 2>&1; tail -8 gpurun_out/${TAG}_phase_stats.log
for thr in 4 12 16 20; do
B200RT_WF_THRESHOLD=$thr timeout 600 python bench.py --steps 1 --warmup 3 --samples 64 --skip-cpu-baseline > gpurun_out/thr.json 2>/dev/null; summ thr$thr gpurun_out/thr.json
done
timeout 600 python bench.py --config 3 --samples 64 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/${TAG}_c3_short.json 2> gpurun_out/${TAG}_c3_short.err; summ c3 gpurun_out/${TAG}_c3_short.json
timeout 600 python bench.py --config 4 --samples 16 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/${TAG}_c4_short.json 2> gpurun_out/${TAG}_c4_short.err; summ c4 gpurun_out/${TAG}_c4_short.json
tail -3 gpurun_out/${TAG}*.err | tail -30
