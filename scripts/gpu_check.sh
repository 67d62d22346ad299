#!/bin/bash
# parity + short benches (+ optional ncu) after a trace-path change
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${TAG:-r02d}
exec > gpurun_out/${TAG}_run.log 2>&1
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), [(k['kernel'], round(k.get('avg_launch_ms') or 0,4), k.get('launches_per_step'), round(k['frac'],3)) for k in d['roofline_kernels'][:3]])
PY
}
timeout 600 python bench.py --steps 2 --warmup 3 --samples 64 --skip-cpu-baseline > gpurun_out/${TAG}_c2_64spp.json 2> gpurun_out/${TAG}_c2_64spp.err; summ c2 gpurun_out/${TAG}_c2_64spp.json
timeout 600 python bench.py --config 3 --samples 64 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/${TAG}_c3_short.json 2> gpurun_out/${TAG}_c3_short.err; summ c3 gpurun_out/${TAG}_c3_short.json
timeout 600 python bench.py --config 4 --samples 16 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/${TAG}_c4_short.json 2> gpurun_out/${TAG}_c4_short.err; summ c4 gpurun_out/${TAG}_c4_short.json
if [ -n "$NCU" ]; then
python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU -s 1 -c 1 -f -o gpurun_out/${TAG}_$NCU python bench.py --steps 1 --warmup 0 --samples 32 --skip-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log | cut -c1-300
fi
