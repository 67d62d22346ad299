#!/bin/bash
# one GPU visit: parity tests, A/B of the variants in scripts/ab_cases.txt, short benches
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${TAG:-r02i}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
bash scripts/gpu_ab2.sh
cp gpurun_out/ab2_run.log gpurun_out/${TAG}_ab.log
timeout 600 python bench.py --steps 2 --warmup 3 --samples 64 --skip-cpu-baseline > gpurun_out/${TAG}_c2_64spp.json 2> gpurun_out/${TAG}_c2_64spp.err
python - gpurun_out/${TAG}_c2_64spp.json <<'PY' >> gpurun_out/${TAG}_ab.log
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('c2 64spp value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [(k['kernel'], round(k.get('avg_launch_ms') or 0,4), k.get('launches_per_step'), round(k['frac'],3)) for k in d['roofline_kernels'][:3]])
PY
