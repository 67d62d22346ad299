#!/bin/bash
# one GPU visit: parity tests, A/B of the variants in scripts/ab_cases.txt, short benches
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${TAG:-r02i}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
bash scripts/gpu_ab2.sh
cp gpurun_out/ab2_run.log gpurun_out/${TAG}_ab.log
