#!/bin/bash
# round-2 GPU check 1: parity tests, then the headline bench (short), config 1, scheduler statistics
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_bench_c2_short.json 2> gpurun_out/r02_bench_c2_short.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; echo "bench c1 rc=$?"
B200RT_WF_PHASE_STATS=1 timeout 300 python scripts/wf_phase_stats.py > gpurun_out/r02_phase_stats.log 2>&1; echo "phase rc=$?"
tail -c 1500 gpurun_out/r02_bench_c2_short.json; tail -c 600 gpurun_out/r02_bench_c1.json; tail -8 gpurun_out/r02_phase_stats.log
