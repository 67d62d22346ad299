#!/bin/bash
# round-2 GPU check 2: parity tests on the pipelined host-streaming path, ncu --set full of wf_trace (bounce 1)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu2.log
tail -5 gpurun_out/r02_pytest_gpu2.log
timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1b.json 2> gpurun_out/r02_bench_c1b.err; echo "bench c1 rc=$?"
python bench.py --steps 1 --warmup 0 --samples 32 --skip-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel -s 1 -c 1 -f -o gpurun_out/r02_wftrace python bench.py --steps 1 --warmup 0 --samples 32 --skip-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
tail -c 400 gpurun_out/r02_bench_c1b.json
