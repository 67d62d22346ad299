#!/bin/bash
# round-2 evidence run: full default bench (+ reference arm), ncu launch list with DRAM bytes, ncu --set full of the
# three dominant kernels (bounce-1 wf_trace / wf_shade, the NIF kernel). Numbers printed under ncu are never bench values.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${TAG:-r02}
exec > gpurun_out/${TAG}_prof_run.log 2>&1
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
# ncu serialises kernels anyway; chunk overlap off so that the launches are the ones bench.py's roofline step times
export B200RT_OVERLAP=0
timeout 900 ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 0 --samples 128 --skip-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
for k in wf_trace_kernel wf_shade_kernel nif_mlp; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_full_$k python bench.py --steps 1 --warmup 0 --samples 32 --skip-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
