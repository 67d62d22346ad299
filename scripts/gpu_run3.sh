#!/bin/bash
# round-2 GPU check 3: all GPU tests (TMA shadow kernel, pipelined streaming, CLI), config 1 with/without TMA staging,
# the CUB-sorted ray order experiment
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu3.log
tail -15 gpurun_out/r02_pytest_gpu3.log
timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1_tma.json 2> gpurun_out/r02_bench_c1_tma.err; echo "bench c1 tma rc=$?"
B200RT_SHADOW_TMA=0 timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1_notma.json 2> gpurun_out/r02_bench_c1_notma.err; echo "bench c1 no-tma rc=$?"
for f in tma notma; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c1_$f.json').read().strip().splitlines()[-1])
print('$f', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'kernel avg ms', d['roofline_kernels'][0]['avg_launch_ms'], 'frac', d['roofline_kernels'][0]['frac'])
PY
done
# sort experiment: 64 spp, no cpu baseline
for srt in 0 1; do
B200RT_LIB=$PWD/ipu_ray_lib_b200/libb200rt_exp.so B200RT_SORT=$srt timeout 600 python bench.py --steps 1 --warmup 1 --samples 64 --skip-cpu-baseline > gpurun_out/r02_sort$srt.json 2> gpurun_out/r02_sort$srt.err; echo "sort$srt rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_sort$srt.json').read().strip().splitlines()[-1])
print('sort=$srt', 'ms/step', round(d['ms_per_step'],2), [(k['kernel'], round(k.get('avg_launch_ms') or 0,3), k.get('launches_per_step')) for k in d['roofline_kernels'][:3]])
PY
done
B200RT_LIB=$PWD/ipu_ray_lib_b200/libb200rt_exp.so B200RT_SORT=1 B200RT_WF_PHASE_STATS=1 timeout 300 python scripts/wf_phase_stats.py > gpurun_out/r02_phase_stats_sorted.log 2>&1; tail -8 gpurun_out/r02_phase_stats_sorted.log
