#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu4.log
tail -4 gpurun_out/r02_pytest_gpu4.log
python scripts/nif_tolerance.py > gpurun_out/r02_nif_tolerance.log 2>&1; cat gpurun_out/r02_nif_tolerance.log
python scripts/pcie_floor.py > gpurun_out/r02_pcie_floor.log 2>&1; cat gpurun_out/r02_pcie_floor.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), [(k['kernel'], round(k.get('avg_launch_ms') or 0,4), k.get('launches_per_step'), round(k['frac'],3)) for k in d['roofline_kernels'][:3]])
PY
}
timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1_tma.json 2> gpurun_out/r02_bench_c1_tma.err; summ tma gpurun_out/r02_bench_c1_tma.json
B200RT_SHADOW_TMA=0 timeout 300 python bench.py --config 1 --skip-cpu-baseline > gpurun_out/r02_bench_c1_notma.json 2> gpurun_out/r02_bench_c1_notma.err; summ notma gpurun_out/r02_bench_c1_notma.json
for t in 32768 65536 262144 524288; do B200RT_TILE_RAYS=$t timeout 300 python bench.py --config 1 --skip-cpu-baseline --steps 8 > gpurun_out/tile.json 2>/dev/null; summ tile$t gpurun_out/tile.json; done
for srt in 0 1; do
B200RT_LIB=$PWD/ipu_ray_lib_b200/libb200rt_exp.so B200RT_SORT=$srt timeout 600 python bench.py --steps 1 --warmup 1 --samples 64 --skip-cpu-baseline > gpurun_out/r02_sort$srt.json 2> gpurun_out/r02_sort$srt.err; echo "sort$srt rc=$?"
summ sort$srt gpurun_out/r02_sort$srt.json
done
B200RT_LIB=$PWD/ipu_ray_lib_b200/libb200rt_exp.so B200RT_SORT=1 B200RT_WF_PHASE_STATS=1 timeout 300 python scripts/wf_phase_stats.py > gpurun_out/r02_phase_stats_sorted.log 2>&1; tail -8 gpurun_out/r02_phase_stats_sorted.log
timeout 600 python bench.py --config 3 --samples 64 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_c3_short.json 2> gpurun_out/r02_c3_short.err; summ c3 gpurun_out/r02_c3_short.json
timeout 600 python bench.py --config 4 --samples 16 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_c4_short.json 2> gpurun_out/r02_c4_short.err; summ c4 gpurun_out/r02_c4_short.json
timeout 900 python bench.py --config 5 --width 4096 --height 4096 --samples 8 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_c5_short.json 2> gpurun_out/r02_c5_short.err; summ c5 gpurun_out/r02_c5_short.json
tail -3 gpurun_out/*.err | tail -30
