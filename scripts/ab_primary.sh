set -e
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for pp in 2 1; do
  B200RT_PRIMARY_PASS=$pp python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline > /dev/null 2>&1
  B200RT_PRIMARY_PASS=$pp timeout 600 ncu --metrics $M --clock-control none -k regex:"path_trace|primary_hit" -c 4 --csv --log-file gpurun_out/ab_pp$pp.csv python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline > /dev/null 2>&1
done
