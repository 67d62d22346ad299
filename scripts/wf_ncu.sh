# per-kernel metrics of one wavefront chunk (32 spp)
set -e
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
python bench.py --steps 1 --warmup 1 --samples 32 --skip-cpu-baseline --traversal 4 > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:"wf_" -c 24 --csv --log-file gpurun_out/ncu_wf2.csv python bench.py --steps 1 --warmup 0 --samples 32 --skip-cpu-baseline --traversal 4 > /dev/null 2>&1
