# A/B of the wavefront scheduler against the megakernel on the bench workload (64 spp), then a per-kernel ncu pass.
set -e
run() { python bench.py --steps 2 --warmup 1 --samples 64 --skip-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], [(k['kernel'],k.get('avg_launch_ms')) for k in d['roofline_kernels'][:2]])"; }
echo "megakernel"; run --traversal 2
for thr in 4 8 12; do echo "wavefront thr=$thr"; run --traversal 4; done
