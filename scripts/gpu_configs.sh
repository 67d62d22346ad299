#!/bin/bash
# The named configurations of BASELINE.json on N GPUs (N = 1: every configuration; N > 1: the headline and the two
# configurations north_star names for 1/2/4/8). Lines go to gpurun_out/<tag>_c<k>_n<N>.json
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${N:-1}; TAG=${TAG:-r02}
exec > gpurun_out/${TAG}_configs_n${N}.log 2>&1
run() {  # config, extra args
  local c=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $c "$@" > gpurun_out/${TAG}_c${c}_n1.json 2> gpurun_out/${TAG}_c${c}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + c)) \
      bench.py --gpus $N --config $c "$@" > gpurun_out/${TAG}_c${c}_n${N}.json 2> gpurun_out/${TAG}_c${c}_n${N}.err
  fi
  echo "config $c N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_c${c}_n${N}.json | head -c 300; echo
}
if [ "$N" = 1 ]; then
  run 1
  run 3 --samples 256 --steps 2 --warmup 3
  run 4
  run 5 --samples 16 --steps 1 --warmup 3
  python scripts/pcie_floor.py > gpurun_out/${TAG}_pcie_floor.log 2>&1; cat gpurun_out/${TAG}_pcie_floor.log
else
  run 2
  run 4
  run 5 --samples 16 --steps 1 --warmup 3 --skip-cpu-baseline
fi
