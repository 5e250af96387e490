#!/bin/bash
# BASELINE.json configs[2..4] (bench.py --preset cfg3|cfg4|cfg5) and a strong-scaling line on N GPUs of one box.
# usage: bash tools/run_presets.sh N TAG   -> gpurun_out/TAG_<preset>_nN.json
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
run() {  # name, bench args...
  local name=$1; shift
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" \
      > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  else
    python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  fi
  echo "== $name rc=$?"; tail -c 700 gpurun_out/${TAG}_${name}_n${N}.json; echo
}
run cfg5 --preset cfg5 --steps 2 --warmup 3 --no-cpu-baseline
run cfg3 --preset cfg3 --steps 1 --warmup 1 --no-cpu-baseline
run cfg4 --preset cfg4
run strong8 --global-batch 8 --steps 3 --warmup 3 --no-cpu-baseline
