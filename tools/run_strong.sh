#!/bin/bash
# strong scaling: 8 images split over N GPUs (bench.py --global-batch 8); usage: bash tools/run_strong.sh N TAG
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N \
  --global-batch 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_strong8_n${N}.json 2> gpurun_out/${TAG}_strong8_n${N}.err
echo "rc=$?"; cut -c1-260 gpurun_out/${TAG}_strong8_n${N}.json
