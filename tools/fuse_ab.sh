#!/bin/bash
# Interleaved A/B of the fused-statistics threshold (FIDM_FUSE_MIN_PIXELS), eval-only, three rounds per setting.
for W in ${WORKLOADS:-ref_ffhq256 adm256}; do
  for B in ${BATCHES:-8 1}; do
    for rep in 1 2 3; do
      for cfg in "base" "FIDM_FUSE_MIN_PIXELS=4096" "FIDM_FUSE_MIN_PIXELS=1024"; do
        if [ "$cfg" = "base" ]; then e=""; else e="$cfg"; fi
        ms=$(env $e python bench.py --eval-only --batch $B --workload $W 2>/dev/null | python -c "import sys,json; print('%.3f' % json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_unet_eval'])")
        echo "$W B=$B  $ms ms  [$cfg]"
      done
    done
  done
done
