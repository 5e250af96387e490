#!/bin/bash
# K1h / K1s fill threshold (FIDM_HALO_MIN_FILL) at small per-GPU batch, eval-only, two rounds per setting.
for W in ${WORKLOADS:-adm256}; do
  for B in ${BATCHES:-1 2 4}; do
    for rep in 1 2; do
      for cfg in "base" "FIDM_HALO_MIN_FILL=40" "FIDM_HALO_MIN_FILL=55" "FIDM_HALO_MIN_FILL=90"; do
        if [ "$cfg" = "base" ]; then e=""; else e="$cfg"; fi
        ms=$(env $e python bench.py --eval-only --batch $B --workload $W 2>/dev/null | python -c "import sys,json; print('%.3f' % json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_unet_eval'])")
        echo "$W B=$B  $ms ms  [$cfg]"
      done
    done
  done
done
