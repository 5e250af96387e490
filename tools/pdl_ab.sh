#!/bin/bash
# A/B of programmatic dependent launch with the single-wave early trigger, eval-only: ms per UNet evaluation.
for W in ${WORKLOADS:-adm256 ref_ffhq256}; do
  for B in ${BATCHES:-1 2 4 8}; do
    for cfg in "base" "FIDM_PDL=1"; do
      if [ "$cfg" = "base" ]; then e=""; else e="$cfg"; fi
      ms=$(env $e python bench.py --eval-only --batch $B --workload $W 2>/dev/null | python -c "import sys,json; print('%.3f' % json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_unet_eval'])")
      echo "$W B=$B  $ms ms  [$cfg]"
    done
  done
done
