#!/bin/bash
# A/B of the plan-time knobs at small per-GPU batch (eval-only, ADM256): ms per UNet evaluation.
for B in ${BATCHES:-1 2 4 8}; do
  for cfg in "base" "FIDM_PDL=1" "FIDM_FUSE_MIN_PIXELS=4096" "FIDM_HALO_MIN_FILL=40" "FIDM_FUSE_MIN_PIXELS=4096 FIDM_HALO_MIN_FILL=40" "FIDM_FUSE_MIN_PIXELS=4096 FIDM_HALO_MIN_FILL=40 FIDM_PDL=1" "FIDM_FUSE_MIN_PIXELS=1024 FIDM_HALO_MIN_FILL=20"; do
    if [ "$cfg" = "base" ]; then e=""; else e="$cfg"; fi
    ms=$(env $e python bench.py --eval-only --batch $B --workload ${WORKLOAD:-adm256} 2>/dev/null | python -c "import sys,json; print('%.3f' % json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_unet_eval'])")
    echo "B=$B  $ms ms  [$cfg]"
  done
done
