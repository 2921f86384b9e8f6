#!/bin/bash
# Runs the quick (device-resident) bench for every library variant under perturb_b200/variants.
# usage: scripts/sweep_variants.sh "c2 c3s c4s" [steps]
WL=${1:-c2}
STEPS=${2:-10}
for lib in perturb_b200/variants/*.so; do
  for w in $WL; do
    PERTURB_GPU_LIB=$PWD/$lib python bench.py --quick --workload $w --steps $STEPS --warmup 3 2>&1 | tail -1
  done
done
