#!/bin/bash
# A/B of kernel variants (libraries under topoflow_glacier_b200/lib/variants/, built by scripts/build_variants.sh):
# bit-pattern checksums on the small workload (must agree between variants that claim identical arithmetic), then the
# bench-shaped launch 16 777 216 cells x 128 steps with aggregates, January and July.
O=gpurun_out; mkdir -p $O
for so in topoflow_glacier_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "== $n"
  TFG_LIBRARY=$so python scripts/prof_run.py --steps 24 --launches 3 --agg 1 --sum 2>&1 | tail -4 | cut -c1-400
  TFG_LIBRARY=$so python scripts/prof_run.py --steps 24 --launches 2 --cells 2097150 --start 6000 --sum 2>&1 | tail -3 | cut -c1-400
  TFG_LIBRARY=$so python scripts/prof_run.py --cells 16777216 --steps 128 --agg 1 --launches 3 2>&1 | tail -2
  TFG_LIBRARY=$so python scripts/prof_run.py --cells 16777216 --steps 128 --agg 1 --launches 2 --start 6600 2>&1 | tail -1
done 2>&1 | tee $O/ab_${1:-x}.log
