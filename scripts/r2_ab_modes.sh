#!/bin/bash
# A/B of kernel variants in the strict float64 and the float32 mode (same launch shape as scripts/r2_ab.sh)
O=gpurun_out; mkdir -p $O
for so in topoflow_glacier_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "== $n"
  for m in f64 f32; do
    TFG_LIBRARY=$so python scripts/prof_run.py --mode $m --steps 24 --launches 2 --agg 1 --sum 2>&1 | grep -E "checksum (ring|agg)" | cut -c1-100
    TFG_LIBRARY=$so python scripts/prof_run.py --mode $m --cells 16777216 --steps 128 --agg 1 --launches 3 2>&1 | tail -2
  done
done 2>&1 | tee $O/abm_${1:-x}.log
