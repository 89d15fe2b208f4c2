#!/bin/bash
# Time the fixed profiling workload with every experimental library under topoflow_glacier_b200/lib/variants/.
for so in topoflow_glacier_b200/lib/variants/*.so; do
  echo "== $(basename $so)"; TFG_LIBRARY=$so python scripts/prof_run.py --mode ${1:-f64_fast} --steps 24 --launches 3 2>&1 | tail -1
done
