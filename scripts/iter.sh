#!/bin/bash
# Quick GPU iteration (via gpurun): parity tests, then the fixed profiling workload for the given modes.
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/iter_pytest.log 2>&1; tail -3 $O/iter_pytest.log
for m in ${@:-f64_fast}; do python scripts/prof_run.py --mode $m --steps 24 --launches 3 2>&1 | tail -2; done
