#!/bin/bash
# Round 2, first GPU call: parity over the widened configuration space on the round-1 kernel, a bench line at HEAD,
# and `ncu --set full` of the BENCHED instantiation run_kernel<FastF64,0,1,1,0> at the bench shape (16 777 216 x 128).
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/pytest_r2a.log 2>&1; tail -15 $O/pytest_r2a.log
python bench.py --steps 4 --warmup 3 --no-shared > $O/bench_r2a.json 2> $O/bench_r2a.err; tail -c 600 $O/bench_r2a.json; tail -3 $O/bench_r2a.err
python scripts/prof_run.py --cells 16777216 --steps 128 --agg 1 --launches 2 > $O/plain_r2a_big.log 2>&1; tail -1 $O/plain_r2a_big.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o $O/prof_r2a_bench \
    python scripts/prof_run.py --cells 16777216 --steps 128 --agg 1 --launches 2 > $O/ncu_r2a_big.log 2>&1; tail -2 $O/ncu_r2a_big.log
ls -la $O/prof_r2a_bench.ncu-rep
