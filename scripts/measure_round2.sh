#!/bin/bash
# Round 2, on the GPU box (via gpurun): tests; `ncu --set full` of the BENCHED instantiations at the bench shape
# (16 777 216 cells x 128 steps, aggregates + integrals on, October start like bench.py) for the three arithmetic modes;
# the launch list of the bench command; then the bench lines.  bench.py reads profiles/kernel_mix.json and
# profiles/traffic.json, so scripts/make_profiles.py refreshes them on the box before the bench lines are taken.
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
for m in f64_fast f64 f32; do
  python scripts/prof_run.py --mode $m --cells 16777216 --steps 128 --agg 1 --start 0 --launches 2 > $O/plain_$m.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_$m \
      python scripts/prof_run.py --mode $m --cells 16777216 --steps 128 --agg 1 --start 0 --launches 2 > $O/ncu_$m.log 2>&1
  tail -n 1 $O/plain_$m.log
done
# launch list of the bench command (shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu --no-modes --no-strong > $O/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-modes --no-strong > $O/ncu_launch.log 2>&1
python scripts/make_profiles.py $TAG > /dev/null 2>&1   # refresh kernel_mix.json / traffic.json for the bench lines below
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_$TAG.err
python bench.py > $O/bench_${TAG}_f64fast.json 2>> $O/bench_$TAG.err
TFG_TMA_STAGING=1 python bench.py --no-cpu --no-modes --no-strong --no-shared > $O/bench_${TAG}_f64fast_tma.json 2>> $O/bench_$TAG.err
python bench.py --mode f32 --no-cpu --no-modes --no-strong --no-shared > $O/bench_${TAG}_f32.json 2>> $O/bench_$TAG.err
TFG_TMA_STAGING=1 python bench.py --mode f32 --no-cpu --no-modes --no-strong --no-shared > $O/bench_${TAG}_f32_tma.json 2>> $O/bench_$TAG.err
python bench.py --agg exact --no-cpu --no-modes --no-strong > $O/bench_${TAG}_f64fast_exactagg.json 2>> $O/bench_$TAG.err
python bench.py --e2e-raw float32 --no-cpu --no-modes --no-strong --no-shared > $O/bench_${TAG}_f64fast_e2e_float32.json 2>> $O/bench_$TAG.err
python scripts/water_year.py > $O/water_year_f64fast.json 2> $O/water_year.err
python scripts/bmi_latency.py > $O/bmi_latency.log 2>&1
for f in f64fast f64fast_tma f32 f32_tma; do python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}_$f.json").read().strip().splitlines()[-1])
print("$f", "%.2f G" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "e2e %.2f G" % (d["e2e"]["value"]/1e9), d["clocks"])
PY
done
tail -n 3 $O/bench_$TAG.err
