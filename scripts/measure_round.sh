#!/bin/bash
# Run on the GPU box (via gpurun): tests, one full ncu capture per arithmetic mode on the small fixed workload, DRAM
# traffic and launch list of the bench command, THEN the bench lines (bench.py reads profiles/kernel_mix.json and
# profiles/traffic.json, so they are refreshed on the box first).  Everything lands in gpurun_out/;
# scripts/make_profiles.py condenses it into profiles/ (run it again locally after the call).
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
# full capture on the small fixed workload (scripts/prof_run.py)
for m in f64_fast f64 f32; do
  python scripts/prof_run.py --mode $m --steps 24 --launches 2 > $O/plain_$m.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_$m \
      python scripts/prof_run.py --mode $m --steps 24 --launches 2 > $O/ncu_$m.log 2>&1
  tail -n 1 $O/plain_$m.log
done
# DRAM traffic of one launch of the hot kernel at the bench configuration
python bench.py --steps 1 --warmup 3 --no-cpu > $O/plain_traffic.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:run_kernel -s 3 -c 1 --csv --log-file $O/traffic_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_traffic.log 2>&1
# launch list of the bench command (shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu > $O/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launch.log 2>&1
python scripts/make_profiles.py $TAG > /dev/null 2>&1   # refresh kernel_mix.json / traffic.json for the bench lines below
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_$TAG.err
python bench.py > $O/bench_${TAG}_f64fast.json 2>> $O/bench_$TAG.err
python bench.py --mode f64 --no-cpu > $O/bench_${TAG}_f64.json 2>> $O/bench_$TAG.err
python bench.py --mode f32 --no-cpu > $O/bench_${TAG}_f32.json 2>> $O/bench_$TAG.err
python bench.py --agg exact --no-cpu > $O/bench_${TAG}_f64fast_exactagg.json 2>> $O/bench_$TAG.err
python bench.py --workload regional > $O/bench_${TAG}_regional_1gpu.json 2>> $O/bench_$TAG.err
tail -c 400 $O/bench_${TAG}_f64fast.json; echo; tail -n 3 $O/bench_$TAG.err
