#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench lines for every mode, launch list, dram traffic and one full ncu
# capture of the hot kernel.  Everything lands in gpurun_out/ ; scripts/make_profiles.py condenses it here.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_$TAG.err
python bench.py > $O/bench_${TAG}_f64fast.json 2>> $O/bench_$TAG.err
python bench.py --mode f64 --no-cpu > $O/bench_${TAG}_f64.json 2>> $O/bench_$TAG.err
python bench.py --mode f32 --no-cpu > $O/bench_${TAG}_f32.json 2>> $O/bench_$TAG.err
# launch list of the bench command (shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu > $O/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launch.log 2>&1
# DRAM traffic of one launch of the hot kernel at the bench configuration
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:run_kernel -s 3 -c 1 --csv --log-file $O/traffic_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_traffic.log 2>&1
# full capture on the small fixed workload (scripts/prof_run.py)
for m in f64_fast f64 f32; do
  python scripts/prof_run.py --mode $m --steps 24 --launches 2 > $O/plain_$m.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_$m \
      python scripts/prof_run.py --mode $m --steps 24 --launches 2 > $O/ncu_$m.log 2>&1
  tail -1 $O/plain_$m.log
done
tail -c 400 $O/bench_${TAG}_f64fast.json; echo; tail -3 $O/bench_$TAG.err
