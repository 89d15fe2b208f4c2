#!/bin/bash
# Multi-GPU lines (run with `gpurun --gpus N -- bash scripts/scale_round.sh N`): the raster workload (weak scaling,
# exact aggregates so that the sums are bit-identical for every N) and the 100 M-cell regional grid (strong scaling).
N=${1:-2}
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$RUN bench.py --gpus $N --steps 6 --warmup 3 --agg exact > $O/bench_${N}gpu_raster.json 2> $O/bench_${N}gpu.err
$RUN bench.py --gpus $N --steps 6 --warmup 3 --workload regional --agg exact > $O/bench_${N}gpu_regional.json 2>> $O/bench_${N}gpu.err
python - <<PY
import json
for w in ("raster", "regional"):
    try:
        d = json.loads(open("$O/bench_${N}gpu_%s.json" % w).read().strip().splitlines()[-1])
        print(w, "N=$N value %.4g e2e %.4g kernel_ms %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
    except Exception as e:
        print(w, "failed", e)
PY
tail -n 3 $O/bench_${N}gpu.err
