#!/bin/bash
# 100 M-cell regional grid (strong scaling) on N GPUs: gpurun --gpus N -- bash scripts/scale_regional.sh N
N=${1:-2}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 6 --warmup 3 --workload regional --agg exact > $O/bench_${N}gpu_regional.json 2> $O/bench_${N}gpu.err
python -c "
import json
d = json.loads(open('$O/bench_${N}gpu_regional.json').read().strip().splitlines()[-1])
print('regional N=$N value %.4g e2e %.4g kernel_ms %.2f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))"
