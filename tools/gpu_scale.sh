#!/bin/bash
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench$n.err > gpurun_out/bench_n$n.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1]); print($n, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'wall s/step', round(d['wall_seconds_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['seconds_per_step'],4), d['clocks'])"
done
python tools/gpu_multi_perf.py 2>&1 | tail -3
