#!/bin/bash
# developer tool: bench.py under torchrun at 8 / 4 / 2 ranks, BASELINE configs[4] (4096^2 x 4096 spp) at 8, tpt_render_multi timing
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench$n.err > gpurun_out/bench_n$n.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1]); print($n, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'wall s/step', round(d['wall_seconds_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['seconds_per_step'],4), 'cull', round(d['bundle_cull']['value']), d['clocks'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --size 4096 --spp 4096 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/bench_c5.err > gpurun_out/bench_config5_n8.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_config5_n8.json').read().strip().splitlines()[-1]); print('config5 n8 value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['seconds_per_step'],4), 'cull', round(d['bundle_cull']['value']))"
python tools/gpu_multi_perf.py 2>&1 | tail -3
