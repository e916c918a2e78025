#!/bin/bash
# developer tool: multi-GPU evidence (run under `gpurun --gpus N`): cross-device tests of tpt_render_multi,
# bench.py under torchrun with the in-process arm (e2e_inprocess), outputs kept for profiles/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_multi_n${N}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > gpurun_out/r02_pytest_multi_n${N}.txt 2>&1; tail -6 gpurun_out/r02_pytest_multi_n${N}.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_config4_n${N}.json 2> gpurun_out/r02_bench_config4_n${N}.err; tail -3 gpurun_out/r02_bench_config4_n${N}.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_config4_n${N}.json'))
print('n', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['seconds_per_step'])
print('parity', round(d['parity_mode']['value']), d['parity_mode']['e2e'])
print('inprocess', d.get('e2e_inprocess'))
PY
for extra in "${@:2}"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 2 --warmup 1 --config $extra > gpurun_out/r02_bench_config${extra}_n${N}.json 2> gpurun_out/r02_bench_config${extra}_n${N}.err
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_config${extra}_n${N}.json'));print('config ${extra} n', d['n_gpus'], 'value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['seconds_per_step'], 'parity', d.get('parity_mode',{}).get('value'), 'inprocess', d.get('e2e_inprocess'))"
done
echo SESSION_DONE
