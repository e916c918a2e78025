#!/bin/bash
# developer tool: round-2 session 6 -- full GPU test suite + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r02_pytest_gpu_s6.txt 2>&1; tail -12 gpurun_out/r02_pytest_gpu_s6.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_s6.txt 2>&1; tail -2 gpurun_out/r02_smoke_s6.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_s6.json 2> gpurun_out/r02_bench_default_s6.err; tail -5 gpurun_out/r02_bench_default_s6.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_s6.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_step'])
print('parity', d['parity_mode']['value'], d['parity_mode']['e2e'], d['parity_mode']['roofline']['frac'])
print('cull', d['bundle_cull']['value'])
print('cpu', d.get('cpu_baseline'))
print('program', d.get('program_e2e'))
print('roofline', d['roofline']['frac'], d['roofline']['issue'])
PY
echo SESSION_DONE
