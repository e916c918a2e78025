#!/bin/bash
# developer tool: one bench line per BASELINE.json configuration on one B200 (kept under profiles/)
mkdir -p gpurun_out
for c in 1 2 3a 3b 4; do
  python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02_bench_config${c}_n1.json 2> gpurun_out/r02_bench_config${c}_n1.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_config${c}_n1.json'))
print('config ${c}', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'parity', round(d['parity_mode']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), 'frac', round(d['roofline']['frac'],4), 'prog', (d.get('program_e2e') or {}).get('p3'))
PY
done
python bench.py --config 4 --variant B --steps 3 --warmup 3 > gpurun_out/r02_bench_config4B_n1.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02_bench_config4B_n1.json'));print('config 4B', d['value'], d['e2e']['value'], d['parity_mode']['value'], d['roofline']['frac'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null; head -c 1500 gpurun_out/r02_bench_ref_n1.json
echo SESSION_DONE
