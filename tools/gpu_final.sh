#!/bin/bash
# developer tool: short closing session on ONE B200 after a last source change. The ncu captures come FIRST and are
# summarised on the box into profiles/ncu_summary.json, so that the bench lines taken afterwards carry the
# executed-instruction view of the very build they measure (bench.py gates it on the source fingerprint).
mkdir -p gpurun_out
summarise() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$2_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/r02_ncu_full_$2_details.txt 2>/dev/null
  python profiles/tools/ncu_lines.py gpurun_out/$1.ncu-rep 60 > gpurun_out/r02_ncu_full_$2_lines.txt 2>&1
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_fast_A python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_fast_A.log 2>&1
summarise r02_prof_fast_A render_wave_fast_A
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_parity_A python bench.py --steps 1 --warmup 1 --mode parity --spp 64 --no-cpu-baseline > gpurun_out/r02_ncu_parity_A.log 2>&1
summarise r02_prof_parity_A render_wave_parity_A
cp gpurun_out/r02_ncu_full_render_wave_fast_A_raw.csv gpurun_out/r02_ncu_full_render_wave_parity_A_raw.csv profiles/
python profiles/tools/ncu_summary.py profiles/r02_ncu_full_render_wave_fast_A_raw.csv 368640000 "ncu --set full --clock-control none, bench.py --spp 256 --steps 1 --warmup 1 (launch 2 of render_wave_kernel); the framebuffer traffic of a launch does not depend on spp: R x npix x 12 B" 4A fast > profiles/ncu_summary.json
cp profiles/ncu_summary.json gpurun_out/ncu_summary_on_box.json
rm -f gpurun_out/r02_prof_parity_A.ncu-rep gpurun_out/r02_prof_fast_A.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -1 gpurun_out/r02_smoke.txt
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_config4_n1.json 2> gpurun_out/r02_bench_config4_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_config4_n1.json'))
print('config 4 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'parity', round(d['parity_mode']['value'],1), 'cpu', d['cpu_baseline']['value'], 'issue', json.dumps(d['roofline'].get('issue'))[:300], 'traffic', d['roofline'].get('traffic'))"
python bench.py --config 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_config2_n1.json 2> gpurun_out/r02_bench_config2_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_config2_n1.json'))
print('config 2 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'parity', round(d['parity_mode']['value'],1))"
echo SESSION_DONE
