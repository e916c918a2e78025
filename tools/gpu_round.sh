#!/bin/bash
# developer tool: one evidence session on ONE B200 (round 2; ordered by importance; everything lands in gpurun_out/,
# the summaries are copied to profiles/ afterwards -- see profiles/README.md). About 6 GPU-minutes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x -rs --durations=6 > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -14 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -2 gpurun_out/r02_smoke.txt
for c in 4 1 2 3a 3b; do
  python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02_bench_config${c}_n1.json 2> gpurun_out/r02_bench_config${c}_n1.err
  python -c "
import json
d=json.load(open('gpurun_out/r02_bench_config${c}_n1.json'))
print('config ${c}', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'parity', round(d['parity_mode']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), 'frac', round(d['roofline']['frac'],4), 'prog', (d.get('program_e2e') or {}).get('p3'))"
done
python bench.py --config 4 --variant B --steps 3 --warmup 3 > gpurun_out/r02_bench_config4B_n1.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02_bench_config4B_n1.json'));print('config 4B', d['value'], d['e2e']['value'], d['parity_mode']['value'], d['roofline']['frac'])"
python bench.py --kernel mega --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_config4_mega_n1.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null; head -c 400 gpurun_out/r02_bench_ref_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# ncu --set full captures, summarised ON THE BOX (gpurun copies at most 64 MiB back; a report is ~15 MB): raw counters,
# details page, per-source-line table; only the headline report itself is kept
summarise() {  # $1 = report stem under gpurun_out, $2 = name under which the summaries are stored
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$2_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/r02_ncu_full_$2_details.txt 2>/dev/null
  python profiles/tools/ncu_lines.py gpurun_out/$1.ncu-rep 60 > gpurun_out/r02_ncu_full_$2_lines.txt 2>&1
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_fast_A python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_fast_A.log 2>&1
summarise r02_prof_fast_A render_wave_fast_A
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_parity_A python bench.py --steps 1 --warmup 1 --mode parity --spp 64 --no-cpu-baseline > gpurun_out/r02_ncu_parity_A.log 2>&1
summarise r02_prof_parity_A render_wave_parity_A
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_fast_B python bench.py --steps 1 --warmup 1 --variant B --spp 64 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_fast_B.log 2>&1
summarise r02_prof_fast_B render_wave_fast_B
QUICK=1 KERN=1 bash tools/gpu_prof_scenes.sh > gpurun_out/r02_scene_prof.log 2>&1; tail -3 gpurun_out/r02_scene_prof.log
args=""
for sc in random_scene two_perlin_spheres earth; do [ -f gpurun_out/scene_${sc}_k1.ncu-rep ] && args="$args $sc=gpurun_out/scene_${sc}_k1.ncu-rep"; done
python profiles/tools/ncu_scene_summary.py cornell_A_fast=gpurun_out/r02_prof_fast_A.ncu-rep cornell_A_parity=gpurun_out/r02_prof_parity_A.ncu-rep cornell_B_fast=gpurun_out/r02_prof_fast_B.ncu-rep $args > gpurun_out/r02_scene_ncu_summary.json
python profiles/tools/ncu_lines.py gpurun_out/scene_random_scene_k1.ncu-rep 40 > gpurun_out/r02_ncu_full_random_scene_lines.txt 2>&1
rm -f gpurun_out/scene_*_k1.ncu-rep gpurun_out/r02_prof_parity_A.ncu-rep gpurun_out/r02_prof_fast_B.ncu-rep gpurun_out/*_k[01].ncu-rep gpurun_out/prof_wave*.ncu-rep
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
timeout 300 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.log
echo SESSION_DONE
