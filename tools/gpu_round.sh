#!/bin/bash
# developer tool: one GPU session for the v7 evidence (ordered by importance; everything lands in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
bash tools/gpu_variants.sh > gpurun_out/variants.txt 2>&1; cat gpurun_out/variants.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.txt 2>&1; tail -14 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 600 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/prof_wave python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
TPT_SMALL_OPEN_BLOCKS=0 timeout 300 python -m pytest tests/test_gpu_hits.py tests/test_gpu_radiance.py -m gpu -q -x -k "cornell" > gpurun_out/pytest_gpu_rects.txt 2>&1; tail -3 gpurun_out/pytest_gpu_rects.txt
python bench.py --steps 3 --warmup 3 --variant B --no-cpu-baseline > gpurun_out/bench_ours_B.json 2>/dev/null; tail -c 200 gpurun_out/bench_ours_B.json | head -c 0; python -c "import json;d=json.load(open('gpurun_out/bench_ours_B.json'));print('B', d['value'], d['roofline']['frac'])"
python bench.py --steps 3 --warmup 3 --kernel mega --no-cpu-baseline > gpurun_out/bench_ours_mega.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --mode parity --spp 256 --no-cpu-baseline > gpurun_out/bench_ours_parity256.json 2>/dev/null
python tools/gpu_bvh_perf.py 32 > gpurun_out/bvh_perf.txt 2>&1; cat gpurun_out/bvh_perf.txt
timeout 240 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
echo SESSION_DONE
