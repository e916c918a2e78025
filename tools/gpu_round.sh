#!/bin/bash
# developer tool: one GPU session = tests + smoke + bench (both arms) + ncu evidence
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 400 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json
python bench.py --steps 3 --warmup 3 --variant B --no-cpu-baseline > gpurun_out/bench_ours_B.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --kernel mega --no-cpu-baseline > gpurun_out/bench_ours_mega.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --mode parity --spp 256 --no-cpu-baseline > gpurun_out/bench_ours_parity256.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/prof_wave python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
# BVH / texture scenes: throughput + one full capture each (L1/L2 hit rates, DRAM traffic)
python tools/gpu_bvh_perf.py 32 > gpurun_out/bvh_perf.txt 2>&1; cat gpurun_out/bvh_perf.txt
KERN=1 bash tools/gpu_prof_scenes.sh > gpurun_out/scenes_perf.txt 2>&1; cat gpurun_out/scenes_perf.txt
