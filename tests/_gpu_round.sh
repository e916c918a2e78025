#!/bin/bash
# scratch: one GPU session = tests + bench + ncu evidence
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_mega -s 1 -c 1 -o gpurun_out/prof_mega python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
