#!/bin/bash
# developer tool: round-2 session 1 -- observed fast-mode mismatches, parity kernel profile, baselines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 600 python tools/gpu_mismatch.py > gpurun_out/r02_mismatch.log 2>&1; tail -5 gpurun_out/r02_mismatch.log
python bench.py --steps 2 --warmup 3 --mode parity --spp 256 --no-cpu-baseline > gpurun_out/r02_parity256_s1.json 2>gpurun_out/r02_parity256_s1.err; python -c "import json;d=json.load(open('gpurun_out/r02_parity256_s1.json'));print('parity256', d['value'], d['e2e'])"
python bench.py --steps 2 --warmup 3 --mode parity --no-cpu-baseline > gpurun_out/r02_parity2048_s1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_parity2048_s1.json'));print('parity2048', d['value'], d['e2e'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_fast_s1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_fast_s1.json'));print('fast', d['value'], d['e2e'], d['bundle_cull']['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_parity python bench.py --steps 1 --warmup 1 --mode parity --spp 16 --no-cpu-baseline > gpurun_out/r02_ncu_parity.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo SESSION_DONE
