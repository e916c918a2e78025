#!/bin/bash
# developer tool: round-2 session 2 -- flat parity replay: tests + throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hits.py tests/test_gpu_radiance.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/r02_pytest_s2.txt 2>&1; tail -5 gpurun_out/r02_pytest_s2.txt
TPT_PARITY_FLAT=0 timeout 600 python -m pytest tests/test_gpu_hits.py tests/test_gpu_radiance.py -m gpu -q -x -k parity > gpurun_out/r02_pytest_s2_noflat.txt 2>&1; tail -3 gpurun_out/r02_pytest_s2_noflat.txt
for v in A B; do
python bench.py --steps 2 --warmup 3 --mode parity --spp 512 --variant $v --no-cpu-baseline > gpurun_out/r02_parity512_${v}_s2.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_parity512_${v}_s2.json'));print('parity512 $v', d['value'], d['e2e']['value'], d['bundle_cull']['value'])"
done
TPT_PARITY_FLAT=0 python bench.py --steps 2 --warmup 3 --mode parity --spp 512 --no-cpu-baseline > gpurun_out/r02_parity512_noflat_s2.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_parity512_noflat_s2.json'));print('parity512 noflat', d['value'])"
python bench.py --steps 2 --warmup 3 --mode parity --spp 512 --kernel mega --no-cpu-baseline > gpurun_out/r02_parity512_mega_s2.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_parity512_mega_s2.json'));print('parity512 mega', d['value'])"
for v in A B; do
python bench.py --steps 3 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r02_fast_${v}_s2.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_fast_${v}_s2.json'));print('fast $v', d['value'], d['e2e']['value'], d['bundle_cull']['value'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_parity_flat python bench.py --steps 1 --warmup 1 --mode parity --spp 16 --no-cpu-baseline > gpurun_out/r02_ncu_parity_flat.log 2>&1
echo SESSION_DONE
